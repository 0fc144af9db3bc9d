cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "horse" 2>&1 | tail -8
