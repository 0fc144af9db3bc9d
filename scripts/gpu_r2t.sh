cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_dynamic_cloth.py -m gpu -x -q -k "exp2" 2>&1 | tail -25
