set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
timeout 1200 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -40 gpurun_out/pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -s -k "c1 or c4 or c2" > gpurun_out/pytest_scale2.log 2>&1; echo "scale rc=$?"; tail -40 gpurun_out/pytest_scale2.log | cut -c1-1500
