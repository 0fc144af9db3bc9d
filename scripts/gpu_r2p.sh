set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_asm.py 100 10 0 > gpurun_out/bench_asm.log 2>&1; grep -i galerkin gpurun_out/bench_asm.log
bash scripts/gpu_10m.sh
