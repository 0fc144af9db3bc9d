set -x
cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_asm.py 100 > gpurun_out/r2k_bench_asm.txt 2>&1; echo "bench_asm rc=$?"; tail -12 gpurun_out/r2k_bench_asm.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2k_pytest.log
