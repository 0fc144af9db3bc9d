set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or force_energy_hessian or nonlinear or rows_with_more" > gpurun_out/pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_fused.log
timeout 600 python scripts/bench_fused.py 100 10 0,3,5,8,13 > gpurun_out/bench_fused.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_fused.log
