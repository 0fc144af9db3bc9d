set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lattice_fold or force_energy_hessian or rows_with_more or rest_data" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2i_pytest.log
timeout 600 python scripts/bench_asm.py 100 > gpurun_out/r2i_bench_asm.txt 2>&1; echo "bench_asm rc=$?"; cat gpurun_out/r2i_bench_asm.txt | tail -12
