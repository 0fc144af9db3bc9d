cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "assembly_async or lattice or force_energy or rows_with_more or nonlinear" > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r2z_pytest.log
timeout 600 python scripts/bench_asm.py 100 > gpurun_out/r2z_bench_asm.txt 2>&1; echo "rc=$?"
grep -v "fold plan\|min blocks 3\|min blocks 5" gpurun_out/r2z_bench_asm.txt
