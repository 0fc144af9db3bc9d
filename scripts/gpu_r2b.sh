set -x
cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_asm.py 100 10 > gpurun_out/bench_asm.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_asm.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pdl or fused" > gpurun_out/pytest_pdl.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_pdl.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:element_kernel -c 4 -o gpurun_out/prof_fused python scripts/bench_asm.py 100 1 > gpurun_out/ncu_fused.log 2>&1; echo "ncu rc=$?"
