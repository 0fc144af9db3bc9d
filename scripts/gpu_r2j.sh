set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_asm.py 100 10 01 > gpurun_out/bench_asm.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_asm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"element_kernel|galerkin" -s 30 -c 3 -o gpurun_out/prof_r2j python scripts/bench_asm.py 100 1 01 > gpurun_out/ncu_r2j.log 2>&1; echo "ncu rc=$?"
