set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_asm.py 100 10 01 > gpurun_out/bench_asm.log 2>&1; echo "rc=$?"; grep -i "galerkin\|residual" gpurun_out/bench_asm.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"galerkin" -c 2 -o gpurun_out/prof_galerkin python scripts/bench_asm.py 100 1 01 > gpurun_out/ncu_r2l.log 2>&1; echo "ncu rc=$?"
