cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_asm.py 100 2>&1 | grep -E "max|force level0|level0 lattice fold  "
timeout 600 python scripts/bench_asm.py 100 2>&1 | grep -E "level0 lattice fold  "
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"element_kernel" -s 4 -c 1 -o gpurun_out/r2y_el python scripts/prof_asm.py 100 > gpurun_out/r2y_ncu.log 2>&1; echo rc=$?
