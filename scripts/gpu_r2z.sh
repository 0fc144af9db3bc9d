cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_asm.py 100 > gpurun_out/r2z_bench_asm.txt 2>&1; echo "rc=$?"
grep -v "fold plan\|min blocks" gpurun_out/r2z_bench_asm.txt
