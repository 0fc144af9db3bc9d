cd $GRAFT_REPO_ROOT
timeout 600 python scripts/bench_asm.py 100 2>&1 | grep -E "force level0|hessian level0 lattice fold  "
