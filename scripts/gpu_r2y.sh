cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lattice_fold or force_energy_hessian or rows_with_more or rest_data" 2>&1 | tail -2
timeout 600 python scripts/bench_asm.py 100 2>&1 | grep -E "max|force level0|level0 lattice fold  |level1"
