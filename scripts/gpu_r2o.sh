set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fas_multilevel" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2o_pytest.log
timeout 600 python scripts/bench_asm.py 100 2>&1 | grep -E "force level0|hessian level0 lattice fold  "
