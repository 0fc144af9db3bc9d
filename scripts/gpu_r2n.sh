set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fas_multilevel or nonlinear_steps" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2n_pytest.log
