set -x
cd $GRAFT_REPO_ROOT
timeout 120 python scripts/sanitize.py > gpurun_out/sanitize_plain.log 2>&1; echo "plain rc=$?"; tail -2 gpurun_out/sanitize_plain.log
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/sanitize_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/sanitize.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -5 gpurun_out/sanitize_racecheck.log
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "exp3 or corrected" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_new.log
