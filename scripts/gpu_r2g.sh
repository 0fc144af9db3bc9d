set -x
cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-600
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench_r2g.json; tail -3 gpurun_out/bench_r2g.err
