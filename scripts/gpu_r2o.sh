set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 600 python scripts/bench_asm.py 100 10 0 > gpurun_out/bench_asm.log 2>&1; echo "rc=$?"; head -6 gpurun_out/bench_asm.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fas.json 2> gpurun_out/bench_fas.err; echo "bench rc=$?"; cat gpurun_out/bench_fas.json; tail -3 gpurun_out/bench_fas.err
