set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fas.json 2> gpurun_out/bench_fas.err; echo "bench rc=$?"; cat gpurun_out/bench_fas.json; tail -3 gpurun_out/bench_fas.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-nmg"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_fas.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
