set -x
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nmg.json 2> gpurun_out/bench_nmg.err; echo "bench rc=$?"; cat gpurun_out/bench_nmg.json; tail -3 gpurun_out/bench_nmg.err
timeout 900 python bench.py --steps 10 --warmup 3 --method fasv3 --no-cpu-baseline > gpurun_out/bench_fas.json 2> gpurun_out/bench_fas.err; echo "bench rc=$?"; cat gpurun_out/bench_fas.json; tail -3 gpurun_out/bench_fas.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_nmg.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
