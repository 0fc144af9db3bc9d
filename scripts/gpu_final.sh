set -x
cd $GRAFT_REPO_ROOT
( time timeout 600 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-200 gpurun_out/bench_n1.json
( time timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
