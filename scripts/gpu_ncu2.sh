set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches_nmg.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --method fasv3"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_fas.csv $CMD2 > gpurun_out/ncu2.log 2>&1
echo "ncu launches rc=$?"
