set -x
cd $GRAFT_REPO_ROOT
free -g > gpurun_out/box_mem.txt; nproc >> gpurun_out/box_mem.txt; cat gpurun_out/box_mem.txt
avail=$(awk '/MemAvailable/ {print int($2/1048576)}' /proc/meminfo)
echo "MemAvailable GB: $avail"
if [ "$avail" -lt 110 ]; then echo "not enough host memory for the n=216 setup (needs ~50-60 GB, wanted 110 GB headroom)"; exit 0; fi
( time timeout 1500 python bench.py --n 216 --layers 2 --no-nmg --no-cpu-baseline --steps 5 --warmup 3 ) > gpurun_out/bench_10m.json 2> gpurun_out/bench_10m.err; echo "bench10m rc=$?"; cat gpurun_out/bench_10m.json; tail -5 gpurun_out/bench_10m.err
