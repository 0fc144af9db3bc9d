set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 1850 -c 1100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:row_pass -s 1745 -c 3 -f -o gpurun_out/prof_rowpass $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu rowpass rc=$?"
ncu --set full --clock-control none --import-source on -k "regex:element_kernel|assemble_gather|galerkin|force_gather" -s 12 -c 8 -f -o gpurun_out/prof_assembly $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu assembly rc=$?"
ls -la gpurun_out
