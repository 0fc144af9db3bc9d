set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --method fasv3"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k "regex:element_kernel|assemble_rows" -s 6 -c 4 -f -o gpurun_out/prof_asm $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu rc=$?"
