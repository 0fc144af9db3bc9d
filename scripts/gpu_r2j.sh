set -x
cd $GRAFT_REPO_ROOT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"element_kernel|fold_lattice|assemble_rows|force_gather" --csv --log-file gpurun_out/r2j_times.csv python scripts/prof_asm.py 100 > gpurun_out/r2j_ncu1.log 2>&1; echo rc=$?
grep -v "^==" gpurun_out/r2j_times.csv | cut -d, -f5,14,15 | tail -20
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"element_kernel|fold_lattice" -s 6 -c 2 -o gpurun_out/r2j_asm python scripts/prof_asm.py 100 > gpurun_out/r2j_ncu2.log 2>&1; echo rc=$?
ls -la gpurun_out/
