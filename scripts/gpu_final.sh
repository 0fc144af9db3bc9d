set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cat gpurun_out/bench_n1.json; tail -4 gpurun_out/bench_n1.err
( time timeout 900 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cat gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 700 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"element_kernel|assemble_rows|gs_color" -s 8 -c 4 -o gpurun_out/prof_final python scripts/bench_asm.py 100 1 01 > gpurun_out/ncu_final.log 2>&1; echo "ncu full rc=$?"
