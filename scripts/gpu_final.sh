set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 20 --warmup 3 ) > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench1 rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -4 gpurun_out/bench_n1.err
timeout 600 python bench.py --method nmg --corrected-hierarchy --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_nmg_corrected.json 2> gpurun_out/bench_nmg_corrected.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_nmg_corrected.json
timeout 600 python bench.py --method nmg --layers 3 --no-cpu-baseline --steps 20 --warmup 3 > gpurun_out/bench_nmg_l3.json 2> gpurun_out/bench_nmg_l3.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_nmg_l3.json
