set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2m_bench_fas_n1.json 2> gpurun_out/r2m_bench_fas_n1.err; echo "rc=$?"; cut -c1-3000 gpurun_out/r2m_bench_fas_n1.json; tail -3 gpurun_out/r2m_bench_fas_n1.err
timeout 900 python bench.py --steps 10 --warmup 3 --method nmg > gpurun_out/r2m_bench_nmg_n1.json 2> gpurun_out/r2m_bench_nmg_n1.err; echo "rc=$?"; cut -c1-1500 gpurun_out/r2m_bench_nmg_n1.json; tail -3 gpurun_out/r2m_bench_nmg_n1.err
