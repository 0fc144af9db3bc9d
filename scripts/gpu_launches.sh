cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02b_b.log 2>&1; echo "plain rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02b_launches_fas_nmg.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02b_launches_fas_nmg.csv
