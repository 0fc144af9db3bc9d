set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02d_pytest_full.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r02d_pytest_full.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench_fas_n1.json 2> gpurun_out/r02d_bench_fas_n1.err; echo "rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --method nmg > gpurun_out/r02d_bench_nmg_n1.json 2> gpurun_out/r02d_bench_nmg_n1.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('r02d_bench_fas_n1','r02d_bench_nmg_n1'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],1), 'gs', round(d['roofline_smoother']['frac'],3), d['parity']['x_rel_err'], d['cpu_baseline']['value'])
PY
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"element_kernel|fold_lattice" -c 6 -o gpurun_out/r02d_asm python scripts/prof_asm.py 100 > gpurun_out/r02d_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02d_launches_fas_nmg.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02d_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02d_launches_fas_nmg.csv
