set -x
cd $GRAFT_REPO_ROOT
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench_fas_n1.json 2> gpurun_out/final_bench_fas_n1.err; echo "rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --method nmg > gpurun_out/final_bench_nmg_n1.json 2> gpurun_out/final_bench_nmg_n1.err; echo "rc=$?"
python - <<'PY'
import json
for f in ('final_bench_fas_n1','final_bench_nmg_n1'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],1), 'gs', round(d['roofline_smoother']['frac'],3), d['parity']['x_rel_err'], d['cpu_baseline']['value'])
PY
