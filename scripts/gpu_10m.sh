cd $GRAFT_REPO_ROOT
timeout 1500 python bench.py --edge 216 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/final_bench_fas_10m_n1.json 2> gpurun_out/final_bench_fas_10m_n1.err; echo "rc=$?"; tail -3 gpurun_out/final_bench_fas_10m_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/final_bench_fas_10m_n1.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],1), 'gs', round(d['roofline_smoother']['frac'],3), d['config']['setup_s'])
PY
