set -x
cd $GRAFT_REPO_ROOT
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2u_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2u_bench_fas_n1.json 2> gpurun_out/r2u_bench_fas_n1.err; echo "rc=$?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench_fas_n1.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','parity','clocks')})
print(d['roofline']['frac'], d['roofline']['us_per_launch'], d['config'].get('host_affinity'), d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-600
