cd $GRAFT_REPO_ROOT
timeout 170 python scripts/bench_asm.py 216 > gpurun_out/r02d_bench_asm_10m.txt 2>&1; echo "asm rc=$?"
grep -v "fold plan\|min blocks 3\|min blocks 5" gpurun_out/r02d_bench_asm_10m.txt | head -30
timeout 200 python bench.py --edge 216 --steps 5 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r02d_bench_fas_10m_n1.json 2> gpurun_out/r02d_bench_fas_10m_n1.err; echo "rc=$?"; tail -3 gpurun_out/r02d_bench_fas_10m_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_fas_10m_n1.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],3), round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'frac', round(d['roofline']['frac'],4), round(d['roofline']['us_per_launch'],1), 'gs', round(d['roofline_smoother']['frac'],3), d['config']['setup_s'])
PY
