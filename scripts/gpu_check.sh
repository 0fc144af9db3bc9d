cd $GRAFT_REPO_ROOT
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -k "assembly_async or lattice or nonlinear or whole_step" 2>&1 | tail -2
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_fas_n1_quick.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02e_bench_fas_n1_quick.json').read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['frac'],4), d['parity']['x_rel_err'], d['gpu_launches'])"
