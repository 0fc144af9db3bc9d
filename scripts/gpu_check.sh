cd $GRAFT_REPO_ROOT
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_golden_reference.py tests/test_dynamic_cloth.py -m gpu -x -q -k "transfers or vcycle or nonlinear or galerkin or whole_step or golden or exp2 or three_levels or assembly_async" 2>&1 | tail -2
