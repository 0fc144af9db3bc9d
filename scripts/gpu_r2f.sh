set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for ch in 11840 23680 47360; do
PIPE_CHUNK=$ch timeout 600 python scripts/bench_asm.py 100 10 10 > gpurun_out/bench_asm_$ch.log 2>&1; echo "rc=$?"; cat gpurun_out/bench_asm_$ch.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"element_kernel|assemble_rows" -s 6 -c 2 -o gpurun_out/prof_unfused4 python scripts/bench_asm.py 100 1 0 > gpurun_out/ncu_unfused.log 2>&1; echo "ncu rc=$?"
