set -x
cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/pytest_scale.log 2>&1; echo "scale rc=$?"; tail -30 gpurun_out/pytest_scale.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; echo "bench rc=$?"; cat gpurun_out/bench_r2c.json; tail -5 gpurun_out/bench_r2c.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_r2c.json 2> gpurun_out/bench_ref_r2c.err; echo "ref rc=$?"; cat gpurun_out/bench_ref_r2c.json; tail -5 gpurun_out/bench_ref_r2c.err
