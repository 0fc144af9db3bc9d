set -x
cd $GRAFT_REPO_ROOT
N=${1:-2}
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; grep "^{" gpurun_out/bench_n$N.json | cut -c1-400; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
