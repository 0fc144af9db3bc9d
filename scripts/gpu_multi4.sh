set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_multi.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench4 rc=$?"; cat gpurun_out/bench_n4.json | cut -c1-700; tail -3 gpurun_out/bench_n4.err | cut -c1-300
