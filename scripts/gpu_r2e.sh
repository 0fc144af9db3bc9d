set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/pytest_multi.log 2>&1; echo "multi rc=$?"; tail -30 gpurun_out/pytest_multi.log | cut -c1-400
for m in fasv3 nmg; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --method $m > gpurun_out/bench_n2_$m.json 2> gpurun_out/bench_n2_$m.err; echo "bench $m rc=$?"; cut -c1-2500 gpurun_out/bench_n2_$m.json; tail -5 gpurun_out/bench_n2_$m.err | cut -c1-600
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --method nmg --comm torch > gpurun_out/bench_n2_nmg_torch.json 2> gpurun_out/bench_n2_nmg_torch.err; echo "bench rc=$?"; cut -c1-1200 gpurun_out/bench_n2_nmg_torch.json; tail -3 gpurun_out/bench_n2_nmg_torch.err | cut -c1-400
timeout 300 python -m pytest tests/test_gpu_scale.py -m gpu -x -q -k "c1 or c4 or c2" > gpurun_out/pytest_scale2.log 2>&1; echo "scale rc=$?"; tail -5 gpurun_out/pytest_scale2.log | cut -c1-600
