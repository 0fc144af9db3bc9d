set -x
cd $GRAFT_REPO_ROOT
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8_fasv3.json 2> gpurun_out/bench_n8_fasv3.err; echo "bench fas rc=$?"; cut -c1-1500 gpurun_out/bench_n8_fasv3.json; tail -3 gpurun_out/bench_n8_fasv3.err | cut -c1-500
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --method nmg --n 216 --strong > gpurun_out/bench_n8_nmg_216.json 2> gpurun_out/bench_n8_nmg_216.err; echo "bench nmg216 rc=$?"; cut -c1-1800 gpurun_out/bench_n8_nmg_216.json; tail -3 gpurun_out/bench_n8_nmg_216.err | cut -c1-500
