set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 10 --warmup 3 --method nmg --edge 216 --strong > gpurun_out/bench_n8_nmg_216.json 2> gpurun_out/bench_n8_nmg_216.err; echo "bench nmg216 rc=$?"; cut -c1-2500 gpurun_out/bench_n8_nmg_216.json; tail -3 gpurun_out/bench_n8_nmg_216.err | cut -c1-500
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 8 --steps 10 --warmup 3 --method nmg > gpurun_out/bench_n8_nmg_weak.json 2> gpurun_out/bench_n8_nmg_weak.err; echo "bench nmg weak rc=$?"; cut -c1-1200 gpurun_out/bench_n8_nmg_weak.json; tail -3 gpurun_out/bench_n8_nmg_weak.err | cut -c1-500
