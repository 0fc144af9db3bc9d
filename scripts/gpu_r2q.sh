set -x
cd $GRAFT_REPO_ROOT
nvidia-smi topo -m | head -8
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2q_pytest.log
for c in p2p nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --method nmg --comm $c --no-cpu-baseline > gpurun_out/r2q_nmg_n2_$c.json 2> gpurun_out/r2q_nmg_n2_$c.err; echo "bench nmg $c rc=$?"; tail -1 gpurun_out/r2q_nmg_n2_$c.json | cut -c1-400; tail -2 gpurun_out/r2q_nmg_n2_$c.err | cut -c1-300
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --comm p2p --no-cpu-baseline > gpurun_out/r2q_fas_n2_p2p.json 2> gpurun_out/r2q_fas_n2_p2p.err; echo "bench fas p2p rc=$?"; tail -1 gpurun_out/r2q_fas_n2_p2p.json | cut -c1-400
