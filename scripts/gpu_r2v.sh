set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2v_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 --method nmg --comm p2p --no-cpu-baseline > gpurun_out/r2v_nmg_n2_p2p.json 2> gpurun_out/r2v_nmg_n2_p2p.err; echo "bench nmg rc=$?"; tail -1 gpurun_out/r2v_nmg_n2_p2p.json | cut -c1-330
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 --comm p2p --no-cpu-baseline > gpurun_out/r2v_fas_n2_p2p.json 2> gpurun_out/r2v_fas_n2_p2p.err; echo "bench fas rc=$?"; tail -1 gpurun_out/r2v_fas_n2_p2p.json | cut -c1-330
