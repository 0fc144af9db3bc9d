set -x
cd $GRAFT_REPO_ROOT
run() { # name, args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --no-cpu-baseline "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -1 gpurun_out/$name.json | cut -c1-300; tail -2 gpurun_out/$name.err | cut -c1-300
}
PORT=29541 run r2w_fas_weak_n8 --steps 20 --warmup 5
PORT=29542 run r2w_nmg_weak_n8 --steps 10 --warmup 3 --method nmg
PORT=29543 run r2w_nmg_10m_strong_n8 --steps 10 --warmup 3 --method nmg --edge 216 --strong
PORT=29544 run r2w_nmg_weak_n8_nccl --steps 10 --warmup 3 --method nmg --comm nccl
