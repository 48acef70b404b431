# Usage on a GPU box: bash tools/bench_ngpu.sh N   (torchrun bench.py on N GPUs, JSON record into gpurun_out/)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo rc=$?
tail -1 gpurun_out/bench_${N}gpu.json | cut -c1-200
