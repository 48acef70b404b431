cd $GRAFT_REPO_ROOT
timeout 120 python tools/hift_debug.py --kind unit --operand fp16 --T 300 50 > gpurun_out/r33_debug.log 2>&1; echo "rc=$?"; grep -E "WAV|nan=[1-9]" gpurun_out/r33_debug.log
timeout 120 python tools/hift_timeline.py > gpurun_out/r33_timeline.jsonl 2>&1; cat gpurun_out/r33_timeline.jsonl
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r33_bench2.log 2>&1; tail -1 gpurun_out/r33_bench2.log | cut -c1-900
