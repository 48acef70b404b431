set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r14_pytest.log
tail -15 gpurun_out/r14_pytest.log
timeout 300 python tools/hift_debug.py --kind unit --operand fp16 --T 300 50 > gpurun_out/r14_debug.log 2>&1; tail -30 gpurun_out/r14_debug.log
timeout 300 python tools/hift_timeline.py > gpurun_out/r14_timeline.jsonl 2>&1; cat gpurun_out/r14_timeline.jsonl
