cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_hift_gpu.py -m gpu -x -q > gpurun_out/r39_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r39_pytest.log
tail -4 gpurun_out/r39_pytest.log
timeout 120 python tools/hift_timeline.py > gpurun_out/r39_timeline.jsonl 2>&1; cat gpurun_out/r39_timeline.jsonl
VT_TC_TRACE=resblocks.1.convs1.0 timeout 120 python tools/hift_timeline.py --reps 1 > /dev/null 2> gpurun_out/r39_trace.log; head -14 gpurun_out/r39_trace.log | tail -5
