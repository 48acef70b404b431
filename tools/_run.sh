cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r41_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r41_pytest.log
tail -4 gpurun_out/r41_pytest.log
timeout 120 python tools/hift_timeline.py > gpurun_out/r41_timeline.jsonl 2>&1; cat gpurun_out/r41_timeline.jsonl
