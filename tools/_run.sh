cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r25_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r25_pytest.log
tail -5 gpurun_out/r25_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r25_bench.log 2>&1; tail -2 gpurun_out/r25_bench.log
