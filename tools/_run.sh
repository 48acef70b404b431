cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r37_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r37_pytest.log
tail -4 gpurun_out/r37_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
