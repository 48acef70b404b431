cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r31_bench.log 2>&1; tail -1 gpurun_out/r31_bench.log | cut -c1-600
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r31_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r31_ncu_launch.log 2>&1; tail -2 gpurun_out/r31_ncu_launch.log | cut -c1-300
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pair_tc --launch-skip 18 --launch-count 1 -f -o gpurun_out/r31_pair_c64 python tools/hift_timeline.py --reps 1 > gpurun_out/r31_ncu1.log 2>&1; tail -2 gpurun_out/r31_ncu1.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pair_tc --launch-skip 6 --launch-count 1 -f -o gpurun_out/r31_pair_c128 python tools/hift_timeline.py --reps 1 > gpurun_out/r31_ncu2.log 2>&1; tail -2 gpurun_out/r31_ncu2.log | cut -c1-200
timeout 120 python tools/hift_debug.py --kind init --operand fp16 --T 300 50 > gpurun_out/r31_debug_init.log 2>&1; grep -E "WAV" gpurun_out/r31_debug_init.log
timeout 120 python tools/hift_debug.py --kind unit --operand fp16 --T 300 50 > gpurun_out/r31_debug_unit.log 2>&1; grep -E "WAV" gpurun_out/r31_debug_unit.log
