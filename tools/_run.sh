cd $GRAFT_REPO_ROOT
timeout 120 python tools/hift_debug.py --kind unit --operand fp16 --T 300 50 > gpurun_out/r21_debug.log 2>&1; echo "rc=$?"; grep -E "WAV|nan=[1-9]" gpurun_out/r21_debug.log
timeout 120 python tools/hift_timeline.py > gpurun_out/r21_timeline.jsonl 2>&1; cat gpurun_out/r21_timeline.jsonl
for L in resblocks.6.convs1.0 resblocks.7.convs1.1 resblocks.3.convs1.0; do
VT_TC_TRACE=$L timeout 120 python tools/hift_timeline.py --reps 1 > /dev/null 2> gpurun_out/r21_trace_$L.log; head -14 gpurun_out/r21_trace_$L.log | tail -8
done
