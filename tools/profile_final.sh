#!/bin/bash
# Round-end evidence on one GPU box: full GPU test suite, the driver-shaped bench line, the ncu launch list of a bench
# run and five `ncu --set full` captures (C = 64 k = 7 plain pair = the dominant kernel class of the roofline object, the
# mean-fused launches at C = 64 and C = 128, a C = 128 k = 7 transposed pair, a C = 256 conv2).  Outputs under gpurun_out/ with the given tag.
#   gpurun --timeout 1500 -- bash tools/profile_final.sh r02c
TAG=${1:-r02c}
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputest.log 2>&1; tail -3 gpurun_out/${TAG}_gputest.log; fi
python bench.py > gpurun_out/${TAG}_bench_n1.json 2> gpurun_out/${TAG}_bench_n1.err; tail -c 300 gpurun_out/${TAG}_bench_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_launches_ncu.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/${TAG}_ncu_bench.log 2>&1
# one forward = 16 k_pair_tc launches: 0-9 level 1 (0 = source ResBlock k = 7 plain, 9 = mean-fused last pairs), 10 level-2
# source last pair, 11-12 k = 3 plain, 13-14 k = 7 plain, 15 mean-fused last pairs; 24 k_convT_tc launches: 13 = conv2 of the
# first k = 7 pair of the C = 256 level
for spec in "c64_k7:k_pair_tc:13" "mean_fused:k_pair_tc:15" "c128_k7:k_pair_tc:0" "c128_mean_fused:k_pair_tc:9" "c256_k7_conv2:k_convT_tc:13"; do
  name=${spec%%:*}; rest=${spec#*:}; kern=${rest%%:*}; skip=${rest##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$kern --launch-skip $skip --launch-count 1 \
    -f -o /tmp/${TAG}_${name} python tools/hift_timeline.py --reps 1 > gpurun_out/${TAG}_ncu_${name}.log 2>&1
  # the reports (30 MB each with sources) stay on the box: gpurun_out/ is capped at 64 MiB; export what the summaries need
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page details --csv > gpurun_out/${TAG}_${name}_details.csv 2>/dev/null
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page raw --csv > gpurun_out/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page source --csv 2>/dev/null | gzip -9 > gpurun_out/${TAG}_${name}_source.csv.gz
done
ls -la gpurun_out | tail -12
