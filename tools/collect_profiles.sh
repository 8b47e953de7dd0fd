#!/bin/bash
# Collect the round's profile artefacts on a GPU box into gpurun_out/<tag>/ (copied to profiles/ by hand afterwards).
#   tools/collect_profiles.sh <tag>
TAG=${1:-rXX}
O=gpurun_out/$TAG
mkdir -p $O
# 1. ncu launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --worker-threads 0 > $O/ncu_bench.log 2>&1
python tools/ncu_summary.py $O/launches.csv > $O/ncu_launch_list.txt 2>&1
# 2. one full-set capture of a whole eager prediction() -> per-launch table + the roofline traffic record
ncu --set full --clock-control none --import-source on -c 64 -o $O/pred_full -f \
    python tools/profile_net.py --batch 8 --iters 1 --reps 1 > $O/ncu_full.log 2>&1
python tools/ncu_traffic.py $O/pred_full.ncu-rep $O/ncu_conv_per_launch.txt $O/conv_traffic.json > $O/traffic.log 2>&1
rm -f $O/pred_full.ncu-rep     # ~80 MB: over the copy-back limit; the text summaries are what is kept
# 3. pointwise / streaming kernels: achieved GB/s (CUDA events) and the DRAM-side counters of one launch each
python tools/bench_pointwise.py > $O/pointwise_gbs.txt 2>&1
ncu --set full --clock-control none -k regex:b200 -c 40 -o $O/pointwise -f python tools/bench_pointwise.py --once > $O/ncu_pointwise.log 2>&1
python tools/ncu_keymetrics.py $O/pointwise.ncu-rep > $O/ncu_pointwise_keymetrics.txt 2>&1
rm -f $O/pointwise.ncu-rep
# 3b. the classifier head kernel alone (small report, kept)
ncu --set full --clock-control none --import-source on -k regex:head_pool -c 1 -o $O/head -f \
    python tools/profile_net.py --batch 8 --iters 1 --reps 1 > /dev/null 2>&1
# 4. steady-state per-op tables
python tools/profile_net.py --batch 8 > $O/per_op_b8.txt 2>&1
python tools/profile_net.py --batch 1 > $O/per_op_b1.txt 2>&1
