#!/bin/bash
# Collect the round's profile artefacts on a GPU box into gpurun_out/ (copied to profiles/ by hand afterwards).
#   tools/collect_profiles.sh <tag>
TAG=${1:-rXX}
O=gpurun_out
# 1. ncu launch list of the bench command (cold-cache, serialised: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --worker-threads 0 > $O/${TAG}_ncu_bench.log 2>&1
python tools/ncu_summary.py $O/${TAG}_launches.csv > $O/${TAG}_ncu_launch_list.txt 2>&1
# 2. one full-set capture of a whole prediction() (eager pass of profile_net: 57 kernels) for DRAM traffic
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -c 60 -o $O/${TAG}_conv_full -f \
    python tools/profile_net.py --batch 8 --iters 1 --reps 1 > $O/${TAG}_ncu_full.log 2>&1
python tools/ncu_keymetrics.py $O/${TAG}_conv_full.ncu-rep > $O/${TAG}_conv_keymetrics.txt 2>&1
rm -f $O/${TAG}_conv_full.ncu-rep     # ~75 MB: over the copy-back limit; the text summary is what is kept
# 2b. a small report (first three conv launches) that can be opened with `ncu -i` for the source page
ncu --set full --clock-control none --import-source on -k regex:conv_igemm -s 10 -c 3 -o $O/${TAG}_conv_3launches -f \
    python tools/profile_net.py --batch 8 --iters 1 --reps 1 > /dev/null 2>&1
# 3. steady-state per-op table and the phase timeline
python tools/profile_net.py --batch 8 > $O/${TAG}_per_op_b8.txt 2>&1
python tools/timeline.py run --batch 8 > $O/${TAG}_timeline_b8.txt 2>&1
