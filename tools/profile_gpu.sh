#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/profile_gpu.sh TAG'): plain bench first (must exit 0), then the ncu launch
# list of the same command, then one `--set full` capture of the two launches of the path kernel (primary phase, bounce
# phase) of one frame.  Reports land in gpurun_out/.
set -u
TAG=${1:-cur}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_l.log 2>&1
# launches of k_wf_path come in pairs (primary phase, bounce phase); skip the counting / warm-up frames, take one pair
ncu --set full --import-source on --clock-control none -k regex:k_wf_path -s 10 -c 2 -f -o $OUT/${TAG}_k_wf_path \
    python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_k_wf_path.log 2>&1
tail -c 600 $OUT/${TAG}_bench.json
