#!/bin/bash
# Runs ON THE GPU BOX (gpurun -- 'bash tools/profile_gpu.sh TAG'): plain bench first (must exit 0), then the ncu launch
# list of the same command, then one `--set full` capture of each persistent kernel.  Reports land in gpurun_out/.
set -u
TAG=${1:-cur}
OUT=gpurun_out
mkdir -p $OUT
python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err || { echo "bench failed"; tail -5 $OUT/${TAG}_bench.err; exit 1; }
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_l.log 2>&1
for K in k_wf_bounce k_wf_trace; do
  ncu --set full --import-source on --clock-control none -k regex:$K -c 1 -s 3 -f -o $OUT/${TAG}_$K \
      python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/${TAG}_ncu_$K.log 2>&1
done
tail -c 400 $OUT/${TAG}_bench.json
