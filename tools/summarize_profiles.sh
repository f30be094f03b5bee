#!/bin/bash
# Runs in the build container after tools/profile_gpu.sh: turns gpurun_out/TAG_* into the tracked summaries profiles/TAG_*.
set -eu
TAG=${1:-cur}
IN=gpurun_out
OUT=profiles
mkdir -p $OUT
cp $IN/${TAG}_bench.json $OUT/${TAG}_bench.json
grep -v "^==" $IN/${TAG}_launches.csv > $OUT/${TAG}_launches.csv || true
k=0
for PHASE in primary bounce; do
  ncu -i $IN/${TAG}_k_wf_path.ncu-rep --launch-skip $k --launch-count 1 --page details 2>/dev/null | grep -v "^ *$" > $OUT/${TAG}_k_wf_path_${PHASE}_details.txt
  ncu -i $IN/${TAG}_k_wf_path.ncu-rep --launch-skip $k --launch-count 1 --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; u=rows[1]; v=rows[2]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','l1tex__throughput.avg.pct_of_peak_sustained_active','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts.avg','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_lgds.avg','SM_A.TriageCompute.l1tex__data_pipe_lsu_wavefronts_mem_shared.avg','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','dram__throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed']
for w in want:
    if w in h: print(f'{w},{u[h.index(w)]},{v[h.index(w)]}')
" > $OUT/${TAG}_k_wf_path_${PHASE}_raw.csv
  k=$((k+1))
done
python tools/ncu_lines.py $IN/${TAG}_k_wf_path.ncu-rep 50 > $OUT/${TAG}_k_wf_path_lines.txt
ls -la $OUT | grep ${TAG}
