#!/bin/bash
# round 2, run AF: does more L1 (bounded shared-memory stack) change the L1 hit rate of the path kernel?
mkdir -p gpurun_out
M=l1tex__t_sector_hit_rate.pct,l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,launch__shared_mem_config_size,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active
for s in 0 4; do
  RTB_WF_STACK=$s timeout 600 ncu --metrics $M --clock-control none -k regex:k_wf_path -c 4 --csv --log-file gpurun_out/r2_af_stack$s.csv python bench.py --steps 1 --warmup 1 --no-cpu > /dev/null 2>&1
  echo "stack $s rc=$?"
done
