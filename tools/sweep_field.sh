for cfg in "X=1" "RTB_WF_REFILL_P=16" "RTB_WF_REFILL_P=20" "RTB_WF_REFILL_P=28" "RTB_WF_REFILL=16" "RTB_WF_REFILL=24" "RTB_WF_DESCEND=3" "RTB_WF_DESCEND=6" "RTB_WF_STACK=8"; do
  env $cfg python bench.py --steps 5 --warmup 3 --workload field1m --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), {k: round(v,3) for k,v in d['stages_ms'].items()})"
done
