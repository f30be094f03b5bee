#!/bin/bash
# usage: tools/runs/gpurun_retry.sh LOG TIMEOUT -- command...   (retries while the pod answers "busy"/"transient")
LOG=$1; TMO=$2; shift 3
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun --timeout $TMO -- "$@" > $LOG 2>&1
  rc=$?
  if grep -q "status=transient\|status=busy" $LOG || [ $rc -eq 3 ]; then sleep 90; continue; fi
  break
done
exit $rc
