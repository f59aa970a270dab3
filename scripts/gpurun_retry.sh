#!/bin/bash
# usage: scripts/gpurun_retry.sh LOGFILE [gpurun args...] -- 'command'   : retries while the pod answers busy / transient (nothing charged)
LOG=$1; shift
for attempt in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  rc=$?
  if grep -q '"status": "transient"' gpurun_out/.last_call.json 2>/dev/null; then
    echo "$(date +%T) attempt $attempt: busy, retrying in 75 s" >> "$LOG.retries"
    sleep 75
    continue
  fi
  exit $rc
done
exit 3
