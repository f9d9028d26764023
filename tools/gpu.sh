#!/bin/bash
# usage: tools/gpu.sh <timeout_s> '<command>'  -- gpurun with retries while the pod answers busy (nothing is charged for those)
t=$1; shift
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  st=$(python -c "import json;print(json.load(open('/root/repo/gpurun_out/.last_call.json')).get('status'))" 2>/dev/null)
  if [ "$st" != "transient" ] && [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
