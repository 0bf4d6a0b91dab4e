#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (exit code 3, nothing charged).
#   tools/gpurun_retry.sh <logfile> <gpurun args...>
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then echo "attempt $i rc=$rc"; exit $rc; fi
  sleep 150
done
echo "gave up"; exit 3
