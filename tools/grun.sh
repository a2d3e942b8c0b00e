#!/bin/bash
# Retry wrapper around gpurun: a busy pod answers exit code 3 (nothing charged); wait and ask again.
#   tools/grun.sh <log-file> [gpurun options] -- '<command>'
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
