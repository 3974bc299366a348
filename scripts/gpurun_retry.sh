#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>
# Retries ONLY while the pod answers "busy" (exit 3: no box or slot free, nothing charged, the command never started), every
# 3 minutes, up to 40 times.  Anything else ends the loop -- in particular a "transient" verdict: that is also what a box LOST
# under the command looks like, and re-running a command that took a box down costs a strike each time (round 2 lost its GPU
# access exactly that way).
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 180
done
exit 3
