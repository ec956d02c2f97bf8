#!/bin/bash
# Retry gpurun while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh [gpurun args...]
for attempt in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    sleep 90
done
exit 3
