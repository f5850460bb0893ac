#!/bin/bash
# usage: tools/gpu_jobs/retry.sh <gpurun args...>   — retries while the pod answers "busy" (rc 3)
for attempt in 1 2 3 4 5 6 7 8; do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[retry.sh] busy (attempt $attempt); sleeping 90 s"
    sleep 90
done
exit 3
