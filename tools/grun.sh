#!/usr/bin/env bash
# gpurun with retry while the pod is busy (exit code 3 = nothing charged): tools/grun.sh [--gpus N] <timeout> <script> <logfile>
GP=""
if [ "$1" = "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T="$1"; S="$2"; L="$3"
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "bash $S" > "$L" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
