#!/bin/bash
# build the in-tree library (hash-stamped, no-op when current) and run a command on a B200 box via gpurun
# usage: [GPUS=N] [RETRIES=k] tools/gpu.sh <timeout_s> '<command>'      (exit 3 = no box free: retried after a pause, nothing charged)
cd /root/repo
python -m easynlp_b200.build > /dev/null || exit 1
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
n=${RETRIES:-8}
while :; do
  timeout $((T + 1900)) /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] || [ $n -le 0 ]; then exit $rc; fi
  n=$((n - 1)); sleep 150
done
