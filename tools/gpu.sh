#!/bin/bash
# build the in-tree library (hash-stamped, no-op when current) and run a command on a B200 box via gpurun
# usage: [GPUS=N] tools/gpu.sh <timeout_s> '<command>'
set -e
cd /root/repo
python -m easynlp_b200.build > /dev/null
T=$1; shift
G=""
if [ -n "$GPUS" ]; then G="--gpus $GPUS"; fi
exec timeout $((T + 1900)) /usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@"
