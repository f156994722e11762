#!/bin/bash
# build the in-tree library (hash-stamped, no-op when current) and run a command on a B200 box via gpurun
# usage: tools/gpu.sh <timeout_s> '<command>'
set -e
cd /root/repo
python -m easynlp_b200.build > /dev/null
T=$1; shift
exec timeout $((T + 1900)) /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
