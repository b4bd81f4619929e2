#!/bin/bash
# ncu --set full capture of ONE launch of a named kernel inside an arbitrary python command (run under gpurun).
# usage: scripts/ncu_kernel.sh <tag> <kernel regex> <skip> <python args...>
set -u
TAG=$1; KRE=$2; SKIP=$3; shift 3
mkdir -p gpurun_out
python "$@" > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o gpurun_out/${TAG} python "$@" > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}.ncu-rep
