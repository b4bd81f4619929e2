#!/bin/bash
# ncu --set full capture of one step-kernel launch for an arbitrary bench configuration.
# usage: scripts/ncu_cfg.sh <tag> <skip> <bench args...>   (run under gpurun)
set -u
TAG=$1; SKIP=$2; shift 2
mkdir -p gpurun_out
python bench.py "$@" > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:k_step_dmma -s $SKIP -c 1 -f -o gpurun_out/${TAG}_step python bench.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
ls -la gpurun_out/${TAG}_step.ncu-rep
