#!/bin/bash
# ncu evidence for the step kernel (run under gpurun; B200_PROFILING.md recipe).
# usage: scripts/ncu_step.sh <tag> [bench args...]
set -u
TAG=${1:-r01}; shift || true
ARGS="--steps 1 --warmup 1 --n-steps 40 --no-cpu $*"
mkdir -p gpurun_out
python bench.py $ARGS > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
python bench.py $ARGS > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_step_dmma -s 1 -c 1 -f -o gpurun_out/${TAG}_step python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
ls -la gpurun_out/
