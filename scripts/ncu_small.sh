#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:k_step_small -s 1 -c 1 -f -o gpurun_out/r03q_small_chi16 python bench.py --steps 1 --warmup 1 --n-steps 100 --no-cpu --chi 16 > gpurun_out/r03q_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step_small -s 1 -c 1 -f -o gpurun_out/r03q_small_chi32 python bench.py --steps 1 --warmup 1 --n-steps 100 --no-cpu --chi 32 >> gpurun_out/r03q_ncu.log 2>&1
ls -la gpurun_out/r03q*
