#!/usr/bin/env python
"""DMMA.8x8x4 throughput against warps per SM sub-partition (one block per SM, 8 independent accumulator chains per warp)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyaceqd_b200.engine import default_engine
eng = default_engine(0)
print("full occupancy (4 blocks x 8 warps per SM):", eng.fp64_peak("dmma", 20000))
for w in (1, 2, 3, 4, 8):
    print(f"{w} warps per sub-partition:", eng.fp64_peak(10 + w, 20000), "TFLOP/s")
