import sys, time
sys.path.insert(0, '.')
from pyaceqd_b200.engine import default_engine
eng = default_engine(0)
for it in (2000, 20000):
    print("DMMA peak TFLOP/s", it, eng.fp64_peak("dmma", it), flush=True)
    print("DFMA peak TFLOP/s", it, eng.fp64_peak("dfma", it), flush=True)
