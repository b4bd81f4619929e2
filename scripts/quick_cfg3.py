#!/usr/bin/env python
"""cfg3 G2 map in one line: grid wall ms, branch / trunk kernel ms, branch-launch fraction of the DMMA peak.
usage: scripts/quick_cfg3.py [label]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rec = bench.measure_g2_grid(0, n_t=256, chi=128, steps=5, warmup=3, dist=None, cpu=False)
print("CFG3", sys.argv[1] if len(sys.argv) > 1 else "", "wall %.3f branch %.3f trunk %.3f frac %.4f | %s | %s" % (
    rec["wall_ms"], rec["branch_kernel_ms"], rec["trunk_kernel_ms"], rec["roofline"]["frac"], rec.get("branch_kernel"), rec.get("trunk_kernel")),
    "checks", rec.get("checks"))
