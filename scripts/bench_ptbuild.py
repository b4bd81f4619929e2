#!/usr/bin/env python
"""Time the on-device process-tensor builder (csrc/ptbuild.cu) next to the NumPy builder on the host's cores, for the
parameter sets the reference's adapters pass to ACE: tls (threshold 1e-8, t_mem 6.4, dt 0.1),
tls at threshold 1e-10, biexciton (threshold 1e-10, a_e 3 nm, t_mem 20.48, dt 0.5).  One JSON line per case.
usage: scripts/bench_ptbuild.py [--host-limit-s 120] [--svd 0|1]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from pyaceqd_b200 import pt_builder as pb

ap = argparse.ArgumentParser()
ap.add_argument("--host-limit-s", type=float, default=120.0, help="skip the host build of cases expected to take longer")
ap.add_argument("--svd", type=int, default=None)
ap.add_argument("--cases", default="tls8,tls10,bx8,bx10")
args = ap.parse_args()
if args.svd is not None:
    os.environ["ACEQD_PT_SVD"] = str(args.svd)
CASES = {
    "tls8": dict(coupling_diag=[0.0, 1.0], dt=0.1, t_mem=6.4, a_e=5.0, temperature=4.0, threshold=1e-8, host_s=1),
    "tls10": dict(coupling_diag=[0.0, 1.0], dt=0.1, t_mem=6.4, a_e=5.0, temperature=4.0, threshold=1e-10, host_s=5),
    "bx8": dict(coupling_diag=[0.0, 1.0, 1.0, 2.0], dt=0.5, t_mem=20.48, a_e=3.0, temperature=4.0, threshold=1e-8, host_s=60),
    "bx10": dict(coupling_diag=[0.0, 1.0, 1.0, 2.0], dt=0.5, t_mem=20.48, a_e=3.0, temperature=4.0, threshold=1e-10, host_s=400),
}
pb.build_qd_phonon_pt([0.0, 1.0], dt=0.1, t_mem=1.0, backend="device")     # library handles, context
for name in args.cases.split(","):
    kw = dict(CASES[name])
    host_s = kw.pop("host_s")
    t = time.perf_counter()
    dev = pb.build_qd_phonon_pt(backend="device", **kw)
    t_dev = time.perf_counter() - t
    rec = {"case": name, **{k: v for k, v in kw.items()}, "chi_device": dev.chi_max, "device_s": t_dev,
           "device_stats": dev.meta["device_build"], "host_cores": os.cpu_count()}
    if host_s <= args.host_limit_s:
        t = time.perf_counter()
        host = pb.build_qd_phonon_pt(backend="host", **kw)
        rec["host_s"] = time.perf_counter() - t
        rec["chi_host"] = host.chi_max
        rec["speedup"] = rec["host_s"] / t_dev
    print(json.dumps(rec), flush=True)
