"""Throughput of the time-local-map chain kernel on a (t, tau) grid of the size the reference's
Fortran/OpenMP helper handles (calc_onetime_parallel, propagate_tau.f90:110-187), next to the
NumPy restatement on a bounded sample.  Run on a GPU box."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
from pyaceqd_b200.engine import default_engine
from pyaceqd_b200.two_time import propagate_tau_module as gm
import tlmap_oracle as fo

dim = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n_t = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
n_tau = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
NL = dim * dim
rng = np.random.default_rng(0)
n_full = n_t + n_tau + 1
dm = rng.standard_normal((NL, NL, n_full - 1)) + 1j * rng.standard_normal((NL, NL, n_full - 1))
dm *= 0.9 / np.sqrt(2 * NL)
dm = np.asfortranarray(dm)
rho = rng.standard_normal(NL) + 1j * rng.standard_normal(NL)
A, B, C = (rng.standard_normal((dim, dim)) + 1j * rng.standard_normal((dim, dim)) for _ in range(3))
times = np.round(0.1 * np.arange(n_full), 6)
sparse = times[:n_t]
eng = default_engine(0)
gm.calc_onetime_parallel(dm, rho, n_tau, dim, A, B, C, times, sparse)
t = time.perf_counter(); res = gm.calc_onetime_parallel(dm, rho, n_tau, dim, A, B, C, times, sparse); wall = time.perf_counter() - t
k_ms = eng.tlmap_last_ms()
steps = n_t * n_tau
ns = min(n_t, 8)
t = time.perf_counter(); ref = fo.calc_onetime(dm, rho, n_tau, dim, A, B, C, times, sparse[:ns]); cpu = time.perf_counter() - t
print(json.dumps({"what": "tl-map chains calc_onetime_parallel", "NL": NL, "n_t": n_t, "n_tau": n_tau,
                  "kernel_ms": k_ms, "wall_ms": 1e3 * wall, "matvec_steps_per_s_kernel": steps / (k_ms * 1e-3),
                  "l2_read_GBps_kernel": steps * NL * NL * 16 / (k_ms * 1e-3) / 1e9,
                  "gflops_kernel": steps * 8 * NL * NL / (k_ms * 1e-3) / 1e9,
                  "numpy_restatement_steps_per_s_1core": ns * n_tau / cpu,
                  "max_abs_diff_vs_restatement": float(np.abs(res[:ns] - ref).max())}))
