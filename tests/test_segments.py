"""Segment schedule of the step kernel (more tiles than SMs: tiles laid end to end, cut into one equal
piece of steps per CTA; a tile that straddles a cut is handed from one CTA to the next through HBM).
CPU: the planner's invariants through the C ABI (aceqd_segment_plan is pure host code).
GPU: results do not depend on the schedule and match the oracle (ACEQD_SEG_SMS shrinks the CTA count so
that small batches are cut)."""
import ctypes
import os

import numpy as np
import pytest

from pyaceqd_b200 import engine as eng_mod

INF = 0x7FFFFFFF


def _plan(lens_per_tile, T, n_sm, step0=None):
    """lens_per_tile: list of lists of n_steps (<= T entries per tile)."""
    lib = eng_mod.load_library()
    lib.aceqd_segment_plan.argtypes = [ctypes.POINTER(eng_mod._Batch), ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                       ctypes.c_void_p, ctypes.POINTER(ctypes.c_int32), ctypes.POINTER(ctypes.c_int32)]
    n_tiles = len(lens_per_tile)
    trajs = np.zeros(sum(len(x) for x in lens_per_tile), dtype=eng_mod.TRAJ_DT)
    tile_traj = np.full(n_tiles * T, -1, dtype=np.int32)
    k = 0
    for i, lens in enumerate(lens_per_tile):
        for j, n in enumerate(lens):
            trajs[k]["n_steps"] = n
            trajs[k]["step0"] = 0 if step0 is None else step0[i][j]
            tile_traj[i * T + j] = k
            k += 1
    b = eng_mod._Batch()
    b.n_traj, b.trajs = len(trajs), trajs.ctypes.data
    b.tile_T, b.n_tiles, b.tile_traj = T, n_tiles, tile_traj.ctypes.data
    max_segs = 2 * n_tiles + n_sm
    segs = np.zeros((max_segs, 5), dtype=np.int32)
    off = np.zeros(n_sm + 2, dtype=np.int32)
    n_ctas, n_slots = ctypes.c_int32(), ctypes.c_int32()
    rc = lib.aceqd_segment_plan(ctypes.byref(b), n_sm, max_segs, segs.ctypes.data, off.ctypes.data,
                                ctypes.byref(n_ctas), ctypes.byref(n_slots))
    assert rc == 0, lib.aceqd_last_error()
    return segs, off[:n_ctas.value + 1], n_ctas.value, n_slots.value


def _check_plan(lens_per_tile, T, n_sm, step0=None):
    segs, off, n_ctas, n_slots = _plan(lens_per_tile, T, n_sm, step0)
    assert 1 <= n_ctas <= n_sm
    begin, end = [], []
    for i, lens in enumerate(lens_per_tile):
        s0 = [0] * len(lens) if step0 is None else step0[i]
        begin.append(min(s0) if lens else None)
        end.append(max(a + b for a, b in zip(s0, lens)) if lens else None)
    cover = {}      # tile -> list of (lo, hi, cta, position in cta, count in cta, save, load)
    load_of_cta = []
    for c in range(n_ctas):
        tot = 0
        for pos, k in enumerate(range(off[c], off[c + 1])):
            tile, lo, hi, save, load = (int(v) for v in segs[k])
            lo_c = begin[tile] if lo == -INF else lo
            hi_c = end[tile] if hi == INF else hi
            assert begin[tile] <= lo_c and hi_c <= end[tile]
            tot += max(1, hi_c - lo_c)
            cover.setdefault(tile, []).append((lo_c, hi_c, c, pos, off[c + 1] - off[c], save, load))
        load_of_cta.append(tot)
    slots_seen = set()
    for i, lens in enumerate(lens_per_tile):
        if not lens:
            assert i not in cover
            continue
        parts = sorted(cover[i])
        assert parts[0][0] == begin[i] and parts[-1][1] == end[i]
        assert len(parts) <= 2
        if len(parts) == 1:
            assert parts[0][5] == -1 and parts[0][6] == -1
        else:
            head, tail = parts
            assert head[1] == tail[0]                      # the cut
            assert head[5] >= 0 and head[5] == tail[6] and head[6] == -1 and tail[5] == -1
            assert head[5] not in slots_seen
            slots_seen.add(head[5])
            assert head[3] == 0                            # a head is the FIRST thing its CTA does
            assert tail[3] == tail[4] - 1                  # a tail the LAST thing
            assert head[2] < tail[2]                       # producer has the lower block index
            assert head[1] - head[0] >= 8 and tail[1] - tail[0] >= 8
    assert len(slots_seen) == n_slots
    total = sum(max(1, e - b) for b, e in zip(begin, end) if b is not None)
    longest = max(max(1, e - b) for b, e in zip(begin, end) if b is not None)
    ideal = max(longest, -(-total // n_sm))
    assert max(load_of_cta) <= ideal + 8 * 64              # the planner may pad the piece to avoid tiny segments
    return max(load_of_cta), ideal


def test_plan_cfg2_shape_is_balanced():
    # 256 tiles x 400 steps on 148 SMs: two waves would cost 800 steps, the wrap-around schedule 692
    worst, ideal = _check_plan([[400] * 16] * 256, 16, 148)
    assert ideal == 692 and worst <= 700


def test_plan_random_ragged_tiles():
    rng = np.random.default_rng(5)
    for n_sm in (1, 2, 3, 7, 148):
        for _ in range(6):
            n_tiles = int(rng.integers(n_sm + 1, 4 * n_sm + 8))
            T = int(rng.choice([1, 2, 4]))
            lens, s0 = [], []
            for _ in range(n_tiles):
                k = int(rng.integers(0, T + 1))
                lens.append([int(rng.integers(0, 300)) for _ in range(k)])
                s0.append([int(rng.integers(0, 50)) for _ in range(k)])
            if not any(lens):
                continue
            _check_plan(lens, T, n_sm, s0)


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture
def seg_sms():
    old = {k: os.environ.get(k) for k in ("ACEQD_SEG_SMS", "ACEQD_SEGMENTS")}

    def set_(n, on=True):
        os.environ["ACEQD_SEG_SMS"] = str(n)
        os.environ["ACEQD_SEGMENTS"] = "1" if on else "0"
    yield set_
    for k, v in old.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.gpu
@pytest.mark.parametrize("n_sm", [1, 2, 3, 5])
def test_segmented_launch_matches_oracle(engine, seg_sms, n_sm):
    import oracle
    from helpers import make_tables, sweep_jobs, tls_problem
    from pyaceqd_b200.jobs import Job
    from pyaceqd_b200.process_tensor import synthetic_growing_pt, synthetic_pt
    from pyaceqd_b200.pulses import ChirpedPulse
    seg_sms(n_sm)
    prob = tls_problem()
    # ragged lengths on a growing PT (slices change with the absolute step: the resumed segment must pick the right ones)
    pt = synthetic_growing_pt(24, len(prob.cls_keys), n_initial=5, n_repeat=3)
    p = ChirpedPulse(tau_0=1, e_start=0.5, alpha=0, t0=3, e0=2)
    jobs = [Job(0.0, te, 0.1, tables=make_tables([p], 0.0, te, 0.1))
            for te in (0.0, 0.1, 4.0, 1.0, 2.7, 5.0, 3.3, 4.9, 0.3, 5.0, 2.2, 3.0, 4.4)]
    for T in (1, 2):
        got = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=T)
        for g, jb in zip(got, jobs):
            assert np.abs(g - oracle.propagate(prob, pt, jb)).max() < 1e-10
    # operators staged in shared memory (double buffered) and chi = 64
    pt2 = synthetic_pt(64, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    jobs2 = sweep_jobs(5, 5, t_end=6.0)
    got = engine.run_jobs(prob, pt2, jobs2, kernel="dmma", tile_T=2)
    seg_sms(n_sm, on=False)
    ref = engine.run_jobs(prob, pt2, jobs2, kernel="dmma", tile_T=2)
    assert max(np.abs(a - b).max() for a, b in zip(got, ref)) == 0.0      # same arithmetic, bit for bit
    for k in (0, 7, 24):
        assert np.abs(got[k] - oracle.propagate(prob, pt2, jobs2[k])).max() < 1e-10


@pytest.mark.gpu
@pytest.mark.parametrize("n_sm", [2, 3])
def test_segmented_trunks_with_snapshots_and_branches(engine, seg_sms, n_sm):
    """Forked G2-style batches: trunk tiles take bond-state snapshots (the cursors travel with a cut tile);
    branches start from snapshots at different absolute steps; single-buffered operators at NL=16."""
    import oracle
    from helpers import biexciton_problem, make_tables
    from pyaceqd_b200.jobs import Job
    from pyaceqd_b200.process_tensor import synthetic_pt
    from pyaceqd_b200.pulses import ChirpedPulse
    prob = biexciton_problem(outputs=["|1><1|_4", "(|3><1|_4*|1><1|_4*|1><3|_4)", "|0><3|_4"])
    pt = synthetic_pt(40, len(prob.cls_keys), n_slices=2, kind="unitary", scale=0.999)
    dt, tau_max = 0.25, 5.0
    jobs = []
    for e0 in (2.0, 3.0, 4.0, 5.0, 6.0):          # five drives -> five trunks
        p = ChirpedPulse(tau_0=1.0, e_start=-2.0, alpha=0, t0=2.0, e0=e0, polar_x=0.8)
        tabs = make_tables([p], 0.0, 40.0, dt)
        for i in range(5):
            t1 = 2.5 * i + 5.0
            mt = prob.parse_mtos([{"operator": "|3><1|_4", "applyFrom": "_right", "time": t1},
                                  {"operator": "|1><3|_4", "applyFrom": "_left", "time": t1}])
            jobs.append(Job(0.0, t1 + tau_max, dt, tables=tabs, mtos=mt))
    seg_sms(n_sm)
    got = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=1)
    seg_sms(n_sm, on=False)
    ref = engine.run_jobs(prob, pt, jobs, kernel="dmma", tile_T=1)
    assert max(np.abs(a - b).max() for a, b in zip(got, ref)) == 0.0
    for k in (0, 9, 24):
        assert np.abs(got[k] - oracle.propagate(prob, pt, jobs[k])).max() < 1e-10
