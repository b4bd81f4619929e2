"""Array planner of the batch engine: jobs -> levels of trajectory descriptors, in NumPy.

A sweep of the reference is thousands of ``system(...)`` calls that differ in a few numbers: the end time and the
times of one to three multi-time operators (``two_time/correlations.py:153-170``, ``pol_entanglement/G2.py:488-530``,
``timebin/twophoton_new.py:515-557``).  :class:`JobArrays` holds exactly those numbers, one row per call;
:func:`plan_levels` turns them into launches without a Python loop over the calls:

* calls that share drive, start and initial state form a *group*; their operator lists form a trie (two calls share
  a node of depth d when their first d operator events coincide);
* the root of a group is propagated once and its full system x bond state is stored ("snapshot") at every step where
  a call's first operator acts -- and, new with this planner, so is every inner node that at least two calls share:
  a triangular sweep over (t1, t2) forks at t1 AND again at t2, so the stretch between the two operators is
  propagated once per t1 instead of once per pair;
* level 0 holds the roots, level d the depth-d nodes (they start from a snapshot of level d-1 and write snapshots
  themselves), the last level the calls.  Levels run as consecutive launches over one snapshot pool.

Exactness is the same as for one-level forking: a snapshot is the complete state, no approximation enters.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

MAX_OVR = 6
BIG = np.iinfo(np.int32).max


def _c128(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.complex128)


@dataclass
class JobArrays:
    """One row per job.  Events are a job's multi-time operators merged per step (products in file order), sorted
    by step, followed by the rows that need explicit operators only because the job's drive table ends early
    (``sb = sa = -1``); unused columns hold step ``BIG``."""
    dt: float
    t0: float                      # earliest t_start of the batch: time origin of step_shift
    packed: np.ndarray             # [n_sets, 3, n_samples] drive tables
    grid: Tuple[float, float]      # (t0, dt) of the tables
    mats: list                     # operator products referenced by ev_sb / ev_sa
    rho0s: np.ndarray              # [n_rho0, NL]
    set_id: np.ndarray
    shift: np.ndarray              # (t_start - t0) / dt
    r0: np.ndarray                 # initial-state slot
    n_steps: np.ndarray
    tail: np.ndarray               # rows kept (0 = all)
    clamp: np.ndarray              # samples of the job's own drive table (0 = the whole table)
    ev_step: np.ndarray            # [J, E]
    ev_sb: np.ndarray
    ev_sa: np.ndarray
    n_ev: np.ndarray
    n_fork: np.ndarray             # leading events that are real operators (fork candidates)

    @property
    def n_jobs(self) -> int:
        return len(self.n_steps)

    def rows(self, a: int, b: int) -> "JobArrays":
        """Jobs ``a .. b-1`` (shared tables, operators and initial states stay whole)."""
        import dataclasses
        per_job = ("set_id", "shift", "r0", "n_steps", "tail", "clamp", "ev_step", "ev_sb", "ev_sa", "n_ev", "n_fork")
        return dataclasses.replace(self, **{k: getattr(self, k)[a:b] for k in per_job})


# ------------------------------------------------------------------------------------------ from Job objects
def pack_tables(jobs):
    """Pack the jobs' drive tables into ``[n_sets, 3, n_samples]``; identical table objects share a set.  All
    tables of one batch must live on one sampling grid."""
    set_of_job, sets, key_to_set = [], [], {}
    grid = None
    nmax = 1
    for jb in jobs:
        key = tuple(id(jb.tables.get(p)) if jb.tables.get(p) is not None else 0 for p in ("x", "y", "rf"))
        if key not in key_to_set:
            key_to_set[key] = len(sets)
            sets.append(jb.tables)
            for tb in jb.tables.values():
                if tb is None:
                    continue
                g = (float(tb.t0), float(tb.dt))
                if grid is None:
                    grid = g
                elif abs(grid[0] - g[0]) > 1e-12 or abs(grid[1] - g[1]) > 1e-15:
                    raise ValueError("all drive tables of one batch must share t0 and dt")
                nmax = max(nmax, len(tb.values))
        set_of_job.append(key_to_set[key])
    if grid is None:
        grid = (0.0, 1.0)
    packed = np.zeros((len(sets), 3, nmax), dtype=np.complex128)
    n_tab = np.zeros(len(sets), dtype=np.int64)
    for s, tabs in enumerate(sets):
        for k, pol in enumerate(("x", "y", "rf")):
            tb = tabs.get(pol)
            if tb is None or len(tb.values) == 0:
                continue
            n = len(tb.values)
            n_tab[s] = max(n_tab[s], n)
            packed[s, k, :n] = tb.values
            packed[s, k, n:] = tb.values[-1]  # end value held (oracle.sample_field)
    return packed, np.asarray(set_of_job, dtype=np.int32), grid, n_tab


class MatPool:
    """Products of multi-time superoperators that act at one step, deduplicated by value."""

    def __init__(self, NL: int):
        self.NL, self.mats, self._cache = NL, [], {}

    def product_id(self, superops: Sequence[np.ndarray]) -> int:
        if not superops:
            return -1
        idkey = tuple(id(s) for s in superops)       # superoperators are cached objects (Problem.parse_mtos)
        hit = self._cache.get(idkey)
        if hit is not None:
            return hit[0]
        prod = np.eye(self.NL, dtype=complex)
        for s in superops:                           # file order: first listed acts first
            prod = s @ prod
        key = prod.tobytes()
        if key not in self._cache:
            self._cache[key] = len(self.mats)
            self.mats.append(_c128(prod))
        self._cache[idkey] = (self._cache[key], list(superops))   # keeps the operands alive: ids stay unique
        return self._cache[key]


def arrays_from_jobs(prob, jobs) -> JobArrays:
    """The general (per-job Python) route into the planner; sweeps build :class:`JobArrays` directly."""
    if not jobs:
        raise ValueError("no jobs")
    dt = float(jobs[0].dt)
    NL = prob.NL
    for jb in jobs:
        if abs(jb.dt - dt) > 1e-15:
            raise ValueError("all jobs of one batch must share dt")
    packed, set_id, grid, n_tab = pack_tables(jobs)
    t0_ref = min(jb.t_start for jb in jobs)
    J = len(jobs)
    pool = MatPool(NL)
    rho0s = [_c128(prob.rho0).reshape(NL)]
    rho0_slot: Dict[bytes, int] = {}
    shift = np.zeros(J, dtype=np.int32)
    r0 = np.zeros(J, dtype=np.int32)
    n_steps = np.zeros(J, dtype=np.int32)
    tail = np.zeros(J, dtype=np.int32)
    clamp = np.zeros(J, dtype=np.int32)
    events: List[list] = []
    n_fork = np.zeros(J, dtype=np.int32)
    for i, jb in enumerate(jobs):
        n_steps[i] = jb.n_steps
        tail[i] = jb.tail_rows or 0
        s = int(round((jb.t_start - t0_ref) / dt))
        if abs(jb.t_start - t0_ref - s * dt) > 1e-9 * max(1.0, abs(dt)):
            raise ValueError(f"t_start={jb.t_start} is not a whole number of steps (dt={dt}) after the earliest "
                             f"start {t0_ref} of the batch: run it as a separate batch")
        shift[i] = s
        if jb.rho0 is not None:
            key = _c128(jb.rho0).reshape(NL).tobytes()
            if key not in rho0_slot:
                rho0_slot[key] = len(rho0s)
                rho0s.append(_c128(jb.rho0).reshape(NL))
            r0[i] = rho0_slot[key]
        by_step: Dict[int, Tuple[list, list]] = {}
        for m in jb.mtos:
            by_step.setdefault(jb.mto_step(m), ([], []))[0 if m.before else 1].append(m.superop)
        if len(by_step) > MAX_OVR - 2:
            raise ValueError(f"more than {MAX_OVR - 2} distinct multitime-operator times in one job")
        ev = {k: (pool.product_id(bef), pool.product_id(aft)) for k, (bef, aft) in by_step.items()}
        # a run that shares a longer drive table with others must not see samples past its own last one (the
        # reference writes the pulse file of every run on np.arange(t_start, t_end, dt)): the rows whose half
        # steps reach beyond it -- the last two -- get explicit entries evaluated on the truncated table
        if 0 < jb.table_len < n_tab[set_id[i]]:
            clamp[i] = int(jb.table_len)
            for k in (jb.n_steps - 1, jb.n_steps):
                if k >= 0:
                    ev.setdefault(k, (-1, -1))
        lst = sorted(ev.items())
        nf = 0
        while nf < len(lst) and lst[nf][1] != (-1, -1):
            nf += 1
        n_fork[i] = nf
        events.append(lst)
    E = max([len(e) for e in events] + [0])
    ev_step = np.full((J, E), BIG, dtype=np.int32)
    ev_sb = np.full((J, E), -1, dtype=np.int32)
    ev_sa = np.full((J, E), -1, dtype=np.int32)
    n_ev = np.asarray([len(e) for e in events], dtype=np.int32)
    for i, lst in enumerate(events):
        for c, (k, (sb, sa)) in enumerate(lst):
            ev_step[i, c], ev_sb[i, c], ev_sa[i, c] = k, sb, sa
    return JobArrays(dt=dt, t0=t0_ref, packed=packed, grid=grid, mats=pool.mats, rho0s=np.asarray(rho0s),
                     set_id=set_id, shift=shift, r0=r0, n_steps=n_steps, tail=tail, clamp=clamp,
                     ev_step=ev_step, ev_sb=ev_sb, ev_sa=ev_sa, n_ev=n_ev, n_fork=n_fork)


# ------------------------------------------------------------------------------------------ levels
@dataclass
class Level:
    """Trajectory descriptors of one launch, as columns."""
    seqs: np.ndarray                 # [n_seq, 4] (set, step0, len, first_has_prev)
    entries: np.ndarray              # [n_ent, 6] (set, step, sb, sa, has_prev, clamp)
    seq: np.ndarray
    off: np.ndarray                  # first entry inside the sequence
    step0: np.ndarray
    n_steps: np.ndarray
    init_kind: np.ndarray
    init_index: np.ndarray
    n_ovr: np.ndarray
    ovr_step: np.ndarray             # [n, MAX_OVR]
    ovr_ent: np.ndarray
    out_from: np.ndarray
    snap_off: np.ndarray
    snap_cnt: np.ndarray
    snap_slot0: np.ndarray
    snap_steps: np.ndarray           # local rows at which snapshots are taken
    job: Optional[np.ndarray] = None     # last level: the job of each trajectory
    row0: Optional[np.ndarray] = None    # last level: rows of the job's block filled from the root
    group: Optional[np.ndarray] = None   # root level: group of each trajectory

    @property
    def n_traj(self) -> int:
        return len(self.seq)


@dataclass
class Plan:
    arrays: JobArrays
    levels: List[Level]              # the last one holds the jobs
    n_slots: int                     # snapshot pool size of the whole plan
    out_off: np.ndarray              # [J] element offset of each job's block in the output buffer
    n_rows: np.ndarray               # [J] rows kept
    out_elems: int
    copies: np.ndarray               # [n, 4] (job, n_rows, root trajectory, first row): rows taken from a root
    depth_of_job: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=np.int32))


def _unique_cols(cols, bounds=None):
    """Distinct rows of a table given as 1-D int columns: ``(inverse, first index)``, rows ordered
    lexicographically.  The columns are packed into one mixed-radix int64 key when their value ranges allow (they do
    for step / id columns) and small key spaces are resolved with a lookup table instead of a sort.  ``bounds``:
    known inclusive (lo, hi) per column, saving the min / max passes."""
    n = len(cols[0])
    if n == 0:
        return np.zeros(0, dtype=np.int64), np.zeros(0, dtype=np.int64)
    key, total = None, 1
    for k, c in enumerate(cols):
        lo, hi = bounds[k] if bounds is not None else (int(c.min()), int(c.max()))
        if hi == lo:
            continue                                   # a constant column does not distinguish rows
        total *= hi - lo + 1
        if total >= (1 << 62):
            key = None
            break
        term = c - lo if lo else c
        key = term if key is None else key * (hi - lo + 1) + term
    if total == 1:
        return np.zeros(n, dtype=np.int64), np.zeros(1, dtype=np.int64)
    if key is not None and total <= max(1 << 20, 8 * n):
        # small key space: a lookup table instead of a sort (the first occurrence wins the reversed assignment)
        pos = np.full(total, -1, dtype=np.int64)
        pos[key[::-1]] = np.arange(n - 1, -1, -1, dtype=np.int64)
        used = np.flatnonzero(pos >= 0)
        first = pos[used]
        pos[used] = np.arange(len(used), dtype=np.int64)
        return pos[key], first
    if key is not None:
        _, first, inv = np.unique(key, return_index=True, return_inverse=True)
        return inv, first
    order = np.lexsort(cols[::-1])
    new = np.ones(n, dtype=bool)
    for c in cols:
        cs = c[order]
        new[1:] |= cs[1:] != cs[:-1]
    new[0] = True
    inv = np.empty(n, dtype=np.int64)
    inv[order] = np.cumsum(new) - 1
    return inv, order[new]


def _entries_for(set_id, shift, step, sb, sa, clamp, bounds=None):
    """Deduplicated explicit-operator entries for (trajectory, event) pairs; returns (entries, id per pair)."""
    cols = [set_id, shift + step, sb, sa, (step > 0).astype(np.int64), clamp]
    inv, first = _unique_cols(cols, bounds)
    return np.stack([c[first] for c in cols], axis=1) if len(first) else np.zeros((0, 6), dtype=np.int64), inv


def plan_levels(arr: JobArrays, n_out: int, fork: bool = True) -> Plan:
    J = arr.n_jobs
    E = arr.ev_step.shape[1]
    n_steps = arr.n_steps.astype(np.int64)
    n_rows = np.where(arr.tail > 0, np.minimum(n_steps + 1, arr.tail), n_steps + 1).astype(np.int64)
    g0 = n_steps + 1 - n_rows                                  # first global row kept
    out_off = np.zeros(J, dtype=np.int64)
    out_off[1:] = np.cumsum(n_rows[:-1] * n_out)
    out_elems = int(np.sum(n_rows * n_out))
    if np.any(arr.n_ev > MAX_OVR):
        raise ValueError(f"more than {MAX_OVR} explicit operator rows in one job")

    grp, grp_first = _unique_cols([arr.set_id.astype(np.int64), arr.shift.astype(np.int64), arr.r0.astype(np.int64)])
    G = len(grp_first)
    step = arr.ev_step.astype(np.int64)
    first = step[:, 0] if E else np.full(J, BIG, dtype=np.int64)
    has_first = (arr.n_fork > 0) if E else np.zeros(J, dtype=bool)
    cnt = np.bincount(grp, minlength=G)
    any_late = np.bincount(grp, weights=(has_first & (first > 0)).astype(float), minlength=G) > 0
    forked_g = (cnt > 1) & any_late & bool(fork)
    forked = forked_g[grp]

    # ---- trie: node id of every job at every depth, and which nodes are worth a trajectory of their own
    D = int(arr.n_fork.max()) if (E and J) else 0              # deepest fork event index + 1
    # event d may be forked at: a real operator, after step 0 (the root has nothing to share before it), and not in
    # the job's clamped last rows (shared states are propagated on the full drive table)
    elig = np.zeros((J, max(D, 1)), dtype=bool)
    for d in range(D):
        elig[:, d] = (forked & (arr.n_fork > d) & (step[:, 0] > 0) &
                      ((arr.clamp == 0) | (step[:, d] < n_steps - 1)))
    node = np.full((J, max(D, 1)), -1, dtype=np.int64)          # node[:, d]: id among depth-d nodes (d >= 1)
    n_nodes = [G] + [0] * max(D, 1)
    mat = [forked_g] + [np.zeros(0, dtype=bool)] * max(D, 1)    # materialised?
    parent = [np.zeros(0, dtype=np.int64)] * (max(D, 1) + 1)
    node_first = [grp_first] + [np.zeros(0, dtype=np.int64)] * max(D, 1)
    node[:, 0] = grp
    for d in range(1, D):
        # members: jobs with a fork event BEYOND the d events of the node
        m = elig[:, d] & elig[:, d - 1]
        idx = np.nonzero(m)[0]
        if len(idx) == 0:
            D = d
            break
        inv, fst = _unique_cols([node[:, d - 1][idx], step[:, d - 1][idx], arr.ev_sb[:, d - 1][idx].astype(np.int64),
                                 arr.ev_sa[:, d - 1][idx].astype(np.int64)])
        node[idx, d] = inv
        n_nodes[d] = len(fst)
        node_first[d] = idx[fst]
        parent[d] = node[idx[fst], d - 1]
        # a node pays off when at least two jobs continue from it; its parent must exist (closure)
        mat[d] = (np.bincount(inv, minlength=len(fst)) >= 2) & mat[d - 1][parent[d]]
        node[idx[~mat[d][inv]], d] = -1
        if not mat[d].any():
            D = d
            break
    D = max(D, 1) if E else 0

    # ---- where every job starts: the deepest node it may use
    depth = np.full(J, -1, dtype=np.int64)
    for d in range(D):
        ok = elig[:, d] & (node[:, d] >= 0)
        if d == 0:
            ok &= forked
        else:
            ok &= mat[d][np.maximum(node[:, d], 0)] & (g0 >= step[:, d]) & (depth == d - 1)
        depth[ok] = d
    jidx = np.arange(J)
    start = np.where(depth >= 0, step[jidx, np.maximum(depth, 0)] if E else 0, 0)

    # ---- snapshot requests (depth, node, step): from jobs and from materialised child nodes
    BASE = int(n_steps.max()) + 2 if J else 2
    req_d, req_n, req_s = [], [], []
    sel = depth >= 0
    req_d.append(depth[sel]); req_n.append(node[jidx[sel], depth[sel]]); req_s.append(start[sel])
    node_start = [np.zeros(G, dtype=np.int64)] + [None] * D
    for d in range(1, D):
        keep = np.nonzero(mat[d])[0]
        s_d = step[node_first[d], d - 1]                        # the node's own (last) event step = its start
        node_start[d] = s_d
        req_d.append(np.full(len(keep), d - 1)); req_n.append(parent[d][keep]); req_s.append(s_d[keep])
    req = [np.concatenate(req_d).astype(np.int64), np.concatenate(req_n).astype(np.int64),
           np.concatenate(req_s).astype(np.int64)]
    _, req_first = _unique_cols(req)                            # sorted by (depth, node, step): slot id = row
    slots = np.stack([c[req_first] for c in req], axis=1) if len(req_first) else np.zeros((0, 3), dtype=np.int64)
    n_slots = len(slots)
    slot_key = (slots[:, 0] * (max(n_nodes) + 1) + slots[:, 1]) * BASE + slots[:, 2] if n_slots else np.zeros(0, np.int64)

    def slot_of(d, n, s):
        return np.searchsorted(slot_key, (d * (max(n_nodes) + 1) + n) * BASE + s)

    levels: List[Level] = []
    copies = np.zeros((0, 4), dtype=np.int64)
    root_traj_of_group = np.full(G, -1, dtype=np.int64)
    gset, gshift, gr0 = arr.set_id[grp_first].astype(np.int64), arr.shift[grp_first].astype(np.int64), arr.r0[grp_first].astype(np.int64)

    def seqs_for(set_of, shift_of, end_of):
        """One MTO-free operator sequence per (set, shift): trajectories index into it by absolute group step."""
        inv, fst = _unique_cols([set_of, shift_of])
        ln = np.zeros(len(fst), dtype=np.int64)
        np.maximum.at(ln, inv, end_of + 1)
        return np.stack([set_of[fst], shift_of[fst], ln, np.zeros(len(fst), dtype=np.int64)], axis=1), inv

    for d in range(D):
        rows = np.nonzero(slots[:, 0] == d)[0] if n_slots else np.zeros(0, dtype=np.int64)
        if len(rows) == 0:
            continue
        inv_d, fst_d = _unique_cols([slots[rows, 1]])
        nodes_d = slots[rows, 1][fst_d]
        cnt_d = np.bincount(inv_d, minlength=len(nodes_d))
        s0 = node_start[d][nodes_d]
        last = np.zeros(len(nodes_d), dtype=np.int64)
        np.maximum.at(last, inv_d, slots[rows, 2])
        n_t = len(nodes_d)
        if d == 0:
            set_of, shift_of, init_kind, init_index = gset[nodes_d], gshift[nodes_d], np.zeros(n_t, np.int64), gr0[nodes_d]
            root_traj_of_group[nodes_d] = np.arange(n_t)
            n_ovr = np.zeros(n_t, dtype=np.int64)
            ovr_step = np.zeros((n_t, MAX_OVR), dtype=np.int64)
            ovr_ent = np.zeros((n_t, MAX_OVR), dtype=np.int64)
            entries = np.zeros((0, 6), dtype=np.int64)
            out_from = np.zeros(n_t, dtype=np.int64)            # roots keep every row (jobs may need rows before their fork)
        else:
            j_of = node_first[d][nodes_d]                       # a job that carries the node's events
            set_of, shift_of = arr.set_id[j_of].astype(np.int64), arr.shift[j_of].astype(np.int64)
            init_kind = np.ones(n_t, dtype=np.int64)
            init_index = slot_of(d - 1, parent[d][nodes_d], s0)
            entries, ent_id = _entries_for(set_of, shift_of, s0, arr.ev_sb[j_of, d - 1].astype(np.int64),
                                           arr.ev_sa[j_of, d - 1].astype(np.int64), np.zeros(n_t, dtype=np.int64))
            n_ovr = np.ones(n_t, dtype=np.int64)
            ovr_step = np.zeros((n_t, MAX_OVR), dtype=np.int64)
            ovr_ent = np.zeros((n_t, MAX_OVR), dtype=np.int64)
            ovr_ent[:, 0] = ent_id
            out_from = last - s0                                # inner nodes keep their last row only
        seqs, seq = seqs_for(set_of, shift_of, last)
        levels.append(Level(seqs=seqs, entries=entries, seq=seq, off=s0, step0=s0, n_steps=last - s0,
                            init_kind=init_kind, init_index=init_index, n_ovr=n_ovr, ovr_step=ovr_step, ovr_ent=ovr_ent,
                            out_from=out_from, snap_off=fst_d.astype(np.int64), snap_cnt=cnt_d,
                            snap_slot0=rows[fst_d], snap_steps=slots[rows, 2] - s0[inv_d],
                            group=nodes_d if d == 0 else None))

    # ---- the jobs
    e0 = np.maximum(depth, 0)                                   # first event a job applies itself
    n_ovr = arr.n_ev.astype(np.int64) - e0
    ovr_ent = np.zeros((J, MAX_OVR), dtype=np.int64)
    ovr_step = np.zeros((J, MAX_OVR), dtype=np.int64)
    entries = np.zeros((0, 6), dtype=np.int64)
    if E:
        # (job, event) pairs the jobs apply themselves: everything as [J, E] tables, one packed key per pair
        col = np.arange(E, dtype=np.int64)[None, :]
        valid = (col >= e0[:, None]) & (col < arr.n_ev[:, None])
        nm = len(arr.mats) + 2
        shift64, set64 = arr.shift.astype(np.int64), arr.set_id.astype(np.int64)
        n_abs = int(shift64.max()) + BASE
        clamped = bool(arr.clamp.any())
        n_cl = int(arr.clamp.max()) + 1 if clamped else 1
        safe = np.where(valid, step, 0)
        key = (set64[:, None] * n_abs + (shift64[:, None] + safe)) * nm + (arr.ev_sb + 1)
        key = (key * nm + (arr.ev_sa + 1)) * 2 + (safe > 0)
        if clamped:
            key = key * n_cl + np.where(safe >= n_steps[:, None] - 1, arr.clamp[:, None], 0)
        total = (int(set64.max()) + 1) * n_abs * nm * nm * 2 * n_cl
        if total >= (1 << 62):
            raise ValueError("operator-entry key space too large for one batch")
        inv, first = _unique_cols([key[valid]], bounds=[(0, total - 1)])
        uk = key[valid][first]
        entries = np.zeros((len(uk), 6), dtype=np.int64)
        if clamped:
            uk, entries[:, 5] = np.divmod(uk, n_cl)
        uk, entries[:, 4] = np.divmod(uk, 2)
        uk, sa1 = np.divmod(uk, nm)
        uk, sb1 = np.divmod(uk, nm)
        entries[:, 0], entries[:, 1] = np.divmod(uk, n_abs)
        entries[:, 2], entries[:, 3] = sb1 - 1, sa1 - 1
        ent2d = np.zeros((J, E), dtype=np.int64)
        ent2d[valid] = inv
        loc2d = np.where(valid, step - start[:, None], 0)
        for v in range(min(E, int(e0.max()) + 1)):     # a job's own events start at column e0 of its event list
            rows = e0 == v
            if v == 0 and rows.all():
                ovr_ent[:, :E], ovr_step[:, :E] = ent2d, loc2d
                break
            ovr_ent[rows, :E - v], ovr_step[rows, :E - v] = ent2d[rows, v:], loc2d[rows, v:]
    from_root = (depth == 0) & (g0 < start)                     # rows g0 .. start-1 come from the root
    row0 = np.where(from_root, start - g0, 0)
    out_from = np.where(depth >= 0, np.maximum(g0 - start, 0), g0)
    if from_root.any():
        k = np.nonzero(from_root)[0]
        copies = np.stack([k, start[k] - g0[k], root_traj_of_group[grp[k]], g0[k]], axis=1)
    seqs, seq = seqs_for(arr.set_id.astype(np.int64), arr.shift.astype(np.int64), n_steps)
    init_index = np.where(depth >= 0, slot_of(np.maximum(depth, 0), node[jidx, np.maximum(depth, 0)], start),
                          arr.r0.astype(np.int64))
    zeros = np.zeros(J, dtype=np.int64)
    levels.append(Level(seqs=seqs, entries=entries, seq=seq, off=start, step0=start, n_steps=n_steps - start,
                        init_kind=(depth >= 0).astype(np.int64), init_index=init_index, n_ovr=n_ovr, ovr_step=ovr_step,
                        ovr_ent=ovr_ent, out_from=out_from, snap_off=zeros, snap_cnt=zeros, snap_slot0=zeros,
                        snap_steps=np.zeros(0, dtype=np.int64), job=jidx, row0=row0))
    return Plan(arrays=arr, levels=levels, n_slots=n_slots, out_off=out_off, n_rows=n_rows, out_elems=out_elems,
                copies=copies, depth_of_job=depth)


# ------------------------------------------------------------------------------------------ from sweep arrays
def arrays_from_sweep(prob, *, dt: float, t_start: float, t_end: np.ndarray, superops: Sequence[np.ndarray],
                      before: Sequence[bool], mto_times: np.ndarray, tails, tables: dict,
                      table_len: Optional[np.ndarray] = None) -> JobArrays:
    """:class:`JobArrays` of a sweep whose jobs differ only in their end time and in the times of the SAME ``M``
    multi-time operators (``mto_times[J, M]``, NaN = this job does not apply the operator) -- no Python loop over
    the jobs.  Semantics as :func:`arrays_from_jobs`: operators that act at one step are merged into one product in
    file (= column) order, 'before' and 'after' operators separately."""
    t_end = np.asarray(t_end, dtype=float)
    J, M = len(t_end), len(superops)
    NL = prob.NL
    n_steps = np.rint((t_end - t_start) / dt).astype(np.int64)       # ACE: N = round((te - ta)/dt)
    times = np.asarray(mto_times, dtype=float).reshape(J, M)
    absent = np.isnan(times)
    k = np.rint(np.where(absent, 0.0, (times - t_start) / dt)).astype(np.int64)
    off_grid = ~absent & (np.abs(t_start + k * dt - times) > 1e-6 * max(1.0, abs(dt)))
    if off_grid.any():
        raise ValueError(f"multitime operator time {times[off_grid][0]} is not on the dt grid")
    outside = ~absent & ((k < 0) | (k > n_steps[:, None]))
    if outside.any():
        j = int(np.nonzero(outside.any(axis=1))[0][0])
        raise ValueError(f"multitime operator time {times[j][outside[j]][0]} outside [{t_start}, {t_end[j]}]")
    jidx = np.arange(J)
    E = M + 2
    ev_step = np.full((J, E), BIG, dtype=np.int64)
    code_b = np.zeros((J, E), dtype=np.int64)
    code_a = np.zeros((J, E), dtype=np.int64)
    n_ev = np.zeros(J, dtype=np.int64)
    if M:
        key = np.where(absent, BIG, k)
        order = np.argsort(key, axis=1, kind="stable")                 # by step, file order within a step
        ss = np.take_along_axis(key, order, axis=1)
        ei = np.zeros(J, dtype=np.int64) - 1
        for c in range(M):
            live = ss[:, c] < BIG
            new = live & ((ss[:, c] != ss[:, c - 1]) if c else True)
            ei = ei + new
            rows = jidx[live]
            e = ei[live]
            ev_step[rows, e] = ss[live, c]
            o = order[live, c]
            isb = np.asarray(before, dtype=bool)[o]
            cb, ca = code_b[rows, e], code_a[rows, e]
            code_b[rows, e] = np.where(isb, cb * (M + 1) + o + 1, cb)
            code_a[rows, e] = np.where(isb, ca, ca * (M + 1) + o + 1)
        n_ev = ei + 1
    if np.any(n_ev > MAX_OVR - 2):
        raise ValueError(f"more than {MAX_OVR - 2} distinct multitime-operator times in one job")
    # operator products of every distinct (ordered) combination
    pool = MatPool(NL)
    ev_sb = np.full((J, E), -1, dtype=np.int64)
    ev_sa = np.full((J, E), -1, dtype=np.int64)
    for code, dest in ((code_b, ev_sb), (code_a, ev_sa)):
        for c in np.unique(code):
            if c == 0:
                continue
            digits, r = [], int(c)
            while r:
                r, d = divmod(r, M + 1)
                digits.append(d - 1)
            dest[code == c] = pool.product_id([superops[d] for d in digits[::-1]])
    # drive tables: one set; jobs that must not see samples past their own pulse file get clamp rows
    fake = [type("J", (), {"tables": tables})()]
    packed, _, grid, n_tab = pack_tables(fake)
    clamp = np.zeros(J, dtype=np.int64)
    n_fork = n_ev.copy()
    if table_len is not None:
        tl = np.asarray(table_len, dtype=np.int64)
        clamp = np.where((tl > 0) & (tl < n_tab[0]), tl, 0)
        for row in (n_steps - 1, n_steps):                             # ascending: they stay sorted behind the events
            need = (clamp > 0) & (row >= 0) & ~np.any(ev_step == row[:, None], axis=1)
            # an operator AT the last row sorts behind the clamp row n_steps - 1
            late = need & np.any((ev_step > row[:, None]) & (ev_step < BIG), axis=1)
            plain = need & ~late
            ev_step[jidx[plain], n_ev[plain]] = row[plain]
            if late.any():
                for j in np.nonzero(late)[0]:                          # rare: shift the later event one column right
                    e = int(n_ev[j])
                    p = int(np.searchsorted(ev_step[j, :e], row[j]))
                    for a in (ev_step, ev_sb, ev_sa):
                        a[j, p + 1:e + 1] = a[j, p:e].copy()
                    ev_step[j, p], ev_sb[j, p], ev_sa[j, p] = row[j], -1, -1
            n_ev = n_ev + need
        real = (ev_sb >= 0) | (ev_sa >= 0)
        n_fork = np.where(real.all(axis=1), E, np.argmin(real, axis=1))
        n_fork = np.minimum(n_fork, n_ev)
    z = np.zeros(J, dtype=np.int32)
    tails = np.broadcast_to(np.asarray(tails, dtype=np.int32), (J,))
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)
    return JobArrays(dt=float(dt), t0=float(t_start), packed=packed, grid=grid, mats=pool.mats,
                     rho0s=_c128(prob.rho0).reshape(1, NL), set_id=z, shift=z, r0=z, n_steps=i32(n_steps), tail=i32(tails),
                     clamp=i32(clamp), ev_step=i32(ev_step), ev_sb=i32(ev_sb), ev_sa=i32(ev_sa), n_ev=i32(n_ev),
                     n_fork=i32(n_fork))
