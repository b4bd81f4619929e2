"""Test double for ``aceqd_propagate_batch``: executes the LEVELS of a :class:`pyaceqd_b200.planner.Plan` with the
oracle's arithmetic (snapshot pool, explicit operator entries, kept rows, root copies), so that the planner's
forking can be checked against un-forked oracle runs without a GPU.  Test infrastructure only."""
import types

import numpy as np
from scipy.linalg import expm

import oracle


def run_plan(prob, pt, jobs, plan, t_eval="half_mid"):
    """Returns one ``[n_out, rows]`` array per job, like ``Engine.run_jobs``."""
    arr = plan.arrays
    NL, n_out = prob.NL, prob.out_w.shape[0]
    blk_of_alpha = pt.block_of_class(prob.cls_keys)[np.asarray(prob.cls)]
    tables_of_set = {}
    for j, jb in enumerate(jobs):
        tables_of_set.setdefault(int(arr.set_id[j]), jb.tables)
    snaps = {}
    out = np.zeros(plan.out_elems, dtype=complex)
    root_out = None

    def half(set_id, shift, step, which, clamp):
        job = types.SimpleNamespace(tables=tables_of_set[set_id], table_len=int(clamp))
        t_n = arr.t0 + (shift + step) * arr.dt
        t = oracle.half_step_times(t_n, arr.dt, t_eval)[which]
        return expm(oracle.liouvillian_at(prob, job, t) * (0.5 * arr.dt))

    for lv_i, lv in enumerate(plan.levels):
        last = lv_i == len(plan.levels) - 1
        rows_kept = lv.n_steps + 1 - lv.out_from
        if last:
            base = plan.out_off[lv.job] + lv.row0 * n_out
            buf = out
        else:
            base = np.zeros(lv.n_traj, dtype=np.int64)
            base[1:] = np.cumsum(rows_kept[:-1] * n_out)
            buf = np.zeros(int(np.sum(rows_kept * n_out)), dtype=complex)
        new_snaps = {}
        for b in range(lv.n_traj):
            sset, shift = int(lv.seqs[lv.seq[b], 0]), int(lv.seqs[lv.seq[b], 1])
            assert lv.off[b] + lv.n_steps[b] < lv.seqs[lv.seq[b], 2]
            ovr = {int(lv.ovr_step[b, q]): lv.entries[lv.ovr_ent[b, q]] for q in range(int(lv.n_ovr[b]))}
            if lv.init_kind[b] == 0:
                state = np.zeros((NL, 1), dtype=complex)
                state[:, 0] = arr.rho0s[lv.init_index[b]]
                closure = np.ones(1, dtype=complex)
            else:
                state, closure = (x.copy() for x in snaps[int(lv.init_index[b])])
            snap_rows = {int(lv.snap_steps[lv.snap_off[b] + q]): int(lv.snap_slot0[b] + q) for q in range(int(lv.snap_cnt[b]))}
            s0 = int(lv.step0[b])
            for i in range(int(lv.n_steps[b]) + 1):
                e = ovr.get(i)
                if e is not None:
                    assert (e[0], e[1]) == (sset, shift + s0 + i)
                    if e[2] >= 0:
                        state = arr.mats[e[2]] @ state
                if i >= lv.out_from[b]:
                    rho = state[:, :len(closure)] @ closure
                    o = base[b] + (i - int(lv.out_from[b])) * n_out
                    buf[o: o + n_out] = prob.out_w @ rho
                if i in snap_rows:
                    assert e is None
                    new_snaps[snap_rows[i]] = (state.copy(), closure.copy())
                if e is not None and e[3] >= 0:
                    state = arr.mats[e[3]] @ state
                if i == lv.n_steps[b]:
                    break
                nxt = ovr.get(i + 1)
                state = half(sset, shift, s0 + i, 0, e[5] if e is not None else 0) @ state
                s = int(pt.slice_of_step(s0 + i))
                state = oracle.apply_slice(state, pt.slices[s], blk_of_alpha)
                state = half(sset, shift, s0 + i, 1, nxt[5] if nxt is not None else 0) @ state
                closure = pt.closures[s]
        snaps.update(new_snaps)
        if lv.group is not None:
            root_out = (buf, base)
    for (job, n_copy, ti, row_first) in plan.copies:
        a = root_out[1][ti] + row_first * n_out
        out[plan.out_off[job]: plan.out_off[job] + n_copy * n_out] = root_out[0][a: a + n_copy * n_out]
    return [np.ascontiguousarray(out[o: o + r * n_out].reshape(r, n_out).T) for o, r in zip(plan.out_off, plan.n_rows)]
