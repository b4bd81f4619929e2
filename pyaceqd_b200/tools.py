"""Host helpers that define the I/O contract of the engine: time grids, operator-string
generators, density-matrix (de)composition, CSV export, concurrence.

Public names and argument meaning follow the reference's ``pyaceqd/tools.py`` (cited per
function); plotting helpers are out of scope.
"""
from __future__ import annotations

import itertools
import re

import numpy as np


# ------------------------------------------------------------------------------------ grids
def _merge_intervals(intervals):
    """Merge sorted ``[start, end]`` intervals that overlap or touch (reference ``tools.py:9-26``;
    cases pinned by ``pyaceqd/tests/test_merge_interval.py:5-23``).  Works in place and returns
    the same list, like the reference."""
    merged = []
    for lo, hi in intervals:
        if merged and lo <= merged[-1][1]:
            merged[-1][1] = max(merged[-1][1], hi)
        else:
            merged.append([lo, hi])
    intervals[:] = merged
    return intervals


def get_gaussian_t(t0, tend, *pulses, dt_max=1.0, dt_min=0.01, interval_per_step=0.05):
    """Adaptive grid: a new point whenever the accumulated pulse area since the last point
    reaches ``interval_per_step`` or ``dt_max`` has elapsed (reference ``tools.py:28-44``)."""
    fine = np.arange(t0, tend, dt_min)
    area = np.zeros(len(fine))
    for p in pulses:
        area = area + p.get_integral(fine)
    n_max = int(dt_max / dt_min)
    pts = [t0]
    since, acc = 0, 0.0
    for i in range(1, len(fine)):
        acc += area[i] - area[i - 1]
        since += 1
        if acc >= interval_per_step or since == n_max:
            pts.append(fine[i])
            since, acc = 0, 0.0
    return np.array(pts)


def construct_t(t0, tend, dt_small=0.1, dt_big=1.0, dt_exp=None, *pulses, factor_tau=4, simple_exp=False,
                gaussian_t=False, add_tend=True):
    """Time axis with spacing ``dt_small`` within ``factor_tau*tau`` of every pulse centre and
    ``dt_big`` elsewhere (reference ``tools.py:46-108``).  Note the positional ``dt_exp`` before
    ``*pulses`` -- a reference quirk that callers rely on."""
    if dt_exp is None:
        dt_exp = dt_small
    windows = []
    for p in pulses:
        if t0 < p.t0 < tend:
            windows.append([p.t0 - factor_tau * p.tau, p.t0 + factor_tau * p.tau])
        elif p.t0 > tend:
            print("WARNING: tend is smaller than the end of a pulse")
        elif p.t0 < t0:
            print("WARNING: t0 is greater than the start of a pulse")
    windows = _merge_intervals(sorted(windows))
    if windows[0][0] < t0:
        print("WARNING: t0 is greater than the start of the first pulse")
    if windows[-1][1] > tend:
        print("WARNING: tend is smaller than the end of the last pulse")
    parts = [np.arange(t0, windows[0][0], dt_big)]
    if simple_exp and len(windows) == 1 and windows[0][1] != 0:
        lo, hi = windows[0]
        if gaussian_t:
            parts.append(get_gaussian_t(lo, hi, *pulses, dt_max=dt_big, dt_min=dt_small, interval_per_step=0.05))
        else:
            parts.append(np.arange(lo, hi, dt_small))
        parts.append(np.round(np.exp(np.arange(np.log(hi), np.log(tend), dt_exp))))
        parts.append(np.array([tend]))
        return np.concatenate(parts)
    for k, (lo, hi) in enumerate(windows):
        if k > 0:
            parts.append(np.arange(windows[k - 1][1], lo, dt_big))
        parts.append(np.arange(lo, hi, dt_small))
    parts.append(np.arange(windows[-1][1], tend, dt_big))
    if add_tend:
        parts.append(np.array([tend]))
    return np.concatenate(parts)


def round_to_dt(t, dt):
    """Snap to multiples of ``dt`` and drop duplicates, keeping order (reference ``tools.py:110-118``)."""
    snapped = np.round(np.asarray(t) / dt) * dt
    _, first = np.unique(snapped, return_index=True)
    return snapped[np.sort(first)]


def simple_t_gaussian(t0, texp, tend, dt_small=0.1, dt_big=1.0, *pulses, decimals=2, exp_part=True, add_tend=True):
    """Adaptive grid up to ``texp`` then exponential (or ``dt_big``) spacing (reference ``tools.py:120-135``)."""
    parts = [get_gaussian_t(t0, texp, *pulses, dt_max=dt_big, dt_min=dt_small, interval_per_step=0.05)]
    if exp_part:
        parts.append(np.exp(np.arange(np.log(texp - t0), np.log(tend - t0), dt_small)) + t0)
    else:
        parts.append(np.arange(texp, tend, dt_big))
    if add_tend:
        parts.append(np.array([tend]))
    return round_to_dt(np.concatenate(parts), dt_small)


# ------------------------------------------------------------------------------------ I/O
def export_csv(filename, *arg, precision=4, delimit=',', verbose=False):
    """Write columns with fixed ``%.{precision}f`` formatting (reference ``tools.py:137-165``; the
    ACE pulse-file format uses ``precision=8, delimit=' '``, ``general_system.py:69-70``)."""
    fmt = ["%.{}f".format(precision)] * len(arg)
    np.savetxt(filename, np.column_stack(arg), fmt=fmt, delimiter=delimit, newline="\n")
    if verbose:
        print("[i] csv saved to {}".format(filename))


# ------------------------------------------------------------------------------------ density matrices
def concurrence(rho):
    """Wootters concurrence of a two-qubit density matrix (reference ``tools.py:167-172``)."""
    flip = np.fliplr(np.diag([-1.0, 1.0, 1.0, -1.0]))
    ev = np.real(np.linalg.eigvals(rho @ flip @ np.conjugate(rho) @ flip))
    lam = np.sqrt(np.sort(ev))
    return max(0.0, lam[-1] - np.sum(lam[:-1]))


def serialize_dm(rho):
    return np.concatenate((np.real(rho).flatten(), np.imag(rho).flatten()))


def deserialize_dm(rho):
    dim = int(np.sqrt(len(rho) / 2))
    return rho[:dim ** 2].reshape(dim, dim) + 1j * rho[dim ** 2:].reshape(dim, dim)


def compose_dm(outputs, dim=2):
    """Assemble ``rho[t, j, k]`` from the upper-triangle outputs of :func:`output_ops_dm`
    (reference ``tools.py:188-201``, layout pinned by ``tests/test_output_ops.py:26-41``):
    output ``n`` (1-based, row 0 is time) fills ``rho[:, j, k]`` and its conjugate ``rho[:, k, j]``."""
    outputs = np.asarray(outputs)
    rho = np.zeros((len(outputs[0]), dim, dim), dtype=np.complex128)
    iu, ku = np.triu_indices(dim)
    for n, (j, k) in enumerate(zip(iu, ku), start=1):
        rho[:, j, k] = outputs[n]
        rho[:, k, j] = np.conjugate(outputs[n])
    return np.real(outputs[0]), rho


def generate_basis_states(dim):
    """Product-basis index tuples, rightmost factor fastest (reference ``tools.py:203-211``)."""
    return list(itertools.product(*[range(d) for d in dim]))


def basis_states(dim):
    if not isinstance(dim, list):
        dim = [dim]
    return ["|" + ",".join(str(i) for i in st) + "⟩" for st in generate_basis_states(dim)]


def matrix_element_operators(basis_states, dim, readable=False):
    """Operator strings ``|bra><ket|`` for every upper-triangle pair (reference ``tools.py:229-246``)."""
    ops = []
    for i, bra in enumerate(basis_states):
        for ket in basis_states[i:]:
            if readable:
                ops.append(" ⊗ ".join(f"|{b}⟩⟨{k}|_{d}" for b, k, d in zip(bra, ket, dim)))
            else:
                ops.append(" otimes ".join(f"|{b}><{k}|_{d}" for b, k, d in zip(bra, ket, dim)))
    return ops


def output_ops_dm(dim=[2, 2], readable=False):
    """Output operators whose expectation values give the full density matrix via
    :func:`compose_dm` (reference ``tools.py:248-258``; strings pinned by
    ``tests/test_output_ops.py:11-24,43-73``)."""
    if not isinstance(dim, (list, tuple)):
        dim = [dim]
    return matrix_element_operators(generate_basis_states(dim), dim, readable=readable)


def op_to_matrix(op):
    """Matrix of a single ``|n><m|_dim`` string, optionally parenthesised (reference ``tools.py:260-304``)."""
    dm = re.search(r"_(\d+)(?:\[.*\])?", op)
    if not dm:
        raise ValueError(f"Invalid dimension format in operator: {op}")
    dim = int(dm.group(1))
    m = re.match(r"[(]*\|(\d+)><(\d+)\|_[\d)]*", op)
    if not m:
        return None
    ket, bra = int(m.group(1)), int(m.group(2))
    if ket >= dim or bra >= dim:
        raise ValueError(f"Index out of bounds: ket_idx={ket}, bra_idx={bra}, dim={dim}")
    out = np.zeros((dim, dim), dtype=complex)
    out[ket, bra] = 1.0
    return out


# ------------------------------------------------------------------------------------ units
def nm_to_mev(lambda_light):
    return 1239.84198 / lambda_light  # hc in eV*nm -> meV


def mev_to_nm(energy_light):
    return 1239.84198 / energy_light


# ------------------------------------------------------------------------------------ dynamical-map algebra
def calc_tl_dynmap_pseudo(dm, times, debug=False):
    """Time-local maps from maps-since-t0: ``tl[0] = dm[0]`` and ``tl[i] = dm[i] pinv(dm[i-1])`` so that
    ``tl[i] rho(t_i) = rho(t_{i+1})`` (reference ``tools.py:446-484``; ``dm[i] = E_{t_{i+1}, t_0}``,
    pseudo-inverse with ``rcond=1e-12`` because decayed maps are singular).  The pseudo-inverses are
    computed in one batched SVD instead of the reference's per-step loop."""
    dm = np.asarray(dm, dtype=complex)
    n_tl = len(times) - 1
    tl = np.zeros((n_tl,) + dm.shape[1:], dtype=complex)
    if n_tl <= 0:
        return tl
    tl[0] = dm[0]
    if n_tl > 1:
        tl[1:] = dm[1:n_tl] @ np.linalg.pinv(dm[:n_tl - 1], rcond=1e-12)
    return tl


def extract_dms(dm, times, tau_c, t_MTOs):
    """Split time-local maps into the stationary map (first step beyond the memory time ``tau_c``) and
    the explicit blocks of ``tau_c`` length at the start and after every ``t_MTO`` (reference
    ``tools.py:486-545``)."""
    beyond = np.where(times > times[0] + tau_c)[0]
    n_c = int(beyond[0])
    starts = []
    for t_mto in t_MTOs:
        hit = np.where(times == t_mto)[0]
        if len(hit) == 0:
            print(f"Available times: {times}")
            print(f"Requested t_MTO: {t_mto}")
            raise ValueError(f"t_MTO {t_mto} not found in times array. Make sure that t_MTO is included in the times array.")
        starts.append(int(hit[0]))
    return dm[n_c], [dm[:n_c]] + [dm[s:s + n_c] for s in starts]


def check_tl_map_params(tl_map, rho0):
    n = int(rho0.shape[0])
    if rho0.shape[1] != n:
        raise ValueError("rho0 must be a {n}x{n} matrix")
    if tl_map.shape != (n ** 2, n ** 2):
        raise ValueError("tl_map must be a {}x{} matrix, is {}".format(n ** 2, n ** 2, np.shape(tl_map)))
    return n


def _chain(maps, v0, n_out):
    """``[v0, M_0 v0, M_1 M_0 v0, ...]`` for an iterable of maps (row-major vectorisation)."""
    out = np.zeros((n_out, len(v0)), dtype=complex)
    out[0] = v0
    for k, m in zip(range(1, n_out), maps):
        out[k] = m @ out[k - 1]
    return out


def use_tl_map(tl_map, times, rho0):
    """Stationary propagation ``rho_{k+1} = tl_map rho_k`` on ``times`` (reference ``tools.py:567-588``)."""
    n = check_tl_map_params(tl_map, rho0)
    return _chain(itertools.repeat(tl_map), rho0.reshape(n * n), len(times)).reshape(len(times), n, n)


def use_dm_block(dm, rho0):
    """One explicit map per step (reference ``tools.py:590-609``); returns ``len(dm) + 1`` states."""
    n = check_tl_map_params(dm[0], rho0)
    return _chain(dm, rho0.reshape(n * n), len(dm) + 1).reshape(len(dm) + 1, n, n)


def tl_pad_stationary_nsteps(tl_map, n_steps, rho):
    """Extend a state sequence to ``n_steps`` entries with the stationary map (reference ``:622-631``)."""
    n = rho.shape[-1]
    have = len(rho)
    out = np.zeros((n_steps, n * n), dtype=complex)
    out[:have] = rho.reshape(have, n * n)
    for i in range(have, n_steps):
        out[i] = tl_map @ out[i - 1]
    return out.reshape(n_steps, n, n)


def tl_pad_stationary(tl_map, times, rho):
    return tl_pad_stationary_nsteps(tl_map, len(times), rho)


def use_tl_map_mto(tl_map, dm_1, dm_2, times, rho0, t_MTO, debug=False):
    """Piecewise propagation with one multi-time operator (reference ``tools.py:633-675``): explicit
    maps ``dm_1`` for the first memory time, the stationary map up to ``t_MTO``, explicit maps
    ``dm_2`` (which contain the operator) for one memory time after it, stationary again."""
    n = check_tl_map_params(tl_map, rho0)
    times = np.round(times, 5)
    i_mto = int(np.where(times >= t_MTO)[0][0])
    n1 = min(i_mto, len(dm_1))
    if i_mto < len(dm_1):
        print("caution: t_MTO is smaller than tau_c")
    if debug:
        print("info on piecewise application: ", i_mto, times[i_mto], len(dm_1), len(dm_2))
    schedule = itertools.chain(dm_1[:n1], itertools.repeat(tl_map, i_mto - n1), dm_2, itertools.repeat(tl_map))
    return _chain(schedule, rho0.reshape(n * n), len(times)).reshape(len(times), n, n)


def read_calibration_file(calibration_file):
    """Quantum-dot parameters from an experimental calibration file (INI sections ``EMISSION``, ``SPLITTING``,
    ``LIFETIMES``, ``G_FACTORS``; reference ``tools.py:307-344``).  Returns ``(E_X, E_Y, E_S, E_F, E_B, gamma_e,
    gamma_b, gamma_d, g_ex, g_hx, g_ez, g_hz)``: exciton energies relative to the mean bright exciton, the binding
    energy negative, rates in 1/ps."""
    import configparser
    config = configparser.ConfigParser()
    config.read(calibration_file)
    to_mev = lambda wavelength_nm: 1239.8 * 1e3 / wavelength_nm
    exciton = to_mev(float(config['EMISSION']['exciton_wavelength']))
    biexciton = to_mev(float(config['EMISSION']['biexciton_wavelength']))
    dark = to_mev(float(config['EMISSION']['dark_wavelength']))
    fss_bright = float(config['SPLITTING']['fss_bright']) * 1e-3
    fss_dark = float(config['SPLITTING']['fss_dark']) * 1e-3
    gamma_e = 1 / float(config['LIFETIMES']['exciton'])
    gamma_b = 1 / (float(config['LIFETIMES']['biexciton']) * 2)
    g = [float(config['G_FACTORS'][k]) for k in ('g_ex', 'g_hx', 'g_ez', 'g_hz')]
    dark_energy = dark - exciton
    return (fss_bright / 2, -fss_bright / 2, dark_energy + fss_dark / 2, dark_energy - fss_dark / 2,
            -(exciton - biexciton), gamma_e, gamma_b, 0, *g)


# ---------------------------------------------------------------------------------------------- post-processing helpers
def ghz_to_mev(ghz):
    """Frequency in GHz -> photon energy in meV, ``E = h f`` (reference ``tools.py:746-757``)."""
    return ghz * (2 * np.pi * 0.6582119514) * 1e-3


def mev_to_ghz(mev):
    """Photon energy in meV -> frequency in GHz (reference ``tools.py:759-770``)."""
    return mev / ((2 * np.pi * 0.6582119514) * 1e-3)


def rotate_basis(rho, U_rot):
    """``U rho U^dagger``: density matrix in a rotated basis, e.g. the field-mixed eigenbasis of the six-level
    system (reference ``tools.py:375-398``)."""
    return U_rot @ rho @ np.conj(U_rot).T


def resample(x, y, z, s_x, s_y):
    """Every ``s_x``-th / ``s_y``-th sample of a map ``z[y, x]`` and of its axes (reference ``tools.py:352-373``)."""
    ix = (np.arange(int(len(x) / s_x)) * s_x).astype(int)
    iy = (np.arange(int(len(y) / s_y)) * s_y).astype(int)
    return np.asarray(x, dtype=float)[ix], np.asarray(y, dtype=float)[iy], np.asarray(z, dtype=float)[np.ix_(iy, ix)]


def with_filename(func):
    """Decorator of :func:`get_sparse_range`: with ``filename`` the result comes back as ``(range, filename +
    "_sparse" | "_inverse")`` (reference ``tools.py:772-788``)."""
    import functools

    @functools.wraps(func)
    def wrapper(start=0.1, stop=12, num=101, nth=10, get_inverse=False, round_to=8, filename=None):
        result = func(start, stop, num, nth, get_inverse, round_to)
        if filename is not None:
            return result, filename + ("_inverse" if get_inverse else "_sparse")
        return result
    return wrapper


@with_filename
def get_sparse_range(start=0.1, stop=12, num=101, nth=10, get_inverse=False, round_to=8):
    """Every ``nth`` point of ``linspace(start, stop, num)``, or with ``get_inverse`` all the others, sorted and
    rounded (reference ``tools.py:790-802``): coarse and fill-in halves of a parameter sweep."""
    full = np.linspace(start, stop, num)
    if get_inverse:
        keep = np.ones(num, dtype=bool)
        keep[::nth] = False
        return np.round(np.sort(full[keep]), round_to)
    return full[::nth]


def get_union(arr_x1, arr_x2, arr_z1, arr_z2, axis_z=None):
    """Merge two sweeps: sorted union of the abscissae and the values reordered with it, duplicates taken from the first
    sweep (reference ``tools.py:804-831``)."""
    z1, z2 = np.asarray(arr_z1), np.asarray(arr_z2)
    if z1.ndim == 1:
        z1 = z1.reshape(len(arr_x1), 1)
    if z2.ndim == 1:
        z2 = z2.reshape(len(arr_x2), 1)
    if axis_z is None:
        if z1.shape[0] == z1.shape[1]:
            raise ValueError("Cannot determine axis for z arrays.")
        if z1.shape[0] == len(arr_x1) and z2.shape[0] == len(arr_x2):
            axis_z = 0
        elif z1.shape[1] == len(arr_x1) and z2.shape[1] == len(arr_x2):
            axis_z = 1
        else:
            raise ValueError("Cannot determine axis for z arrays.")
    x, idx = np.unique(np.concatenate((arr_x1, arr_x2)), return_index=True)
    return x, np.take(np.concatenate((z1, z2), axis=axis_z), idx, axis=axis_z)


def check_tlmap_frobenius(tl_map, times, filename="dynmap_tl_frobenius", xlim=25, check_against_i=None):
    """Frobenius distances between adjacent time-local maps (or to map ``check_against_i``) and the norms of the maps:
    how fast they become stationary (reference ``tools.py:677-743`` plots them; here the numbers are returned and the
    plots are written only where matplotlib is available)."""
    tl_map = np.asarray(tl_map)
    times = np.asarray(times, dtype=float)
    n = len(times) - 3
    if check_against_i is not None:
        diffs = np.array([np.linalg.norm(tl_map[i] - tl_map[check_against_i]) for i in range(n)])
    else:
        diffs = np.array([np.linalg.norm(tl_map[i] - tl_map[i + 1]) for i in range(n)])
    norms = np.array([np.linalg.norm(m) for m in tl_map])
    try:
        import matplotlib.pyplot as plt
    except ImportError:
        return diffs, norms
    ix = np.where((times - times[0] > 0) & (times - times[0] < xlim))[0]
    ix = ix[ix - 1 < len(diffs)]
    for ydata, suffix, title in ((diffs[ix - 1], "_diff", "difference of adjacent dynamical maps"),
                                 (norms[np.minimum(ix, len(norms) - 1)], "_norm", "norm of dynamical maps")):
        plt.clf()
        plt.xlabel("Time")
        plt.ylabel("Norm")
        plt.title(title)
        plt.plot(times[ix] - times[0], ydata)
        plt.yscale("log")
        plt.xlim(0, xlim)
        plt.savefig(filename + suffix + ".png")
    plt.clf()
    return diffs, norms
