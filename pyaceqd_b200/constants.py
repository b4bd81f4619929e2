"""Unit contract of the engine (ps, meV, K, nm).

Mirrors the three names the reference exposes in ``pyaceqd/constants.py:1-3``.
``pybind_path`` is kept only so that user scripts that set it keep working; the
engine never imports ACEutils.
"""
hbar = 0.6582119569  # meV*ps  (reference: pyaceqd/constants.py:1)
kB = 0.0861733326  # meV/K
pybind_path = ""
temp_dir = ""
# "_right" multi-time operators: False = rho -> rho A with A as given (the reference author's own
# restatement, two_time/propagate_tau.f90:91-92); True = rho -> rho A^T (the transposition the
# legacy comment at four_level_system/dark_model.py:267-268 attributes to ACE).  SURVEY App. C.3.
mto_right_transposed = False
# DynamicalMap.E layout: the reference's consumers treat dm[0] as E_{t1,t0} (tools.py:470-479), so the
# identity at t0 is not returned by default.
dynmap_includes_t0 = False
# two_time/propagate_tau.f90:492,524 return to the maps of the period start when ``j + j_start == n_tb + 1``, one step
# BEFORE the period boundary (the phonon-free routine, :286, resets at the boundary).  False keeps the reference's
# numbers; True resets at the boundary, which is what the direct multi-time-operator sweeps agree with.
phonon_block_aligned = False
