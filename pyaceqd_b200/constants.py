"""Unit contract of the engine (ps, meV, K, nm).

Mirrors the three names the reference exposes in ``pyaceqd/constants.py:1-3``.
``pybind_path`` is kept only so that user scripts that set it keep working; the
engine never imports ACEutils.
"""
hbar = 0.6582119569  # meV*ps  (reference: pyaceqd/constants.py:1)
kB = 0.0861733326  # meV/K
pybind_path = ""
temp_dir = ""
