"""The C-ABI library loads on a CPU-only box, exports every symbol include/aceqd.h declares,
and refuses to compute without a GPU (no silent fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from pyaceqd_b200 import engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "aceqd.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(aceqd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = engine.load_library()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libaceqd.so does not export {n}"
    assert lib.aceqd_version().startswith(b"aceqd-b200")


def test_struct_layouts_match_binding():
    lib = engine.load_library()
    sizes = (ctypes.c_int32 * 4)()
    lib.aceqd_struct_sizes(sizes)
    assert tuple(sizes) == (engine.SEQ_DT.itemsize, engine.ENTRY_DT.itemsize, engine.TRAJ_DT.itemsize,
                            ctypes.sizeof(engine._Batch))
    assert engine.TRAJ_DT.fields["ovr_step"][0].shape == (engine.MAX_OVR,)


def test_tile_budget_is_pure_host_logic():
    lib = engine.load_library()
    assert lib.aceqd_max_tile(4, 128) == 16      # cfg2
    assert lib.aceqd_max_tile(16, 128) == 4      # cfg3
    assert lib.aceqd_max_tile(36, 128) == 2      # cfg4
    assert lib.aceqd_max_tile(25, 256) == 1      # cfg5
    assert lib.aceqd_max_tile_global_pt(25, 256) == 2    # ... two trajectories without a shared-memory PT ring
    assert lib.aceqd_max_tile_global_pt(16, 128) == 4 and lib.aceqd_max_tile_global_pt(36, 128) == 2
    assert lib.aceqd_max_tile(64, 256) == 0      # does not fit: reported, not truncated


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(engine.EngineError) as e:
        engine.Engine(0)
    assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "pyaceqd_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle", src, flags=re.M), f
                assert "liboracle_c" not in src, f
