"""Build libaceqd.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

Every ``.cu`` is compiled to its own object (in parallel, only when it or a header it includes is newer than the
object), then the objects are linked into ``csrc/libaceqd.so``."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(CSRC, "libaceqd.so")
SOURCES = ["api.cu", "expm.cu", "step_kernel.cu", "splitk_kernel.cu", "small_kernel.cu", "tlmap.cu", "peak.cu"]
# the on-device process-tensor builder is a library of its own: it links cuBLAS and cuSOLVER, the propagation
# library links nothing
PTBUILD_LIB = os.path.join(CSRC, "libaceqd_ptbuild.so")
PTBUILD_SRC = os.path.join(CSRC, "ptbuild.cu")
PTBUILD_HDR = os.path.join(CSRC, "..", "..", "include", "aceqd_ptbuild.h")
HEADERS = ["common.cuh", "kernel_common.cuh", os.path.join("..", "..", "include", "aceqd.h")]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-Xcompiler", "-fPIC"] + ARCH + ["-lineinfo", "-O3", "-std=c++17", "-diag-suppress", "177"]
LINK_LIBS: list = []


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _obj(src: str) -> str:
    return os.path.join(OBJ, os.path.splitext(src)[0] + ".o")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build() -> bool:
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    return _stale(LIB, [os.path.join(CSRC, s) for s in _sources()] + hdrs)


def build_library(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    todo = [s for s in _sources() if force or _stale(_obj(s), [os.path.join(CSRC, s)] + hdrs)]

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", _obj(src)]
        return src, subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(todo)))) as ex:
        for src, res in ex.map(compile_one, todo):
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError(f"nvcc failed compiling {src}")
            if verbose:
                sys.stderr.write(res.stderr)
    cmd = [nvcc, "-shared"] + ARCH + ["-o", LIB] + [_obj(s) for s in _sources()] + LINK_LIBS
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed linking libaceqd.so")
    return LIB


def build_ptbuild_library(force: bool = False) -> str:
    if not force and not _stale(PTBUILD_LIB, [PTBUILD_SRC, PTBUILD_HDR]):
        return PTBUILD_LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, "-shared"] + NVCC_FLAGS + ["-o", PTBUILD_LIB, PTBUILD_SRC, "-lcublas", "-lcusolver"]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building libaceqd_ptbuild.so")
    return PTBUILD_LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_ptbuild_library(force="--force" in sys.argv))
