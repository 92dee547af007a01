"""In-tree build of libbgm_b200.so (sm_100a only) with nvcc.

`python -m bayesgm_b200._build` or `__graft_entry__.build()`.  Every `csrc/*.cu` is one
translation unit, compiled in parallel into `build/*.o` (only the stale ones) and linked into
the .so next to this file, so that it travels to the GPU box with the repo snapshot.
"""
import os
import re
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libbgm_b200.so")
HEADER = os.path.join(os.path.dirname(HERE), "include", "bgm_b200.h")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def _units():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu")]


def _deps(path, seen=None):
    """The .cu / .cuh file and everything it includes from csrc/ (transitively), plus the ABI header."""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    for inc in re.findall(r'#include\s+"([^"]+)"', open(path).read()):
        _deps(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)
    return seen


def _obj(unit):
    return os.path.join(OBJ, os.path.basename(unit)[:-3] + ".o")


def _unit_stale(unit):
    o = _obj(unit)
    if not os.path.exists(o):
        return True
    t = os.path.getmtime(o)
    return any(os.path.getmtime(d) > t for d in _deps(unit) | {HEADER})


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for u in _units() for d in _deps(u) | {HEADER})


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libbgm_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    todo = [u for u in _units() if force or _unit_stale(u)]

    def compile_one(unit):
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", unit, "-o", _obj(unit)]
        return unit, cmd, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        for unit, cmd, r in ex.map(compile_one, todo):
            if r.returncode != 0:
                raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
            if verbose:
                sys.stderr.write(r.stderr)
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + [_obj(u) for u in _units()]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (" ".join(cmd), r.stderr))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
