"""The C-ABI library loads on a CPU-only box and exports every symbol the header declares."""
import ctypes
import os
import re

from bayesgm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "bgm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgm_[a-z0-9_]+)\s*\(", src)))


def test_library_loads_and_exports_header_symbols():
    lib = _lib.load()
    names = header_symbols()
    assert "bgm_causal_mh" in names and "bgm_causal_logpost" in names
    for name in names:
        assert hasattr(lib, name), "libbgm_b200.so does not export %s" % name
    assert set(names) == set(_lib.SYMBOLS), "ctypes table and header disagree"
    assert lib.bgm_version() >= 100


def test_mh_args_struct_layout_matches_header():
    # pointers 8 bytes, ints 4, natural alignment:
    # x,y,v | ldv,n | vproj,r0 | ldvproj(+pad) | sched | z,lp | 4 ints | q_sd,eps,u | seed,row_offset | 4 ptrs |
    # prior, ldprior(+pad)
    assert ctypes.sizeof(_lib.MhArgs) == 24 + 8 + 16 + 8 + 8 + 16 + 16 + 24 + 16 + 32 + 16
    assert _lib.MhArgs.sched_dev.offset == 56 and _lib.MhArgs.q_sd_dev.offset == 96
    assert _lib.MhArgs.prior_dev.offset == 168


def test_hmc_args_struct_layout_matches_header():
    # x,ldx,n | z,g,lp | 5 ints (+4 pad) | step,mom,logu | seed,row_offset | 5 pointers
    assert ctypes.sizeof(_lib.HmcArgs) == 16 + 24 + 24 + 24 + 16 + 40
    assert _lib.HmcArgs.step_dev.offset == 64 and _lib.HmcArgs.seed.offset == 88


def test_bad_arguments_fail_loudly_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.bgm_causal_create(ctypes.byref(h), None, 10, 0, -1.0, -1.0, -1.0, None, None, None)
    assert rc < 0 and b"null" in lib.bgm_last_error()
    assert lib.bgm_causal_logpost(None, None, None, None, 0, None, 0, None, None, 0, None, None, None) < 0
    assert lib.bgm_hmc_create(ctypes.byref(h), None) < 0 and b"null" in lib.bgm_last_error()
    assert lib.bgm_hmc_run(None, None, None) < 0
    d = _lib.VarNetDesc(40, 10, 1, None, None, None, None, None)
    assert lib.bgm_hmc_create(ctypes.byref(h), ctypes.byref(d)) < 0


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "bayesgm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
