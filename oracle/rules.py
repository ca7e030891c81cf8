"""ORACLE -- TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED (see oracle/README.md).

Rulebook builders of the CPU oracle: ctypes front-end of oracle/rules.cpp plus an
independent pure-Python (dict) restatement used to cross-check the C++ one on small cases.
Anchors: /root/reference/uresnet/models/uresnet_sparse.py:20-22 (the scn call sites);
algorithm per SURVEY.md Appendix A items 1-6.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_rules.so")
_lib = None


def build(force=False):
    """g++ the C++ restatement into oracle/liboracle_rules.so (git-ignored, travels with gpurun)."""
    src = os.path.join(_HERE, "rules.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", _SO, src])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i64p = ctypes.POINTER(ctypes.c_int64)
        i32p = ctypes.POINTER(ctypes.c_int32)
        L.orc_input_rows.restype = ctypes.c_int64
        L.orc_input_rows.argtypes = [i64p, ctypes.c_int64, ctypes.c_int, i32p, i64p]
        L.orc_subm_rules.restype = ctypes.c_int64
        L.orc_subm_rules.argtypes = [i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, i64p, i32p, i32p]
        L.orc_down_rules.restype = ctypes.c_int64
        L.orc_down_rules.argtypes = [i64p, ctypes.c_int64, ctypes.c_int, ctypes.c_int, i32p, i32p, i64p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(ctypes.POINTER(t))


def input_rows(coords):
    """coords int64 [n, D+1] -> (in2row int32 [n], active int64 [nActive, D+1])."""
    coords = np.ascontiguousarray(coords, np.int64)
    n, D1 = coords.shape
    in2row = np.empty(n, np.int32)
    active = np.empty((max(n, 1), D1), np.int64)
    m = lib().orc_input_rows(_p(coords, ctypes.c_int64), n, D1 - 1, _p(in2row, ctypes.c_int32),
                             _p(active, ctypes.c_int64))
    return in2row, active[:m].copy()


def subm_rules(active, k=3):
    """active int64 [n, D+1] -> list over K=k^D offsets of (in_rows, out_rows) int32 arrays."""
    active = np.ascontiguousarray(active, np.int64)
    n, D1 = active.shape
    K = k ** (D1 - 1)
    off = np.empty(K + 1, np.int64)
    rin = np.empty(max(n * K, 1), np.int32)
    rout = np.empty(max(n * K, 1), np.int32)
    lib().orc_subm_rules(_p(active, ctypes.c_int64), n, D1 - 1, k, _p(off, ctypes.c_int64),
                         _p(rin, ctypes.c_int32), _p(rout, ctypes.c_int32))
    return [(rin[off[f]:off[f + 1]].copy(), rout[off[f]:off[f + 1]].copy()) for f in range(K)]


def down_rules(active, s=2):
    """active int64 [n, D+1] -> (parent int32 [n], kofs int32 [n], coarse int64 [m, D+1])."""
    active = np.ascontiguousarray(active, np.int64)
    n, D1 = active.shape
    parent = np.empty(n, np.int32)
    kofs = np.empty(n, np.int32)
    coarse = np.empty((max(n, 1), D1), np.int64)
    m = lib().orc_down_rules(_p(active, ctypes.c_int64), n, D1 - 1, s, _p(parent, ctypes.c_int32),
                             _p(kofs, ctypes.c_int32), _p(coarse, ctypes.c_int64))
    return parent, kofs, coarse[:m].copy()


# ---------------------------------------------------------------------------------------
# Pure-Python restatement (dict per sample), small cases only: cross-checks the C++ above.
# ---------------------------------------------------------------------------------------
def input_rows_py(coords):
    coords = np.asarray(coords, np.int64)
    seen, in2row, active = {}, [], []
    for c in coords:
        key = tuple(int(v) for v in c)
        if key not in seen:
            seen[key] = len(seen)
            active.append(key)
        in2row.append(seen[key])
    D1 = coords.shape[1]
    return np.array(in2row, np.int32), np.array(active, np.int64).reshape(-1, D1)


def subm_rules_py(active, k=3):
    active = np.asarray(active, np.int64)
    D = active.shape[1] - 1
    grid = {tuple(int(v) for v in c): r for r, c in enumerate(active)}
    K = k ** D
    rules = [([], []) for _ in range(K)]
    for o, c in enumerate(active):
        for f in range(K):
            rem, p = f, [0] * D
            for d in range(D - 1, -1, -1):
                p[d] = int(c[d]) + rem % k - k // 2
                rem //= k
            i = grid.get(tuple(p) + (int(c[D]),))
            if i is not None:
                rules[f][0].append(i)
                rules[f][1].append(o)
    return [(np.array(a, np.int32), np.array(b, np.int32)) for a, b in rules]


def down_rules_py(active, s=2):
    active = np.asarray(active, np.int64)
    D = active.shape[1] - 1
    grid, parent, kofs, coarse = {}, [], [], []
    for c in active:
        p = tuple(int(v) // s for v in c[:D]) + (int(c[D]),)
        f = 0
        for d in range(D):
            f = f * s + int(c[d]) % s
        if p not in grid:
            grid[p] = len(grid)
            coarse.append(p)
        parent.append(grid[p])
        kofs.append(f)
    return (np.array(parent, np.int32), np.array(kofs, np.int32),
            np.array(coarse, np.int64).reshape(-1, D + 1))
