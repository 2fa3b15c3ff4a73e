"""TEST INFRASTRUCTURE ONLY — ctypes wrapper of the plain-C oracle (oracle/c/hamming_oracle.c).

Builds ``oracle/_build/liboracle.so`` on demand with gcc (``make -C oracle``) and fans query ranges out
over host threads (ctypes releases the GIL).  Used by tests at sizes where the numpy oracle is too slow
and by ``bench.py`` as an optional fast CPU baseline.  Never imported by the product package.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "c", "hamming_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.oracle_pack_codes_f32.restype = ctypes.c_int64
        _lib.oracle_pack_labels_i64.restype = ctypes.c_int64
        _lib.oracle_map_tindex.restype = ctypes.c_int
        _lib.oracle_hamming_u16.restype = None
        _lib.oracle_topk.restype = None
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _threads() -> int:
    return max(1, min(os.cpu_count() or 1, 64))


def _fanout(nq: int, fn, threads: Optional[int] = None):
    t = threads or _threads()
    step = max(1, -(-nq // (t * 4)))
    spans = [(lo, min(lo + step, nq)) for lo in range(0, nq, step)]
    if t == 1 or len(spans) == 1:
        return [fn(lo, hi) for lo, hi in spans]
    with ThreadPoolExecutor(max_workers=t) as ex:
        return list(ex.map(lambda s: fn(*s), spans))


def pack_codes(codes: np.ndarray) -> Tuple[np.ndarray, int]:
    codes = np.ascontiguousarray(codes, dtype=np.float32)
    n, nbits = codes.shape
    out = np.empty((n, (nbits + 31) // 32), dtype=np.uint32)
    bad = lib().oracle_pack_codes_f32(_p(codes), ctypes.c_int64(n), ctypes.c_int(nbits), _p(out))
    return out, int(bad)


def pack_labels(labels: np.ndarray) -> Tuple[np.ndarray, int]:
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    n, ncls = labels.shape
    out = np.empty((n, 4), dtype=np.uint32)
    bad = lib().oracle_pack_labels_i64(_p(labels), ctypes.c_int64(n), ctypes.c_int(ncls), _p(out))
    return out, int(bad)


def hamming_u16(qp: np.ndarray, gp: np.ndarray, threads: Optional[int] = None) -> np.ndarray:
    qp = np.ascontiguousarray(qp, dtype=np.uint32)
    gp = np.ascontiguousarray(gp, dtype=np.uint32)
    nq, n, w = qp.shape[0], gp.shape[0], gp.shape[1]
    out = np.empty((nq, n), dtype=np.uint16)

    def run(lo, hi):
        lib().oracle_hamming_u16(_p(qp[lo:hi]), ctypes.c_int64(hi - lo), _p(gp), ctypes.c_int64(n),
                                 ctypes.c_int(w), _p(out[lo:hi]))

    _fanout(nq, run, threads)
    return out


def map_tindex(qp, qlp, gp, glp, nbits: int, k: Optional[int] = None, cap: Optional[int] = None,
               want_hist: bool = False, threads: Optional[int] = None):
    """-> (tindex [Q,cap] int32 (0 beyond totals), totals [Q] int32, tsums [Q] int32[, hist_all, hist_rel])."""
    qp = np.ascontiguousarray(qp, dtype=np.uint32)
    gp = np.ascontiguousarray(gp, dtype=np.uint32)
    qlp = np.ascontiguousarray(qlp, dtype=np.uint32)
    glp = np.ascontiguousarray(glp, dtype=np.uint32)
    nq, n = qp.shape[0], gp.shape[0]
    if k is None:
        k = n
    if cap is None:
        cap = min(int(k), n)
    cap = max(int(cap), 1)
    tindex = np.zeros((nq, cap), dtype=np.int32)
    totals = np.zeros(nq, dtype=np.int32)
    tsums = np.zeros(nq, dtype=np.int32)
    ha = np.zeros((nq, nbits + 1), dtype=np.int32) if want_hist else None
    hr = np.zeros((nq, nbits + 1), dtype=np.int32) if want_hist else None

    def run(lo, hi):
        return lib().oracle_map_tindex(
            _p(qp[lo:hi]), _p(qlp[lo:hi]), ctypes.c_int64(hi - lo), _p(gp), _p(glp),
            ctypes.c_int64(n), ctypes.c_int(nbits), ctypes.c_int64(k), _p(tindex[lo:hi]),
            ctypes.c_int64(cap), _p(totals[lo:hi]), _p(tsums[lo:hi]),
            _p(ha[lo:hi]) if want_hist else None, _p(hr[lo:hi]) if want_hist else None)

    rcs = _fanout(nq, run, threads)
    if any(rc == -2 for rc in rcs):
        raise ValueError("too many bits for the C oracle")
    if want_hist:
        return tindex, totals, tsums, ha, hr
    return tindex, totals, tsums


def topk(qp, gp, nbits: int, k: int, threads: Optional[int] = None):
    qp = np.ascontiguousarray(qp, dtype=np.uint32)
    gp = np.ascontiguousarray(gp, dtype=np.uint32)
    nq, n = qp.shape[0], gp.shape[0]
    od = np.empty((nq, k), dtype=np.int32)
    oi = np.empty((nq, k), dtype=np.int32)

    def run(lo, hi):
        lib().oracle_topk(_p(qp[lo:hi]), ctypes.c_int64(hi - lo), _p(gp), ctypes.c_int64(n),
                          ctypes.c_int(nbits), ctypes.c_int64(k), _p(od[lo:hi]), _p(oi[lo:hi]))

    _fanout(nq, run, threads)
    return od, oi
