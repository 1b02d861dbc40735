"""ctypes wrapper over oracle/liboracle_vadd.so (built by oracle/Makefile)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle_vadd.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "vadd_oracle.c")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return LIB_PATH


if not os.path.exists(LIB_PATH):
    build()
_lib = C.CDLL(LIB_PATH)
_P, _SZ, _U64, _I = C.c_void_p, C.c_size_t, C.c_uint64, C.c_int
for _n, _r, _a in [
    ("oracle_vadd_f32", None, [_P, _P, _P, _SZ]),
    ("oracle_softfloat_add_f32", C.c_uint32, [C.c_uint32, C.c_uint32]),
    ("oracle_softfloat_vadd_f32", None, [_P, _P, _P, _SZ]),
    ("oracle_fill_rand_f32", None, [_P, _P, _SZ]),
    ("oracle_fill_ctr_f32", None, [_P, _SZ, _U64, _U64]),
    ("oracle_verify_sample_tolerance", C.c_longlong, [_P, _P, _P, _SZ]),
    ("oracle_first_mismatch_f32", C.c_longlong, [_P, _P, _SZ]),
    ("oracle_fnv1a64", _U64, [_P, _SZ]),
    ("oracle_bits_digest_u32", None, [_P, _SZ, _P]),
    ("oracle_vadd_digest_f32", None, [_P, _P, _SZ, _P]),
    ("oracle_num_cpus", _I, []),
    ("oracle_cpu_quota", C.c_double, []),
    ("oracle_vadd_f32_mt", None, [_P, _P, _P, _SZ, _I]),
    ("oracle_fill_ctr_pair_mt", None, [_P, _P, _P, _SZ, _U64, _U64, _U64, _I]),
    ("oracle_vadd_digest_f32_mt", None, [_P, _P, _SZ, _I, _P]),
    ("oracle_time_vadd_mt", _I, [_SZ, _I, _I, _I, _P]),
    ("oracle_time_vadd_mt_ex", _I, [_SZ, _I, _I, _I, _I, _P]),
    ("oracle_vadd_f32_mt_nt", None, [_P, _P, _P, _SZ, _I]),
    ("oracle_nt_width", _I, []),
    ("oracle_stream", _I, [_I, _I, _P, _P, _P, _SZ, C.c_double]),
    ("oracle_half_to_float", None, [_P, _P, _SZ]),
    ("oracle_float_to_half", None, [_P, _P, _SZ]),
    ("oracle_bf16_to_float", None, [_P, _P, _SZ]),
    ("oracle_float_to_bf16", None, [_P, _P, _SZ]),
]:
    _f = getattr(_lib, _n)
    _f.restype, _f.argtypes = _r, _a

SEED_A, SEED_B = 0x0A, 0x0B


def _f32(x: np.ndarray) -> np.ndarray:
    assert x.dtype == np.float32 and x.flags.c_contiguous
    return x


def vadd(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """C[i] = A[i] + B[i], scalar single-thread C loop (the reference kernel's arithmetic)."""
    c = np.empty_like(_f32(a))
    _lib.oracle_vadd_f32(a.ctypes.data, _f32(b).ctypes.data, c.ctypes.data, a.size)
    return c


def vadd_mt(a: np.ndarray, b: np.ndarray, threads: int = 0) -> np.ndarray:
    c = np.empty_like(_f32(a))
    _lib.oracle_vadd_f32_mt(a.ctypes.data, _f32(b).ctypes.data, c.ctypes.data, a.size, threads or num_cpus())
    return c


def vadd_mt_nt(a: np.ndarray, b: np.ndarray, threads: int = 0) -> np.ndarray:
    """The non-temporal-store timing variant (SSE/AVX/AVX-512 vaddps + movntps): same bits as vadd."""
    c = np.empty_like(_f32(a))
    _lib.oracle_vadd_f32_mt_nt(a.ctypes.data, _f32(b).ctypes.data, c.ctypes.data, a.size, threads or num_cpus())
    return c


def nt_width() -> int:
    """Vector width (bits) of the non-temporal variant on this host: 512, 256, 128 (0: not x86-64)."""
    return int(_lib.oracle_nt_width())


def softfloat_vadd_bits(ua: np.ndarray, ub: np.ndarray) -> np.ndarray:
    assert ua.dtype == np.uint32 and ub.dtype == np.uint32
    uc = np.empty_like(ua)
    _lib.oracle_softfloat_vadd_f32(np.ascontiguousarray(ua).ctypes.data, np.ascontiguousarray(ub).ctypes.data,
                                   uc.ctypes.data, ua.size)
    return uc


def softfloat_add_bits(ua: int, ub: int) -> int:
    return int(_lib.oracle_softfloat_add_f32(ua, ub))


def fill_rand(n: int) -> tuple[np.ndarray, np.ndarray]:
    a, b = np.empty(n, np.float32), np.empty(n, np.float32)
    _lib.oracle_fill_rand_f32(a.ctypes.data, b.ctypes.data, n)
    return a, b


def fill_ctr(n: int, seed: int, first: int = 0) -> np.ndarray:
    x = np.empty(n, np.float32)
    _lib.oracle_fill_ctr_f32(x.ctypes.data, n, seed, first)
    return x


def verify_sample_tolerance(a, b, c) -> int:
    return int(_lib.oracle_verify_sample_tolerance(_f32(a).ctypes.data, _f32(b).ctypes.data, _f32(c).ctypes.data, a.size))


def first_mismatch(x: np.ndarray, y: np.ndarray) -> int:
    assert x.size == y.size
    return int(_lib.oracle_first_mismatch_f32(_f32(x).ctypes.data, _f32(y).ctypes.data, x.size))


def fnv1a64(x: np.ndarray) -> int:
    x = np.ascontiguousarray(x)
    return int(_lib.oracle_fnv1a64(x.ctypes.data, x.nbytes))


def bits_digest(x: np.ndarray) -> tuple[int, int]:
    out = (C.c_uint64 * 2)()
    _lib.oracle_bits_digest_u32(_f32(x).ctypes.data, x.size, out)
    return int(out[0]), int(out[1])


def vadd_digest(a: np.ndarray, b: np.ndarray, threads: int = 1) -> tuple[int, int]:
    out = (C.c_uint64 * 2)()
    if threads > 1:
        _lib.oracle_vadd_digest_f32_mt(_f32(a).ctypes.data, _f32(b).ctypes.data, a.size, threads, out)
    else:
        _lib.oracle_vadd_digest_f32(_f32(a).ctypes.data, _f32(b).ctypes.data, a.size, out)
    return int(out[0]), int(out[1])


def ctr_vadd_digest(n: int, first: int = 0, threads: int = 0, block: int = 1 << 24) -> tuple[int, int]:
    """Digest of ctr(A)+ctr(B) over [first, first+n) without holding the vectors: the
    full-size (2^28 / 2^30) checker."""
    threads = threads or num_cpus()
    s, x = 0, 0
    a = np.empty(min(block, max(n, 1)), np.float32)
    b = np.empty_like(a)
    done = 0
    while done < n:
        m = min(block, n - done)
        _lib.oracle_fill_ctr_pair_mt(a.ctypes.data, b.ctypes.data, None, m, SEED_A, SEED_B, first + done, threads)
        out = (C.c_uint64 * 2)()
        _lib.oracle_vadd_digest_f32_mt(a.ctypes.data, b.ctypes.data, m, threads, out)
        s = (s + int(out[0])) & 0xFFFFFFFFFFFFFFFF
        x ^= int(out[1])
        done += m
    return s, x


def num_cpus() -> int:
    return int(_lib.oracle_num_cpus())


def cpu_quota() -> float:
    """Container CPU quota in CPUs (0 = unlimited)."""
    return float(_lib.oracle_cpu_quota())


def best_thread_count(n: int = 1 << 26, nt_stores: bool = False) -> tuple[int, dict[int, float]]:
    """All the host threads the add can USE: tries the affinity count, multiples of the cgroup
    quota and a few fixed counts on a short sample (median of 3 passes each) and returns the
    fastest (threads, {threads: elem/s})."""
    ncpu, quota = num_cpus(), cpu_quota()
    cands = {ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 8), min(ncpu, 16), min(ncpu, 32)}
    if quota > 0:
        cands |= {max(1, min(ncpu, int(round(k * quota)))) for k in (1, 1.5, 2, 3, 4)}
    rates = {}
    for t in sorted(cands):
        secs = sorted(time_vadd_mt(n, t, 1, 3, nt_stores))
        rates[t] = n / secs[1]
    return max(rates, key=rates.get), rates


def time_vadd_mt(n: int, threads: int = 0, warmup: int = 1, reps: int = 5, nt_stores: bool = False) -> list[float]:
    """Per-pass seconds of the all-cores add over n elements (buffers first-touched by the workers).
    nt_stores: non-temporal stores (no write-allocate read of C) instead of regular ones."""
    secs = (C.c_double * reps)()
    rc = _lib.oracle_time_vadd_mt_ex(n, threads or num_cpus(), warmup, reps, 1 if nt_stores else 0, secs)
    if rc != 0:
        raise MemoryError("oracle_time_vadd_mt: allocation failed")
    return list(secs)


def best_cpu_config(n: int = 1 << 26) -> dict:
    """The fastest way this host adds two vectors: regular vs non-temporal stores, each at its
    best thread count.  {"stores", "threads", "rate", "tried": {...}}"""
    out = {"tried": {}}
    for nt in (False, True):
        t, rates = best_thread_count(n, nt)
        name = f"non-temporal ({nt_width()}-bit vmovntps)" if nt else "regular (write-allocate)"
        out["tried"][name] = {"threads": t, "elements_per_s": rates[t], "all": {str(k): v for k, v in sorted(rates.items())}}
        if "rate" not in out or rates[t] > out["rate"]:
            out.update(stores=name, threads=t, rate=rates[t], nt=nt)
    return out


# ---------------------------------------------------------------- f4: STREAM-style ops
OPS = {"copy": 0, "scale": 1, "add": 2, "triad": 3}
DTYPES = {"f32": (0, np.float32), "f64": (1, np.float64), "f16": (2, np.uint16), "bf16": (3, np.uint16)}


def stream(op: str, dtype: str, a: np.ndarray, b: np.ndarray | None, scalar: float = 0.0) -> np.ndarray:
    """Oracle of b200va_stream. f16/bf16 arrays are uint16 bit patterns."""
    code, npdt = DTYPES[dtype]
    a = np.ascontiguousarray(a)
    assert a.dtype == npdt, (a.dtype, npdt)
    c = np.empty_like(a)
    pb = np.ascontiguousarray(b).ctypes.data if b is not None else None
    rc = _lib.oracle_stream(OPS[op], code, a.ctypes.data, pb, c.ctypes.data, a.size, float(scalar))
    assert rc == 0
    return c


def half_to_float(h: np.ndarray) -> np.ndarray:
    f = np.empty(h.size, np.float32)
    _lib.oracle_half_to_float(np.ascontiguousarray(h).ctypes.data, f.ctypes.data, h.size)
    return f


def float_to_half(f: np.ndarray) -> np.ndarray:
    h = np.empty(f.size, np.uint16)
    _lib.oracle_float_to_half(np.ascontiguousarray(f, np.float32).ctypes.data, h.ctypes.data, f.size)
    return h


def bf16_to_float(h: np.ndarray) -> np.ndarray:
    f = np.empty(h.size, np.float32)
    _lib.oracle_bf16_to_float(np.ascontiguousarray(h).ctypes.data, f.ctypes.data, h.size)
    return f


def float_to_bf16(f: np.ndarray) -> np.ndarray:
    h = np.empty(f.size, np.uint16)
    _lib.oracle_float_to_bf16(np.ascontiguousarray(f, np.float32).ctypes.data, h.ctypes.data, f.size)
    return h


def first_mismatch_bits(x: np.ndarray, y: np.ndarray, dtype: str) -> int:
    """First index where the bit patterns differ, NaNs (of the given dtype) matching as a class; -1 if none."""
    x, y = np.ascontiguousarray(x), np.ascontiguousarray(y)
    if dtype == "f32":
        ux, uy = x.view(np.uint32), y.view(np.uint32)
        nx, ny = (ux & 0x7FFFFFFF) > 0x7F800000, (uy & 0x7FFFFFFF) > 0x7F800000
    elif dtype == "f64":
        ux, uy = x.view(np.uint64), y.view(np.uint64)
        m, inf = np.uint64(0x7FFFFFFFFFFFFFFF), np.uint64(0x7FF0000000000000)
        nx, ny = (ux & m) > inf, (uy & m) > inf
    elif dtype == "f16":
        ux, uy = x.view(np.uint16), y.view(np.uint16)
        nx, ny = (ux & 0x7FFF) > 0x7C00, (uy & 0x7FFF) > 0x7C00
    else:
        ux, uy = x.view(np.uint16), y.view(np.uint16)
        nx, ny = (ux & 0x7FFF) > 0x7F80, (uy & 0x7FFF) > 0x7F80
    bad = (ux != uy) & ~(nx & ny)
    idx = np.nonzero(bad)[0]
    return int(idx[0]) if idx.size else -1
