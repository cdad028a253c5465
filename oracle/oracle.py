"""ctypes wrapper around oracle/vega_oracle.c — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product package (vega_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libvega_oracle.so")

U64, I64, F64 = 0, 1, 2
GROUP, SUM, MIN, MAX, COUNT = 0, 1, 2, 3, 4
OPS = {"group": GROUP, "sum": SUM, "min": MIN, "max": MAX, "count": COUNT}
DTYPES = {"u64": U64, "i64": I64, "f64": F64}
_NP = {U64: np.uint64, I64: np.int64, F64: np.float64}


def build(force=False):
    src = os.path.join(_HERE, "vega_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    l = ctypes.CDLL(_SO)
    vp, u64, u32, i32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int
    l.vo_metrohash64_1.restype = u64
    l.vo_metrohash64_1.argtypes = [ctypes.c_char_p, u64, u32]
    l.vo_hash_key.restype = u64
    l.vo_hash_key.argtypes = [u64, u32]
    l.vo_get_partition.restype = u32
    l.vo_get_partition.argtypes = [u64, u32, u32]
    l.vo_slice.restype = u64
    l.vo_slice.argtypes = [u64, u64, vp]
    l.vo_shuffle_run.restype = vp
    l.vo_shuffle_run.argtypes = [i32, i32, u32, vp, vp, u64, u64, u64, u64]
    l.vo_job_n_map.restype = u64
    l.vo_job_n_map.argtypes = [vp]
    l.vo_job_bucket_rows.restype = u64
    l.vo_job_bucket_rows.argtypes = [vp, u64, u64]
    l.vo_job_part_sizes.restype = None
    l.vo_job_part_sizes.argtypes = [vp, u64, vp, vp]
    l.vo_job_part_copy.restype = None
    l.vo_job_part_copy.argtypes = [vp, u64, vp, vp, vp, vp]
    l.vo_job_free.restype = None
    l.vo_job_free.argtypes = [vp]
    l.vo_join_run.restype = vp
    l.vo_join_run.argtypes = [u32, vp, vp, u64, u64, vp, vp, u64, u64, u64, u64]
    l.vo_join_part_size.restype = u64
    l.vo_join_part_size.argtypes = [vp, u64]
    l.vo_join_part_copy.restype = None
    l.vo_join_part_copy.argtypes = [vp, u64, vp, vp, vp]
    l.vo_join_free.restype = None
    l.vo_join_free.argtypes = [vp]
    l.vo_sort_by_key.restype = None
    l.vo_sort_by_key.argtypes = [i32, vp, vp, u64, u64, vp, vp, vp]
    l.vo_splitmix64.restype = u64
    l.vo_splitmix64.argtypes = [u64]
    l.vo_gen_uniform.restype = None
    l.vo_gen_uniform.argtypes = [vp, vp, u64, u64, u64, u64, u64]
    _lib = l
    return l


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def metrohash64_1(data: bytes, seed: int = 0) -> int:
    return lib().vo_metrohash64_1(data, len(data), seed)


def get_partition(key: int, n_reduce: int, key_width: int = 8) -> int:
    return lib().vo_get_partition(int(key) & 0xFFFFFFFFFFFFFFFF, key_width, n_reduce)


def slice_starts(n: int, num_slices: int) -> np.ndarray:
    """ParallelCollection::slice boundaries (parallel_collection_rdd.rs:116-145)."""
    buf = np.zeros(min(n, num_slices) + 2, dtype=np.uint64)
    k = lib().vo_slice(n, num_slices, _p(buf))
    return buf[: k + 1].copy()


def _as_u64(a):
    a = np.ascontiguousarray(a)
    assert a.dtype.itemsize == 8, a.dtype
    return a.view(np.uint64)


def shuffle(op, keys, vals, num_slices, n_reduce, vdtype="u64", key_width=8, threads=1):
    """Run combine_by_key end to end on the CPU.

    Returns a list (one entry per reduce partition) of dicts:
      reduce ops: {"keys": u64[nk], "combined": vdtype[nk]}
      group:      {"keys": u64[nk], "offsets": u64[nk+1], "vals": vdtype[nv]}
    Keys are in first-insertion order (the reference's HashMap order is unspecified).
    """
    l = lib()
    opc = OPS[op] if isinstance(op, str) else op
    vdt = DTYPES[vdtype] if isinstance(vdtype, str) else vdtype
    k = _as_u64(keys)
    v = _as_u64(vals) if vals is not None else np.zeros(len(k), dtype=np.uint64)
    n = len(k)
    h = l.vo_shuffle_run(opc, vdt, key_width, _p(k), _p(v), n, num_slices, n_reduce, threads)
    if not h:
        raise ValueError("vo_shuffle_run failed")
    try:
        out = []
        out_dt = np.uint64 if opc == COUNT else _NP[vdt]
        for r in range(n_reduce):
            nk, nv = ctypes.c_uint64(), ctypes.c_uint64()
            l.vo_job_part_sizes(h, r, ctypes.byref(nk), ctypes.byref(nv))
            ks = np.empty(nk.value, dtype=np.uint64)
            if opc == GROUP:
                offs = np.empty(nk.value + 1, dtype=np.uint64)
                vs = np.empty(nv.value, dtype=np.uint64)
                l.vo_job_part_copy(h, r, _p(ks), None, _p(offs), _p(vs))
                out.append({"keys": ks, "offsets": offs, "vals": vs.view(_NP[vdt])})
            else:
                cs = np.empty(nk.value, dtype=np.uint64)
                l.vo_job_part_copy(h, r, _p(ks), _p(cs), None, None)
                out.append({"keys": ks, "combined": cs.view(out_dt)})
        return out
    finally:
        l.vo_job_free(h)


def shuffle_timed(op, keys, vals, num_slices, n_reduce, vdtype="u64", threads=1):
    """Time one CPU shuffle (seconds) without copying results out; for the CPU baseline."""
    import time
    l = lib()
    k = _as_u64(keys)
    v = _as_u64(vals)
    t0 = time.perf_counter()
    h = l.vo_shuffle_run(OPS[op], DTYPES[vdtype], 8, _p(k), _p(v), len(k), num_slices, n_reduce, threads)
    t1 = time.perf_counter()
    tot = 0
    for r in range(n_reduce):
        nk, nv = ctypes.c_uint64(), ctypes.c_uint64()
        l.vo_job_part_sizes(h, r, ctypes.byref(nk), ctypes.byref(nv))
        tot += nk.value
    l.vo_job_free(h)
    return t1 - t0, tot


def join(ka, va, slices_a, kb, vb, slices_b, n_reduce, key_width=8, threads=1):
    """Inner join; returns per-partition (k, v, w) u64 arrays."""
    l = lib()
    ka, va, kb, vb = _as_u64(ka), _as_u64(va), _as_u64(kb), _as_u64(vb)
    h = l.vo_join_run(key_width, _p(ka), _p(va), len(ka), slices_a, _p(kb), _p(vb), len(kb), slices_b, n_reduce, threads)
    if not h:
        raise ValueError("vo_join_run failed")
    try:
        out = []
        for r in range(n_reduce):
            n = l.vo_join_part_size(h, r)
            k, v, w = (np.empty(n, dtype=np.uint64) for _ in range(3))
            l.vo_join_part_copy(h, r, _p(k), _p(v), _p(w))
            out.append((k, v, w))
        return out
    finally:
        l.vo_join_free(h)


def sort_by_key(keys, vals, n_parts, kdtype="u64"):
    l = lib()
    k = _as_u64(keys)
    v = _as_u64(vals) if vals is not None else None
    ok = np.empty(len(k), dtype=np.uint64)
    ov = np.empty(len(k), dtype=np.uint64) if v is not None else None
    ps = np.empty(n_parts + 1, dtype=np.uint64)
    l.vo_sort_by_key(DTYPES[kdtype], _p(k), _p(v), len(k), n_parts, _p(ok), _p(ov), _p(ps))
    return ok.view(_NP[DTYPES[kdtype]]), ov, ps


def gen_uniform(first, n, D, seed_k=1, seed_v=2):
    k = np.empty(n, dtype=np.uint64)
    v = np.empty(n, dtype=np.uint64)
    lib().vo_gen_uniform(_p(k), _p(v), first, n, D, seed_k, seed_v)
    return k, v
