"""Shared helpers for the parity tests (test infrastructure; may use oracle/)."""
import numpy as np

from oracle import oracle as O


def oracle_reduce(op, keys, vals, M, R, vdtype="u64", key_width=8):
    parts = O.shuffle(op, keys, vals, M, R, vdtype=vdtype, key_width=key_width)
    return [dict(zip(p["keys"].tolist(), p["combined"].tolist())) for p in parts]


def oracle_group(keys, vals, M, R, vdtype="u64", key_width=8):
    parts = O.shuffle("group", keys, vals, M, R, vdtype=vdtype, key_width=key_width)
    out = []
    for p in parts:
        o = p["offsets"]
        out.append({int(k): p["vals"][int(o[i]):int(o[i + 1])].tolist() for i, k in enumerate(p["keys"])})
    return out


def gpu_reduce_parts(rdd):
    """Per-partition {key: combined} of a vega_b200 ShuffledRdd (reduce ops)."""
    out = []
    for r in range(rdd.num_slices):
        k, c = rdd.compute(r)
        d = dict(zip(k.tolist(), c.tolist()))
        assert len(d) == len(k), "duplicate key inside one reduce partition"
        out.append(d)
    return out


def gpu_group_parts(rdd):
    out = []
    for r in range(rdd.num_slices):
        k, o, v = rdd.compute(r)
        assert o[0] == 0 and o[-1] == len(v)
        d = {int(kk): v[int(o[i]):int(o[i + 1])].tolist() for i, kk in enumerate(k)}
        assert len(d) == len(k), "duplicate key inside one reduce partition"
        out.append(d)
    return out


def rand_pairs(rng, n, n_keys, vdtype="u64", wide_keys=True):
    ranks = rng.integers(0, max(n_keys, 1), n).astype(np.uint64)
    keys = ranks * np.uint64(0x9E3779B97F4A7C15) + np.uint64(12345) if wide_keys else ranks
    if vdtype == "u64":
        vals = rng.integers(0, 1 << 40, n).astype(np.uint64)
    elif vdtype == "i64":
        vals = rng.integers(-(1 << 40), 1 << 40, n).astype(np.int64)
    else:
        vals = rng.standard_normal(n)
    return keys, vals
