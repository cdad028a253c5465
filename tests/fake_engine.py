"""Oracle-backed stand-in for vega_b200.dist.CudaEngine — TEST INFRASTRUCTURE ONLY.

Lets the world_size-2 gloo tests exercise the exchange logic of vega_b200/dist.py (count swap,
split sizes, source-major = map-id order, partition ownership) on CPU.  Same five calls as the
product engine; all arithmetic comes from the oracle."""
import numpy as np
import torch

from oracle import oracle as O

_OPS = {1: "sum", 2: "min", 3: "max", 4: "count"}


class FakeShuffle:
    def __init__(self, n_map, n_reduce, agg, rank, world, key_width):
        self.n_map, self.n_reduce, self.agg, self.rank, self.world, self.key_width = n_map, n_reduce, agg, rank, world, key_width
        self.maps = {}
        self.rows = None
        self.parts = None

    def free(self):
        pass


class FakeEngine:
    def create(self, n_map, n_reduce, kcode, vcode, agg, rank, world, key_width=8, hint=0):
        return FakeShuffle(n_map, n_reduce, agg, rank, world, key_width)

    def map(self, sh, map_id, keys, vals):
        k = np.asarray(keys).view(np.uint64)
        v = np.asarray(vals).view(np.uint64) if vals is not None else np.ones(len(k), dtype=np.uint64)
        sh.maps[map_id] = (k, v)

    def _local_rows(self, sh):
        ids = sorted(sh.maps)
        k = np.concatenate([sh.maps[m][0] for m in ids]) if ids else np.empty(0, np.uint64)
        v = np.concatenate([sh.maps[m][1] for m in ids]) if ids else np.empty(0, np.uint64)
        if sh.agg in _OPS:            # map-side combine: one row per distinct key
            op = _OPS[sh.agg]
            parts = O.shuffle(op, k, v, 1, 1)
            k, v = parts[0]["keys"], parts[0]["combined"].view(np.uint64)
        return k, v

    def export(self, sh, world):
        k, v = self._local_rows(sh)
        dest = np.array([O.get_partition(int(x), sh.n_reduce, sh.key_width) % world for x in k], dtype=np.int64)
        order = np.argsort(dest, kind="stable")
        counts = [int((dest == r).sum()) for r in range(world)]
        tk = torch.from_numpy(k[order].view(np.int64).copy())
        tv = torch.from_numpy(v[order].view(np.int64).copy())
        return counts, tk, tv

    def import_(self, sh, keys, vals, counts):
        sh.rows = (keys.numpy().view(np.uint64), vals.numpy().view(np.uint64))

    def seal(self, sh):
        if sh.world == 1:
            sh.rows = self._local_rows(sh)
        k, v = sh.rows
        R = sh.n_reduce
        if sh.agg in _OPS:
            op = "sum" if sh.agg == 4 else _OPS[sh.agg]     # merge of counts is a sum
            sh.parts = O.shuffle(op, k, v, 1, R, key_width=sh.key_width)
        else:
            sh.parts = O.shuffle("group", k, v, 1, R, key_width=sh.key_width)

    def reduce(self, sh, r):
        p = sh.parts[r]
        if "offsets" in p:
            return p["keys"], p["offsets"], p["vals"]
        return p["keys"], p["combined"]
