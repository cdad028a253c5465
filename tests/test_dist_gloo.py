"""world_size-2 gloo tests of the N>1 path's host logic (vega_b200/dist.py): contiguous map
blocks, the count swap + all-to-all-v, source-major receive order == map-id order, partition
ownership r % world.  The per-rank compute is the oracle-backed FakeEngine (CPU); on the GPU
the same run_shuffle drives CudaEngine (tests/test_gpu_dist.py, bench.py --gpus N)."""
import os
import pickle
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, agg, M, R, outdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.fake_engine import FakeEngine
    from vega_b200 import dist as vdist
    from vega_b200.rdd import slice_starts
    rng = np.random.default_rng(42)                    # every rank builds the same global dataset
    n = 5000
    keys = (rng.integers(0, 300, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    vals = rng.integers(0, 1 << 30, n).astype(np.uint64)
    starts = slice_starts(n, M)
    lo, hi = vdist.map_block(rank, world, M)
    maps = [(m, keys[starts[m]:starts[m + 1]], vals[starts[m]:starts[m + 1]]) for m in range(lo, hi)]
    stats = {}
    sh = vdist.run_shuffle(FakeEngine(), maps, M, R, 0, 0, agg, rank, world, stats=stats)
    res = {}
    for r in range(R):
        out = FakeEngine().reduce(sh, r)
        if r % world != rank:
            assert len(out[0]) == 0, "a rank holds rows of a partition it does not own"
        res[r] = [np.asarray(x) for x in out]
    with open(os.path.join(outdir, f"r{rank}.pkl"), "wb") as f:
        pickle.dump((res, stats), f)
    dist.destroy_process_group()


@pytest.mark.parametrize("agg,M,R", [(1, 4, 4), (0, 4, 6), (4, 5, 3), (0, 3, 2), (3, 2, 8)])
def test_two_rank_shuffle_matches_single_process_oracle(agg, M, R):
    from oracle import oracle as O
    world = 2
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), agg, M, R, d), nprocs=world, join=True)
        per_rank = [pickle.load(open(os.path.join(d, f"r{r}.pkl"), "rb")) for r in range(world)]
    rng = np.random.default_rng(42)
    n = 5000
    keys = (rng.integers(0, 300, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))
    vals = rng.integers(0, 1 << 30, n).astype(np.uint64)
    op = {0: "group", 1: "sum", 2: "min", 3: "max", 4: "count"}[agg]
    want = O.shuffle(op, keys, vals, M, R)
    for r in range(R):
        got = per_rank[r % world][0][r]
        w = want[r]
        if op == "group":
            gd = {int(k): got[2][int(got[1][i]):int(got[1][i + 1])].tolist() for i, k in enumerate(got[0])}
            wd = {int(k): w["vals"][int(w["offsets"][i]):int(w["offsets"][i + 1])].tolist() for i, k in enumerate(w["keys"])}
            assert gd == wd          # value lists in global input order, across the rank boundary
        else:
            assert dict(zip(got[0].tolist(), got[1].tolist())) == dict(zip(w["keys"].tolist(), w["combined"].tolist()))
    sent = sum(s["sent_rows"] for _, s in per_rank)
    recv = sum(s["recv_rows"] for _, s in per_rank)
    assert sent == recv and sent > 0


def test_map_block_assignment_is_contiguous_and_complete():
    from vega_b200.dist import map_block, owned_partitions
    for world in (1, 2, 3, 8):
        for n_map in (1, 5, 8, 64):
            blocks = [map_block(r, world, n_map) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n_map
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        owned = sorted(p for r in range(world) for p in owned_partitions(r, world, 13))
        assert owned == list(range(13))


def test_p2p_arena_layout_is_a_partition_of_every_arena():
    """Fused exchange: the blocks the ranks write into one arena must tile it exactly, source-rank major."""
    from vega_b200.dist import p2p_layout
    rng = np.random.default_rng(4)
    for world in (2, 3, 8):
        allc = rng.integers(0, 1000, (world, world)).tolist()
        layouts = [p2p_layout(allc, r) for r in range(world)]
        for dst in range(world):
            total = layouts[0][0][dst]
            assert all(l[0][dst] == total for l in layouts) and total == sum(allc[src][dst] for src in range(world))
            blocks = sorted((layouts[src][1][dst], allc[src][dst], src) for src in range(world))
            pos = 0
            for off, cnt, src in blocks:
                assert off == pos            # contiguous, no overlap, no gap
                pos += cnt
            assert pos == total
            assert [b[2] for b in blocks if b[1]] == sorted(b[2] for b in blocks if b[1])   # source-rank major
            assert layouts[dst][2] == [allc[src][dst] for src in range(world)]
