"""One-process-per-GPU shuffle: the data plane between map side and reduce side.

The reference moves shuffle blocks with one HTTP GET per (map, reduce) pair
(src/shuffle/shuffle_fetcher.rs:61-86 → src/shuffle/shuffle_manager.rs:176-251).  Here every
rank packs its map output by destination rank on the device (vb_shuffle_export_prepare: a
stable multisplit, so rows stay in map-id / encounter order), the ranks swap a world×world
count vector, and ONE all-to-all-v per column (torch.distributed → NCCL grouped send/recv over
NVLink/NVSwitch; gloo on CPU for the host-logic tests) delivers them.  Reduce partition r is
owned by rank r % world; map partitions are assigned to ranks in contiguous blocks so that the
source-rank-major receive order IS map-id order (group value order).

reduce ops exchange *combined* rows (≤ distinct keys per rank — map-side combine,
src/dependency.rs:203-209), group ops exchange raw rows.

The exchange itself lives in libvega_b200.so (vb_ctx_comm_init + vb_shuffle_exchange: NCCL grouped
send/recv, or the fused peer-memory scatter, on the library's stream).  `run_shuffle` calls it whenever the
context has a communicator; the torch.distributed code paths below remain for the two cases the library's
NCCL communicator cannot cover: the CPU (gloo) tests of the host logic with an oracle-backed stand-in engine,
and two ranks sharing ONE GPU on a single-GPU test box (NCCL refuses duplicate devices).
"""
import ctypes

import numpy as np

from . import _lib as L
from .rdd import Shuffle, _Col


def map_block(rank, world, n_map):
    """Contiguous block of map partitions owned by `rank`: [lo, hi)."""
    return (rank * n_map) // world, ((rank + 1) * n_map) // world


class _DevArray:
    """Zero-copy view of a library-owned device buffer for torch.as_tensor."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class CudaEngine:
    """libvega_b200 on this rank's GPU."""

    def __init__(self, sc):
        self.sc = sc

    def create(self, n_map, n_reduce, kcode, vcode, agg, rank, world, key_width=8, hint=0):
        return Shuffle(self.sc, n_map, n_reduce, kcode, vcode, agg, key_width=key_width, hint=hint, rank=rank, world=world)

    def map(self, sh, map_id, keys, vals):
        k = keys if isinstance(keys, _Col) else _Col(keys, allow_rows=True)
        v = None if vals is None else (vals if isinstance(vals, _Col) else _Col(vals, role="value"))
        sh.map(map_id, k, v, 0, k.n)

    def export(self, sh, world):
        import torch
        counts = (ctypes.c_uint64 * world)()
        L.check(sh._lib.vb_shuffle_export_prepare(sh._h, counts))
        counts = [int(c) for c in counts]
        kp, vp = ctypes.c_void_p(), ctypes.c_void_p()
        L.check(sh._lib.vb_shuffle_export_buffers(sh._h, ctypes.byref(kp), ctypes.byref(vp)))
        n = sum(counts)
        dev = f"cuda:{sh._lib.vb_ctx_device(self.sc._h)}"
        if n == 0:
            e = torch.empty(0, dtype=torch.int64, device=dev)
            return counts, e, e.clone()
        keys = torch.as_tensor(_DevArray(kp.value, n), device=dev)
        vals = torch.as_tensor(_DevArray(vp.value, n), device=dev)
        return counts, keys, vals

    def import_(self, sh, keys, vals, counts):
        import torch
        if keys.is_cuda:
            # NCCL collectives are stream-ordered on torch's stream; the library reads on its own stream
            torch.cuda.current_stream(keys.device).synchronize()
        sh._imported = (keys, vals)            # keep the tensors alive until seal
        c = (ctypes.c_uint64 * len(counts))(*counts)
        p = lambda t: ctypes.c_void_p(t.data_ptr()) if t.numel() else None
        L.check(sh._lib.vb_shuffle_import(sh._h, p(keys), p(vals), c))

    def seal(self, sh):
        sh.seal()
        sh._imported = None

    def reduce(self, sh, r):
        return sh.reduce(r)


def p2p_layout(allc, rank):
    """Arena layout of the fused exchange from the all-gathered count matrix allc[src][dst]:
    (total rows each rank receives, row offset of `rank`'s block inside every destination's arena,
    rows `rank` receives from every source).  Blocks are source-rank major, so the arena order is map-id order."""
    world = len(allc)
    total_recv = [sum(allc[src][dst] for src in range(world)) for dst in range(world)]
    my_off = [sum(allc[src][dst] for src in range(rank)) for dst in range(world)]
    recv_counts = [allc[src][rank] for src in range(world)]
    return total_recv, my_off, recv_counts


def p2p_exchange(engine, sh, rank, world, group=None, stats=None):
    """Fused exchange of a GROUP/COGROUP shuffle: the partition kernel of every rank stores its rows straight
    into the owners' HBM (CUDA IPC peer mappings over NVLink) — see include/vega_b200.h.  torch.distributed only
    carries the control plane: two small all-gathers (counts, IPC handles) and one barrier."""
    import torch
    import torch.distributed as dist
    lib, sc = sh._lib, engine.sc
    t0 = None
    if stats is not None:
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
    counts = (ctypes.c_uint64 * world)()
    L.check(lib.vb_shuffle_export_counts(sh._h, counts))
    counts = [int(c) for c in counts]
    allc = [None] * world
    dist.all_gather_object(allc, counts, group=group)            # allc[src][dst]
    total_recv, my_off, recv_counts = p2p_layout(allc, rank)
    handle = (ctypes.c_ubyte * 64)()
    gen = ctypes.c_uint64()
    L.check(lib.vb_ctx_arena_reserve(sc._h, 16 * max(total_recv[rank], 1), handle, ctypes.byref(gen)))
    allh = [None] * world
    dist.all_gather_object(allh, (bytes(handle), gen.value), group=group)   # doubles as "every arena is reserved"
    for peer, (hb, g) in enumerate(allh):
        buf = (ctypes.c_ubyte * 64).from_buffer_copy(hb)
        L.check(lib.vb_ctx_peer_open(sc._h, peer, buf, g, int(peer == rank)))
    off = (ctypes.c_uint64 * world)(*my_off)
    tot = (ctypes.c_uint64 * world)(*total_recv)
    L.check(lib.vb_shuffle_export_direct(sh._h, off, tot))        # returns when this rank's stores are done
    dist.barrier(group=group)                                     # everybody's rows have landed
    L.check(lib.vb_ctx_arena_release_retired(sc._h))              # every peer re-opened: outgrown arenas can go
    rc = (ctypes.c_uint64 * world)(*recv_counts)
    L.check(lib.vb_shuffle_import_arena(sh._h, rc))
    if stats is not None:
        torch.cuda.synchronize()
        stats["exchange_ms"] = stats.get("exchange_ms", 0.0) + (time.perf_counter() - t0) * 1e3
        stats["sent_rows"] = sum(counts) - counts[rank]
        stats["recv_rows"] = total_recv[rank] - allc[rank][rank]
        stats["exchanges"] = stats.get("exchanges", 0) + 1
        stats["exchange_kind"] = "p2p"


def all_to_all_v(send_keys, send_vals, send_counts, group=None):
    """The shuffle's single exchange step: counts, then keys and values, each one
    all_to_all_single.  Returns (recv_keys, recv_vals, recv_counts), source-rank major."""
    import torch
    import torch.distributed as dist
    dev = send_keys.device
    cin = torch.tensor(send_counts, dtype=torch.int64, device=dev)
    cout = torch.empty_like(cin)
    dist.all_to_all_single(cout, cin, group=group)
    recv_counts = [int(x) for x in cout.tolist()]
    rk = torch.empty(sum(recv_counts), dtype=send_keys.dtype, device=dev)
    rv = torch.empty(sum(recv_counts), dtype=send_vals.dtype, device=dev)
    dist.all_to_all_single(rk, send_keys, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    dist.all_to_all_single(rv, send_vals, output_split_sizes=recv_counts, input_split_sizes=list(send_counts), group=group)
    return rk, rv, recv_counts


def run_shuffle(engine, local_maps, n_map, n_reduce, kcode, vcode, agg, rank, world, group=None, key_width=8, hint=0,
                stats=None, exchange_device=None, p2p=False):
    """Map side on this rank's partitions → exchange → reduce side for the partitions this rank owns.

    local_maps: [(map_id, keys, vals_or_None)], ascending map ids from map_block(rank, world, n_map).
    Returns the sealed engine handle; engine.reduce(h, r) is non-empty only for r % world == rank.
    exchange_device: move the packed buffers there for the collective (tests: "cpu" + gloo on a 1-GPU box).
    p2p: GROUP/COGROUP shuffles use the fused partition+send over peer memory instead of pack + all-to-all-v.
    """
    sh = engine.create(n_map, n_reduce, kcode, vcode, agg, rank, world, key_width=key_width, hint=hint)
    for map_id, keys, vals in local_maps:
        engine.map(sh, map_id, keys, vals)
    if world > 1 and isinstance(engine, CudaEngine) and getattr(engine.sc, "comm", None) and exchange_device is None:
        # the product path: count exchange + ONE grouped NCCL send/recv (or the fused P2P scatter) inside
        # libvega_b200, on the library's stream; torch.distributed is not involved
        group_op = agg in (L.VB_AGG_GROUP, L.VB_AGG_COGROUP)
        sh.exchange(L.VB_XCHG_P2P if (p2p and group_op) else L.VB_XCHG_NCCL)
        engine.seal(sh)
        if stats is not None:           # after the seal: reading the event timers synchronises the stream
            for k_, v_ in sh.exchange_stats().items():
                stats[k_] = (stats.get(k_, 0) + v_) if k_ in ("exchange_ms", "exchanges", "prepare_wall_ms", "counts_wall_ms", "post_wall_ms") else v_
        return sh
    elif world > 1 and p2p and agg in (L.VB_AGG_GROUP, L.VB_AGG_COGROUP) and isinstance(engine, CudaEngine):
        p2p_exchange(engine, sh, rank, world, group, stats)
    elif world > 1:
        counts, sk, sv = engine.export(sh, world)
        home = sk.device
        if exchange_device is not None:        # e.g. "cpu" for a gloo group: stage the exchange through the host
            sk, sv = sk.to(exchange_device), sv.to(exchange_device)
        ev = None
        if stats is not None and sk.is_cuda:
            import torch
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record()
        rk, rv, rcounts = all_to_all_v(sk, sv, counts, group)
        if ev is not None:
            ev[1].record()
            ev[1].synchronize()
            stats["exchange_ms"] = stats.get("exchange_ms", 0.0) + ev[0].elapsed_time(ev[1])
        if exchange_device is not None:
            rk, rv = rk.to(home), rv.to(home)
        if stats is not None:
            stats["sent_rows"] = sum(counts) - counts[rank]
            stats["recv_rows"] = sum(rcounts) - rcounts[rank]
            stats["exchanges"] = stats.get("exchanges", 0) + 1
        engine.import_(sh, rk, rv, rcounts)
    engine.seal(sh)
    return sh


def owned_partitions(rank, world, n_reduce):
    return [r for r in range(n_reduce) if r % world == rank]
