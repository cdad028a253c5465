"""N4: the tracker-style TCP rendezvous that hands out the NCCL unique id (host logic, no GPU)."""
import multiprocessing as mp
import os
import socket

import pytest

from vega_b200 import rendezvous as rz


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _client(addr, rank, world, q):
    q.put((rank, rz.fetch_unique_id(addr, rank, world, timeout=20.0)))


def test_all_ranks_receive_the_same_id_even_if_they_start_before_the_master():
    world, payload = 4, bytes(range(128))
    addr = ("127.0.0.1", _free_port())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_client, args=(addr, r, world, q)) for r in range(1, world)]
    for p in procs:
        p.start()                      # clients first: they retry until the master listens (map_output_tracker.rs:69-74)
    srv = rz.TrackerServer(addr, world, payload, timeout=30.0).serve()
    got0 = rz.fetch_unique_id(addr, 0, world, timeout=20.0)
    got = dict(q.get(timeout=30) for _ in procs)
    for p in procs:
        p.join(10)
    srv.join(10)
    assert got0 == payload and all(v == payload for v in got.values()) and set(got) == {1, 2, 3}
    assert srv.seen == {0, 1, 2, 3}


def test_wrong_world_is_refused_and_missing_ranks_time_out():
    addr = ("127.0.0.1", _free_port())
    srv = rz.TrackerServer(addr, 2, b"x" * 128, timeout=1.0).serve()
    with pytest.raises(ValueError):
        rz.fetch_unique_id(addr, 0, 3, timeout=5.0)          # world mismatch
    assert rz.fetch_unique_id(addr, 0, 2, timeout=5.0) == b"x" * 128
    with pytest.raises(TimeoutError):
        srv.join(5.0)                                        # rank 1 never came
    with pytest.raises(TimeoutError):
        rz.fetch_unique_id(("127.0.0.1", _free_port()), 0, 2, timeout=0.3)   # nobody listening
