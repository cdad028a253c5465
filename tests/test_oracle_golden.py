"""Pins the CPU oracle against every golden vector the reference's own tests hold
for the shuffle path.  Expected values are transcribed from
/root/reference/tests/*.rs (cited per test); nothing here reads /root/reference."""
import numpy as np
import pytest

from oracle import oracle as O
from oracle import pyref as P


# -- third-party hash (fasthash 0.4.0 / MetroHash64_1): published known-answer vector ------------
def test_metrohash64_1_published_vector():
    # MetroHash's own test key (63 bytes: exercises the 32-byte loop and the 16/8/4/2/1 tails)
    key = b"012345678901234567890123456789012345678901234567890123456789012"
    assert O.metrohash64_1(key, 0).to_bytes(8, "little").hex().upper() == "658F044F5C730E40"
    assert O.metrohash64_1(key, 1).to_bytes(8, "little").hex().upper() == "AE49EBB0A856537B"


def test_hash_partition_like_reference():
    # src/partitioner.rs:63-82 asserts only that 3 keys land in [0, n): the sole reference test
    for n in (1, 2, 3, 4, 8, 64, 1000):
        for k in (1, 2, 3, 2 ** 63, 2 ** 64 - 1):
            assert 0 <= O.get_partition(k, n) < n
    assert O.get_partition(5, 1) == 0


# -- ParallelCollection::slice (parallel_collection_rdd.rs:116-145) ---------------------------
@pytest.mark.parametrize("n,m", [(15, 4), (9, 4), (9, 2), (7, 2), (100, 20), (4, 4), (101, 101), (1000, 7)])
def test_slice_closed_form(n, m):
    st = O.slice_starts(n, m)
    assert len(st) == m + 1
    assert list(st) == [(s * n) // m for s in range(m)] + [n]
    assert [len(x) for x in P.slice_(range(n), m)] == [int(st[i + 1] - st[i]) for i in range(m)]


def test_slice_quirk_fewer_rows_than_slices():
    # make_rdd(0..10, 32) → 11 splits (empty leading slice then singletons)
    assert [len(x) for x in P.slice_(range(10), 32)] == [0] + [1] * 10
    assert list(O.slice_starts(10, 32)) == [0] + list(range(0, 10)) + [10]
    assert list(O.slice_starts(0, 4)) == [0, 0]


# -- tests/test_pair_rdd.rs:8-37 test_group_by_key (ordered groups, F5) -----------------------
GROUP_ROWS = [("x", i) for i in range(1, 8)] + [("y", i) for i in range(1, 9)]


def test_group_by_key_golden():
    res = sorted(P.group_by_key(GROUP_ROWS, 4, 4))
    assert res == [("x", [1, 2, 3, 4, 5, 6, 7]), ("y", [1, 2, 3, 4, 5, 6, 7, 8])]


def test_group_by_key_golden_typed_oracle():
    # same data through the typed C oracle with x→10, y→20
    keys = np.array([10] * 7 + [20] * 8, dtype=np.uint64)
    vals = np.array(list(range(1, 8)) + list(range(1, 9)), dtype=np.uint64)
    parts = O.shuffle("group", keys, vals, 4, 4)
    got = {}
    for p in parts:
        for i, k in enumerate(p["keys"]):
            got[int(k)] = list(p["vals"][int(p["offsets"][i]):int(p["offsets"][i + 1])])
    assert got == {10: [1, 2, 3, 4, 5, 6, 7], 20: [1, 2, 3, 4, 5, 6, 7, 8]}


# -- tests/test_pair_rdd.rs:39-82 test_join (inner join, 6 rows) ------------------------------
COL1 = [(1, ("A", "B")), (2, ("C", "D")), (3, ("E", "F")), (4, ("G", "H"))]
COL2 = [(1, "A1"), (1, "A2"), (2, "B1"), (2, "B2"), (3, "C1"), (3, "C2")]
JOIN_EXPECTED = [(1, ("A1", ("A", "B"))), (1, ("A2", ("A", "B"))), (2, ("B1", ("C", "D"))),
                 (2, ("B2", ("C", "D"))), (3, ("C1", ("E", "F"))), (3, ("C2", ("E", "F")))]


def test_join_golden():
    res = sorted(P.join(COL2, 4, COL1, 4, 4, int_width=4))
    assert res == JOIN_EXPECTED


def test_join_golden_typed_oracle():
    # payload strings → indices into the tables
    kb = np.array([k for k, _ in COL1], dtype=np.uint64)
    vb = np.arange(len(COL1), dtype=np.uint64)
    ka = np.array([k for k, _ in COL2], dtype=np.uint64)
    va = np.arange(len(COL2), dtype=np.uint64)
    rows = []
    for k, v, w in O.join(ka, va, 4, kb, vb, 4, 4, key_width=4):
        rows += [(int(a), (COL2[int(b)][1], COL1[int(c)][1])) for a, b, c in zip(k, v, w)]
    assert sorted(rows) == JOIN_EXPECTED


# -- tests/test_pair_rdd.rs:84-109 test_count_by_value (reduce_by_key(+), R = 4 and 2) --------
@pytest.mark.parametrize("slices", [4, 2])
def test_count_by_value_golden(slices):
    xs = [1, 2, 1, 3, 2, 3, 3, 2, 3]
    assert sorted(P.count_by_value(xs, slices, int_width=4)) == [(1, 2), (2, 3), (3, 4)]
    keys = np.array(xs, dtype=np.uint64)
    parts = O.shuffle("count", keys, None, slices, slices, key_width=4)
    got = sorted((int(k), int(c)) for p in parts for k, c in zip(p["keys"], p["combined"]))
    assert got == [(1, 2), (2, 3), (3, 4)]


# -- tests/test_pair_rdd.rs:111-135 test_group_by (ordered groups) -----------------------------
def test_group_by_golden():
    xs = [-3, -2, -1, 0, 1, 2, 3]
    rows = [("pos" if x > 0 else "neg" if x < 0 else "zero", x) for x in xs]
    res = sorted(P.group_by_key(rows, 2, 2))
    assert res == [("neg", [-3, -2, -1]), ("pos", [1, 2, 3]), ("zero", [0])]


# -- tests/test_rdd.rs:285-322 test_distinct ---------------------------------------------------
@pytest.mark.parametrize("parts", [3, 2, 10])
def test_distinct_golden(parts):
    xs = [1, 2, 2, 2, 3, 3, 3, 4, 4, 5]
    res = P.distinct(xs, 3, parts, int_width=4)
    assert len(res) == 5 and set(res) == {1, 2, 3, 4, 5}


# -- tests/test_rdd.rs:387-432 test_union (two joins → 12 rows) --------------------------------
def test_union_of_joins_golden():
    assert len(P.join(COL2, 4, COL1, 4, 4, 4)) * 2 == 12


# -- tests/test_rdd.rs:434-456 CoGroupedRdd with HashPartitioner(2) → 4 keys per cogroup -------
def test_cogroup_unique_partitioner_golden():
    rdd = [(1, "A"), (2, "B"), (3, "C"), (4, "D")]
    cg = P.cogroup(rdd, 2, rdd, 2, 2, int_width=4)
    assert len(cg) * 2 == 8
    assert sorted(cg) == [(k, [[v], [v]]) for k, v in rdd]


# -- tests/test_rdd.rs:484-521 intersection, :675-699 subtract (cogroup users) -----------------
C1 = [1, 2, 3, 4, 5, 10, 12, 13, 19, 0]
C2 = [3, 4, 5, 6, 7, 8, 11, 13]


def _intersection(a, sa, b, sb, r):
    cg = P.cogroup([(x, None) for x in a], sa, [(x, None) for x in b], sb, r, int_width=4)
    return [k for k, (v1, v2) in cg if len(v1) >= 1 and len(v2) >= 1]


def test_intersection_golden():
    assert sorted(_intersection(C1, 2, C2, 4, 3)) == [3, 4, 5, 13]
    assert sorted(_intersection(C1, 2, C2, 4, 2)) == [3, 4, 5, 13]


def test_subtract_golden():
    # rdd.rs:852-900: xor-cogroup, then intersection with self
    cg = P.cogroup([(x, None) for x in C1], 4, [(x, None) for x in C2], 4, 4, int_width=4)
    xor = [k for k, (v1, v2) in cg if (len(v1) >= 1) ^ (len(v2) >= 1)]
    assert sorted(_intersection(C1, 4, xor, 4, 4)) == [0, 1, 2, 10, 12, 19]


# -- typed C oracle == generic Python restatement on random data -------------------------------
@pytest.mark.parametrize("op", ["sum", "min", "max", "count", "group"])
def test_c_oracle_matches_pyref(op):
    rng = np.random.default_rng(7)
    n, M, R = 3000, 5, 7
    keys = rng.integers(0, 97, n).astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)
    vals = rng.integers(0, 1 << 40, n).astype(np.uint64)
    rows = list(zip(keys.tolist(), vals.tolist()))
    parts = O.shuffle(op, keys, vals, M, R)
    if op == "group":
        want = sorted(P.group_by_key(rows, M, R))
        got = sorted((int(k), p["vals"][int(p["offsets"][i]):int(p["offsets"][i + 1])].tolist())
                     for p in parts for i, k in enumerate(p["keys"]))
    else:
        f = {"sum": lambda a, b: (a + b) & (2 ** 64 - 1), "min": min, "max": max, "count": lambda a, b: a + b}[op]
        r2 = rows if op != "count" else [(k, 1) for k, _ in rows]
        want = sorted(P.reduce_by_key(r2, f, M, R))
        got = sorted((int(k), int(c)) for p in parts for k, c in zip(p["keys"], p["combined"]))
    assert got == want
    # placement: every key sits in the partition the partitioner names
    for r, p in enumerate(parts):
        assert all(O.get_partition(int(k), R) == r for k in p["keys"])


def test_c_oracle_f64_and_i64():
    rng = np.random.default_rng(3)
    n = 2000
    keys = rng.integers(0, 50, n).astype(np.uint64)
    fv = rng.standard_normal(n)
    parts = O.shuffle("sum", keys, fv, 4, 4, vdtype="f64")
    got = {int(k): float(c) for p in parts for k, c in zip(p["keys"], p["combined"])}
    for k in np.unique(keys):
        assert got[int(k)] == pytest.approx(fv[keys == k].sum(), rel=1e-9)
    iv = rng.integers(-10 ** 9, 10 ** 9, n).astype(np.int64)
    for op, fn in (("min", np.min), ("max", np.max), ("sum", np.sum)):
        parts = O.shuffle(op, keys, iv, 3, 5, vdtype="i64")
        got = {int(k): int(c) for p in parts for k, c in zip(p["keys"], p["combined"])}
        assert got == {int(k): int(fn(iv[keys == k])) for k in np.unique(keys)}


def test_sort_oracle():
    rng = np.random.default_rng(5)
    keys = rng.integers(0, 1000, 5000).astype(np.uint64)
    vals = np.arange(5000, dtype=np.uint64)
    ok, ov, ps = O.sort_by_key(keys, vals, 8)
    order = np.argsort(keys, kind="stable")
    assert (ok == keys[order]).all() and (ov == vals[order]).all()
    assert ps[0] == 0 and ps[-1] == 5000 and (np.diff(ps.astype(np.int64)) >= 0).all()
    for c in ps[1:-1]:
        assert c == 0 or c == 5000 or ok[int(c)] != ok[int(c) - 1]


def test_generator_matches_numpy():
    k, v = O.gen_uniform(5, 1000, 37)
    def sm(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        z = x
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))
    i = np.arange(5, 1005, dtype=np.uint64)
    with np.errstate(over="ignore"):
        rank = sm(np.uint64(1) + i) % np.uint64(37)
        assert (k == sm(rank ^ np.uint64(0xA5A5A5A5A5A5A5A5))).all()
        assert (v == (sm(np.uint64(2) + i) & np.uint64(0xFFFFF))).all()


# -- bincode blobs: the reference's only fixture is a round trip -------------
def test_bincode_reference_fixture_round_trip():
    from oracle import bincode_ref as B
    # src/shuffle/shuffle_fetcher.rs:154-166: vec![(0i32, "example data")] at key (11000, 0, 11001)
    blob = B.encode_i32_string_vec([(0, "example data")])
    assert len(blob) == 8 + 4 + 8 + 12 and blob[:8] == (1).to_bytes(8, "little")
    assert B.decode_i32_string_vec(blob) == [(0, "example data")]
    # :168-185 fetch_failure: a payload of another type does not decode
    with pytest.raises(Exception):
        B.decode_i32_string_vec(b"\x0e\x00\x00\x00\x00\x00\x00\x00corrupted data")
    pairs = [(1, 2), (2 ** 64 - 1, 0)]
    assert B.decode_pairs(B.encode_pairs(pairs)) == pairs
    groups = [(7, [1, 2, 3]), (9, [])]
    assert B.decode_groups(B.encode_groups(groups)) == groups
    assert len(B.encode_groups(groups)) == 8 + 16 * 2 + 8 * 3


# -- the same vectors as committed fixtures (tests/golden/reference_vectors.json) -----------------
def test_committed_golden_fixture_matches_oracle():
    import json
    import os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
    c = g["group_by_key"]
    rows = [(k, v) for k, v in c["rows"]]
    assert sorted(P.group_by_key(rows, c["num_slices"], c["num_splits"])) == [(k, v) for k, v in c["expected_sorted"]]
    c = g["join"]
    col1 = [(k, tuple(v)) for k, v in c["col1"]]
    col2 = [(k, v) for k, v in c["col2"]]
    want = [(k, (a, tuple(b))) for k, (a, b) in c["expected_sorted"]]
    assert sorted(P.join(col2, 4, col1, 4, 4, int_width=4)) == want
    c = g["count_by_value"]
    for ns in c["num_slices"]:
        assert sorted(P.count_by_value(c["values_i32"], ns, int_width=4)) == [tuple(x) for x in c["expected_sorted"]]
    c = g["intersection"]
    for nsp in c["num_splits"]:
        assert sorted(_intersection(c["col1"], c["slices"][0], c["col2"], c["slices"][1], nsp)) == c["expected_sorted"]
    c = g["metrohash64_1_kat"]
    assert O.metrohash64_1(c["key"].encode(), 0).to_bytes(8, "little").hex().upper() == c["seed0_le_hex"]
    assert O.metrohash64_1(c["key"].encode(), 1).to_bytes(8, "little").hex().upper() == c["seed1_le_hex"]
