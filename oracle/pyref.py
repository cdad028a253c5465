"""Pure-Python restatement of vega's shuffle path for arbitrary keys/values.

TEST INFRASTRUCTURE ONLY (small cases: the reference's own golden vectors use
String / i32 / tuple payloads that the typed C oracle does not model).
Follows SURVEY.md Appendix A; each function cites the reference file:line.
Python dicts iterate in insertion order, which stands in for the reference's
HashMap (unspecified order) — results are compared sorted, value lists in order.
"""
import struct

from . import oracle as _o


def key_bytes(k, int_width=8):
    """Bytes Rust's `Hash` impl feeds the hasher for the key types the tests use.
    u64/i64 → 8 LE bytes, i32 → 4 LE bytes, String → utf8 + 0xFF (str::hash)."""
    if isinstance(k, str):
        return k.encode() + b"\xff"
    if isinstance(k, int):
        return struct.pack("<q" if int_width == 8 else "<i", k) if k < 0 else \
            struct.pack("<Q" if int_width == 8 else "<I", k)
    if isinstance(k, tuple):
        return b"".join(key_bytes(x, int_width) for x in k)
    if k is None:
        return struct.pack("<Q", 0)
    raise TypeError(type(k))


def get_partition(k, n_reduce, int_width=8):
    """HashPartitioner::get_partition — src/partitioner.rs:54-57."""
    return _o.metrohash64_1(key_bytes(k, int_width), 0) % n_reduce


def slice_(data, num_slices):
    """ParallelCollection::slice — src/rdd/parallel_collection_rdd.rs:116-145 (literal)."""
    if num_slices < 1:
        raise ValueError("Number of slices should be greater than or equal to 1")
    data = list(data)
    n = len(data)
    slice_count, it = 0, 0
    end = ((slice_count + 1) * n) // num_slices
    out, tmp = [], []
    for x in data:
        if it < end:
            tmp.append(x)
            it += 1
        else:
            slice_count += 1
            end = ((slice_count + 1) * n) // num_slices
            out.append(tmp)
            tmp = [x]
            it += 1
    out.append(tmp)
    return out


class Aggregator:
    """src/aggregator.rs:8-52."""

    def __init__(self, create_combiner, merge_value, merge_combiners):
        self.create_combiner, self.merge_value, self.merge_combiners = create_combiner, merge_value, merge_combiners

    @staticmethod
    def default():  # :33-52 — Vec append
        return Aggregator(lambda v: [v], lambda c, v: c + [v], lambda a, b: a + b)

    @staticmethod
    def reducing(f):  # src/rdd/pair_rdd.rs:74-78
        return Aggregator(lambda v: v, f, f)


def map_task(split_rows, n_reduce, agg, int_width=8):
    """ShuffleDependency::do_shuffle_task — src/dependency.rs:164-229."""
    buckets = [dict() for _ in range(n_reduce)]
    for k, v in split_rows:
        b = buckets[get_partition(k, n_reduce, int_width)]
        if k in b:
            b[k] = agg.merge_value(b[k], v)          # :203-206
        else:
            b[k] = agg.create_combiner(v)            # :208
    return [list(b.items()) for b in buckets]        # → SHUFFLE_CACHE[(sid, map, r)] :212-223


def reduce_task(r, map_outputs, agg):
    """ShuffledRdd::compute — src/rdd/shuffled_rdd.rs:149-170; fetch order
    src/shuffle/shuffle_fetcher.rs:34-39,63-97 (map-id ascending in local mode)."""
    comb = {}
    for m in range(len(map_outputs)):
        for k, c in map_outputs[m][r]:
            comb[k] = agg.merge_combiners(comb[k], c) if k in comb else c
    return list(comb.items())


def combine_by_key(rows, num_slices, n_reduce, agg, int_width=8):
    """PairRdd::combine_by_key + collect — pair_rdd.rs:20-33, rdd.rs:420-434."""
    splits = slice_(rows, num_slices)
    outs = [map_task(s, n_reduce, agg, int_width) for s in splits]
    res = []
    for r in range(n_reduce):
        res.extend(reduce_task(r, outs, agg))
    return res


def group_by_key(rows, num_slices, n_reduce, int_width=8):
    return combine_by_key(rows, num_slices, n_reduce, Aggregator.default(), int_width)


def reduce_by_key(rows, f, num_slices, n_reduce, int_width=8):
    return combine_by_key(rows, num_slices, n_reduce, Aggregator.reducing(f), int_width)


def cogroup(rows_a, slices_a, rows_b, slices_b, n_reduce, int_width=8):
    """CoGroupedRdd::compute — src/rdd/co_grouped_rdd.rs:206-249; each parent is shuffled
    with a Vec-append aggregator (:78-124)."""
    agg = Aggregator.default()
    outs = [[map_task(s, n_reduce, agg, int_width) for s in slice_(rows, ns)]
            for rows, ns in ((rows_a, slices_a), (rows_b, slices_b))]
    res = []
    for r in range(n_reduce):
        g = {}
        for dep_num in (0, 1):
            for m in range(len(outs[dep_num])):
                for k, vals in outs[dep_num][m][r]:
                    g.setdefault(k, [[], []])[dep_num].extend(vals)
        res.extend(g.items())
    return res


def join(rows_a, slices_a, rows_b, slices_b, n_reduce, int_width=8):
    """PairRdd::join — src/rdd/pair_rdd.rs:104-121: for v in vs { for w in ws } (inner)."""
    out = []
    for k, (vs, ws) in cogroup(rows_a, slices_a, rows_b, slices_b, n_reduce, int_width):
        for v in vs:
            for w in ws:
                out.append((k, (v, w)))
    return out


def count_by_value(xs, num_slices, int_width=8):
    """Rdd::count_by_value — src/rdd/rdd.rs:450-459: map(x→(x,1u64)).reduce_by_key(+, num_splits)."""
    n_splits = len(slice_(xs, num_slices))
    return reduce_by_key([(x, 1) for x in xs], lambda a, b: a + b, num_slices, n_splits, int_width)


def distinct(xs, num_slices, n_reduce, int_width=8):
    """Rdd::distinct_with_num_partitions — src/rdd/rdd.rs:502-522:
    map(x→(Some(x),None)).reduce_by_key((_, y)→y).map(k→k.unwrap())."""
    rows = [(x, None) for x in xs]
    return [k for k, _ in reduce_by_key(rows, lambda a, b: b, num_slices, n_reduce, int_width)]
