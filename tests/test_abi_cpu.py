"""CPU-side checks of the boundary: the C-ABI library loads and exports every symbol
include/vega_b200.h declares, its host-side pieces of the path agree with the oracle, and the
product fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import vega_b200 as vb
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "vega_b200.h")).read()
    return sorted(set(re.findall(r"VB_API\s+[\w\s\*]+?\b(vb_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(vb.LIB_PATH), "run __graft_entry__.build() first"
    names = _header_symbols()
    assert len(names) >= 30
    l = ctypes.CDLL(vb.LIB_PATH)
    for n in names:
        assert hasattr(l, n), f"{n} declared in include/vega_b200.h but not exported"
    assert set(names) == set(vb.SYMBOLS), "python binding and header disagree"
    assert vb.lib().vb_version().startswith(b"vega_b200")


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(vb.VegaB200Error) as e:
        vb.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under vega_b200/ or include/ may import, link,
    dlopen or #include it."""
    bad = re.compile(r"(^\s*(import|from)\s+oracle\b)|(libvega_oracle)|(#include\s*[\"<][^\">]*oracle)|(oracle\.(oracle|pyref))", re.M)
    for base in ("vega_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp", "Makefile")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not bad.search(txt), f"{f} references the oracle"


@pytest.mark.parametrize("n,m", [(0, 1), (0, 4), (1, 1), (1, 4), (10, 32), (15, 4), (9, 2), (100, 7), (101, 101), (4, 4),
                                 (1000, 999), (5, 5), (5, 6), (12345, 64)])
def test_vb_slice_matches_reference_slicing(n, m):
    assert vb.slice_starts(n, m).tolist() == O.slice_starts(n, m).tolist()


def test_vb_slice_closed_form_random():
    rng = np.random.default_rng(0)
    for _ in range(300):
        n = int(rng.integers(0, 3000))
        m = int(rng.integers(1, 400))
        assert vb.slice_starts(n, m).tolist() == O.slice_starts(n, m).tolist(), (n, m)
    assert vb.slice_starts(10 ** 9, 8).tolist() == [i * 125_000_000 for i in range(9)]


def test_vb_slice_rejects_zero_slices():
    with pytest.raises(ValueError):
        vb.slice_starts(10, 0)


def test_partitioner_matches_oracle():
    rng = np.random.default_rng(1)
    l = vb.lib()
    keys = list(rng.integers(0, 2 ** 63, 2000).astype(np.uint64) * np.uint64(2) + np.uint64(1)) + [0, 1, 2 ** 64 - 1, 2 ** 32, 2 ** 32 - 1]
    for k in keys:
        k = int(k)
        for w in (8, 4):
            assert l.vb_hash_key(k, w) == O.lib().vo_hash_key(k, w)
        for R in (1, 2, 3, 8, 64, 1000, 65536):
            assert l.vb_get_partition(k, 8, R) == O.get_partition(k, R)
    # the 8-byte specialisation equals the general MetroHash64_1 on the key's LE bytes
    for k in keys[:50]:
        assert l.vb_hash_key(int(k), 8) == O.metrohash64_1(int(k).to_bytes(8, "little"), 0)
        assert l.vb_hash_key(int(k), 4) == O.metrohash64_1((int(k) & 0xFFFFFFFF).to_bytes(4, "little"), 0)


def test_host_mirror_argument_checks():
    # argument validation of the operator mirror happens before any device work
    from vega_b200.rdd import _Col
    with pytest.raises(TypeError):
        _Col(np.zeros(4, dtype=np.float32))
    with pytest.raises(ValueError):
        _Col(np.zeros((4, 3), dtype=np.uint64), allow_rows=True)
    c = _Col(np.arange(6, dtype=np.int32))
    assert c.key_width == 4 and c.code == 1 and c.n == 6
    c = _Col(np.zeros((5, 2), dtype=np.uint64), allow_rows=True)
    assert c.rows and c.n == 5


def test_python_constants_match_header_enums():
    """vega_b200/_lib.py mirrors the enums of include/vega_b200.h by value."""
    from vega_b200 import _lib as L
    src = open(os.path.join(ROOT, "include", "vega_b200.h")).read()
    vals = {}
    for body in re.findall(r"enum\s+\w+\s*\{([^}]*)\}", src):
        nxt = 0
        for item in body.split(","):
            item = re.sub(r"/\*.*?\*/", "", item, flags=re.S).strip()
            if not item:
                continue
            if "=" in item:
                name, v = [x.strip() for x in item.split("=")]
                nxt = int(v, 0)
            else:
                name = item
            vals[name] = nxt
            nxt += 1
    for name, v in vals.items():
        if hasattr(L, name):
            assert getattr(L, name) == v, name
    for name in ("VB_OK", "VB_ERR_CUDA", "VB_U64", "VB_F64", "VB_AGG_GROUP", "VB_AGG_SORT", "VB_PART_RANGE", "VB_DEVICE_BORROWED", "VB_GEN_UNIQUE"):
        assert name in vals and getattr(L, name) == vals[name]
    # struct layout of vb_stats: 9 u64 + 3 doubles in header order
    assert ctypes.sizeof(L.vb_stats) == 13 * 8      # 10 u64 + 3 doubles, header order
    fields = re.findall(r"\b(uint64_t|double)\s+(\w+);", src[src.index("typedef struct vb_stats"):src.index("} vb_stats;")])
    assert [f for _, f in fields] == [f for f, _ in L.vb_stats._fields_]
