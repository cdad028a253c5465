"""bincode 1.2.1 (default options) restated for the blob types on the shuffle path — TEST INFRASTRUCTURE ONLY.

vega serialises each (map, reduce) bucket with `bincode::serialize(&Vec<(K,C)>)` (src/dependency.rs:213-214,
Cargo.toml:44 `bincode = "1.2.1"`) and decodes it in the fetcher (src/shuffle/shuffle_fetcher.rs:85).
bincode is a third-party crate absent from /root/reference; its published 1.x default format is:
little-endian, fixed-width integers, `Vec<T>`/`String` = u64 length then the elements/bytes, tuples =
fields in order.  The reference holds no golden bytes for it — only a round trip of
`vec![(0i32, "example data")]` (src/shuffle/shuffle_fetcher.rs:143-166) — so byte-level parity is
anchored on that published format ("parity unpinned" by reference bytes) plus that round trip.
"""
import struct


def encode_pairs(pairs):
    """Vec<(u64,u64)>"""
    out = [struct.pack("<Q", len(pairs))]
    for k, c in pairs:
        out.append(struct.pack("<QQ", k & (2 ** 64 - 1), c & (2 ** 64 - 1)))
    return b"".join(out)


def decode_pairs(blob):
    (n,) = struct.unpack_from("<Q", blob, 0)
    if len(blob) != 8 + 16 * n:
        raise ValueError("corrupted blob")
    return [struct.unpack_from("<QQ", blob, 8 + 16 * i) for i in range(n)]


def encode_groups(groups):
    """Vec<(u64,Vec<u64>)>"""
    out = [struct.pack("<Q", len(groups))]
    for k, vs in groups:
        out.append(struct.pack("<QQ", k & (2 ** 64 - 1), len(vs)))
        out.append(struct.pack(f"<{len(vs)}Q", *[v & (2 ** 64 - 1) for v in vs]))
    return b"".join(out)


def decode_groups(blob):
    (n,) = struct.unpack_from("<Q", blob, 0)
    pos, out = 8, []
    for _ in range(n):
        k, ln = struct.unpack_from("<QQ", blob, pos)
        pos += 16
        out.append((k, list(struct.unpack_from(f"<{ln}Q", blob, pos))))
        pos += 8 * ln
    if pos != len(blob):
        raise ValueError("corrupted blob")
    return out


def encode_i32_string_vec(items):
    """Vec<(i32,String)> — the type of the reference's only bincode fixture (shuffle_fetcher.rs:155)."""
    out = [struct.pack("<Q", len(items))]
    for i, s in items:
        b = s.encode()
        out.append(struct.pack("<iQ", i, len(b)) + b)
    return b"".join(out)


def decode_i32_string_vec(blob):
    (n,) = struct.unpack_from("<Q", blob, 0)
    pos, out = 8, []
    for _ in range(n):
        i, ln = struct.unpack_from("<iQ", blob, pos)
        pos += 12
        out.append((i, blob[pos:pos + ln].decode()))
        pos += ln
    if pos != len(blob):
        raise ValueError("corrupted blob")
    return out
