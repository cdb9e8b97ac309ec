"""Model file codec (model.bin.gz): libmyrrix_model_io.so against a byte-level
restatement of the Java Object Serialization stream that GenerationSerializer produces
(GenerationSerializer.java:96-276), assembled here independently with `struct`. CPU only.
Parity unpinned: no JVM in the image and no model file in the reference tree."""
import os
import struct

import numpy as np
import pytest

from myrrix_recommender_b200 import model_io as MIO


def _java_stream(payload, block=1024):
    """Stream header + GenerationSerializer class descriptor + payload as block-data records of
    `block` bytes (ObjectOutputStream drains its 1024-byte buffer) + TC_ENDBLOCKDATA."""
    utf = lambda s: struct.pack(">H", len(s)) + s.encode()
    out = b"\xac\xed\x00\x05" + b"\x73" + b"\x72" + utf("net.myrrix.online.generation.GenerationSerializer")
    out += struct.pack(">q", 1) + b"\x03" + struct.pack(">H", 1)
    out += b"L" + utf("generation") + b"\x74" + utf("Lnet/myrrix/online/generation/Generation;")
    out += b"\x78" + b"\x70"
    for a in range(0, len(payload), block):
        chunk = payload[a:a + block]
        out += (b"\x77" + struct.pack(">B", len(chunk))) if len(chunk) <= 255 else (b"\x7a" + struct.pack(">i", len(chunk)))
        out += chunk
    return out + b"\x78"


def _payload(g):
    """writeObject (GenerationSerializer.java:96-106), DataOutput big-endian."""
    p = b""
    if g.known_user_ids is None:
        p += struct.pack(">i", -1)
    else:
        p += struct.pack(">i", len(g.known_user_ids))
        for u, uid in enumerate(g.known_user_ids):
            items = g.known_item_ids[g.known_ptr[u]:g.known_ptr[u + 1]]
            p += struct.pack(">qi", int(uid), len(items)) + b"".join(struct.pack(">q", int(i)) for i in items)
    for ids, m in ((g.user_ids, g.X), (g.item_ids, g.Y)):
        p += struct.pack(">i", len(ids))
        for r, rid in enumerate(ids):
            p += struct.pack(">qi", int(rid), m.shape[1]) + m[r].astype(">f4").tobytes()
    for tags in (g.item_tag_ids, g.user_tag_ids):
        p += struct.pack(">i", len(tags)) + b"".join(struct.pack(">q", int(t)) for t in tags)
    return p + struct.pack(">ii", 0, 0)   # no user / item clusters


def _generation(rng, n_users, n_items, k, known=True):
    g = MIO.Generation(rng.integers(-2**62, 2**62, n_users), rng.standard_normal((n_users, k)).astype(np.float32),
                       rng.integers(-2**62, 2**62, n_items), rng.standard_normal((n_items, k)).astype(np.float32))
    if known:
        cnt = rng.integers(0, 5, n_users)
        g.known_user_ids, g.known_ptr = g.user_ids.copy(), np.concatenate([[0], np.cumsum(cnt)])
        g.known_item_ids = rng.choice(g.item_ids, int(cnt.sum()))
    g.item_tag_ids, g.user_tag_ids = rng.integers(0, 2**62, 3), rng.integers(0, 2**62, 2)
    return g


def _same(a, b):
    for f in ("user_ids", "X", "item_ids", "Y", "item_tag_ids", "user_tag_ids"):
        x, y = getattr(a, f), getattr(b, f)
        assert np.array_equal(x, y) or (x.size == 0 and y.size == 0), f   # (no rows: no feature count)
    assert (a.known_user_ids is None) == (b.known_user_ids is None)
    if a.known_user_ids is not None:
        for f in ("known_user_ids", "known_ptr", "known_item_ids"):
            assert np.array_equal(getattr(a, f), getattr(b, f)), f


def test_symbols():
    lib = MIO.load()
    hdr = open(os.path.join(os.path.dirname(__file__), "..", "include", "myrrix_model_io.h")).read()
    for name, _, _ in MIO.SYMBOLS:
        assert hasattr(lib, name) and name + "(" in hdr


@pytest.mark.parametrize("shape", [(0, 0, 4, True), (1, 1, 2, False), (7, 5, 3, True), (200, 90, 30, True)])
def test_bytes_match_the_java_stream_layout(shape):
    """Byte for byte what ObjectOutputStream would emit for this payload, 1024-byte blocks included."""
    g = _generation(np.random.default_rng(sum(shape[:3])), *shape)
    data = MIO.to_bytes(g)
    assert data == _java_stream(_payload(g))
    _same(MIO.from_bytes(data), g)


def test_reader_accepts_other_block_sizes_and_rejects_garbage():
    g = _generation(np.random.default_rng(4), 40, 20, 6)
    for block in (1, 100, 255, 256, 5000):
        _same(MIO.from_bytes(_java_stream(_payload(g), block)), g)
    good = MIO.to_bytes(g)
    for bad in (good[:-1], good[:200], b"\xac\xed\x00\x05\x73\x72\x00\x03abc" + good[20:], b"\xac\xed"):
        with pytest.raises(IOError):
            MIO.from_bytes(bad)
    with pytest.raises(IOError):      # trailing payload bytes: readObject would leave them unread
        MIO.from_bytes(_java_stream(_payload(g) + b"\x00\x00\x00\x01"))
    clustered = _payload(g)[:-8] + struct.pack(">i", 1) + struct.pack(">iq", 1, 42) + struct.pack(">if", 1, 0.5) + struct.pack(">i", 0)
    _same(MIO.from_bytes(_java_stream(clustered)), g)     # clusters are counted and skipped


def test_non_finite_values_are_refused_both_ways():
    g = _generation(np.random.default_rng(5), 5, 4, 3)
    g.Y[2, 1] = np.inf
    with pytest.raises(ValueError):
        MIO.to_bytes(g)                                    # GenerationSerializer.java:205
    with pytest.raises(IOError):
        MIO.from_bytes(_java_stream(_payload(g)))          # :187


def test_gzip_file_round_trip(tmp_path):
    g = _generation(np.random.default_rng(6), 300, 120, 16)
    path = str(tmp_path / "model.bin.gz")
    MIO.write_generation(g, path)
    _same(MIO.read_generation(path), g)
    with pytest.raises(ValueError):
        MIO.write_generation(g, str(tmp_path / "model.bin"))   # IOUtils.java:276
