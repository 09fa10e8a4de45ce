"""f-4: the native TFRecord ground-truth reader (host code, no GPU needed) against the fixture written by the
UNMODIFIED reference converter (dataset/pascalvoc_to_tfrecords.py::run over the shim, oracle/gen_golden_tfrecord.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden

FIXTURE = os.path.join(GOLDEN, "voc_gt_000.tfrecord")


def _read(*a, **k):
    from rodet_b200.dataset.pascalvoc_common import read_ground_truth
    return read_ground_truth(*a, **k)


def test_reader_matches_reference_writer():
    z = golden("voc_gt_expected.npz")
    gt = _read(FIXTURE)
    assert len(gt) == z["offsets"].size - 1 == 12
    assert np.array_equal(gt.offsets, z["offsets"]) and np.array_equal(gt.shape, z["shape"])
    for k in ("ymin", "xmin", "ymax", "xmax"):
        assert np.array_equal(getattr(gt, k).view(np.uint32), z[k].view(np.uint32)), k        # bit-exact float32
    for k in ("label", "difficult", "truncated"):
        assert np.array_equal(getattr(gt, k), z[k]), k
    assert gt.max_objects == int(np.diff(z["offsets"]).max())
    assert (np.diff(z["offsets"]) == 0).any()                          # the fixture holds an image without objects


def test_files_concatenate_and_bytes_input():
    data = open(FIXTURE, "rb").read()
    one, two = _read(data), _read([FIXTURE, data])
    assert len(two) == 2 * len(one)
    assert np.array_equal(two.label, np.concatenate([one.label, one.label]))
    assert np.array_equal(two.offsets, np.concatenate([one.offsets, one.offsets[1:] + one.offsets[-1]]))
    assert len(_read(b"")) == 0


def test_corruption_is_detected():
    data = bytearray(open(FIXTURE, "rb").read())
    bad = bytearray(data)
    bad[40] ^= 0x5A                                                     # inside the first record's payload
    with pytest.raises(ValueError, match="corrupted data in record 0"):
        _read(bytes(bad))
    bad = bytearray(data)
    bad[3] ^= 0x01                                                      # the length field
    with pytest.raises(ValueError, match="corrupted length|truncated"):
        _read(bytes(bad))
    with pytest.raises(ValueError, match="truncated"):
        _read(bytes(data[:-3]))
    # without CRC verification a flipped payload bit inside a float is accepted (the caller's choice)
    assert len(_read(bytes(data), verify_crc=False)) == 12


def test_unpacked_encoding_is_accepted():
    """proto2-style unpacked repeated fields (one tag per value) must parse like the packed form TF writes."""
    import struct

    def varint(v):
        out = bytearray()
        while True:
            b = v & 0x7F
            v >>= 7
            out.append(b | (0x80 if v else 0))
            if not v:
                return bytes(out)

    def ld(field, payload):
        return varint((field << 3) | 2) + varint(len(payload)) + payload
    floats = lambda vs: b"".join(varint((1 << 3) | 5) + struct.pack("<f", v) for v in vs)
    ints = lambda vs: b"".join(varint((1 << 3) | 0) + varint(v) for v in vs)
    feat = lambda key, kind, payload: ld(1, ld(1, key.encode()) + ld(2, ld(kind, payload)))
    ex = ld(1, feat("image/object/bbox/ymin", 2, floats([0.25, 0.5])) + feat("image/object/bbox/xmin", 2, floats([0.125, 0.0])) +
            feat("image/object/bbox/ymax", 2, floats([0.75, 1.0])) + feat("image/object/bbox/xmax", 2, floats([0.5, 0.25])) +
            feat("image/object/bbox/label", 3, ints([8, 300])) + feat("image/shape", 3, ints([720, 1280, 3])))
    from oracle.tf_shim.example_proto import masked_crc32c
    head = struct.pack("<Q", len(ex))
    rec = head + struct.pack("<I", masked_crc32c(head)) + ex + struct.pack("<I", masked_crc32c(ex))
    gt = _read(rec)
    assert len(gt) == 1 and gt.label.tolist() == [8, 300] and gt.ymax.tolist() == [0.75, 1.0]
    assert gt.difficult.tolist() == [0, 0] and gt.shape.tolist() == [[720, 1280, 3]]
    # inconsistent list lengths are rejected
    bad = ld(1, feat("image/object/bbox/ymin", 2, floats([0.25])) + feat("image/object/bbox/label", 3, ints([8, 9])))
    head = struct.pack("<Q", len(bad))
    with pytest.raises(ValueError, match="differ in length"):
        _read(head + struct.pack("<I", masked_crc32c(head)) + bad + struct.pack("<I", masked_crc32c(bad)))


def test_read_gt_rejects_undersized_arrays():
    """rod_tfrecord_read_gt called with fewer objects than the file holds must fail before writing past the arrays
    (the optional difficult / truncated arrays included)."""
    import ctypes
    from rodet_b200 import _abi
    data = np.frombuffer(open(FIXTURE, "rb").read(), dtype=np.uint8)
    z = golden("voc_gt_expected.npz")
    R, O = z["offsets"].size - 1, int(z["offsets"][-1])
    f = lambda n: np.zeros(n, dtype=np.float32)
    i = lambda n: np.full(n + 8, -7, dtype=np.int64)                    # guard words behind the announced size
    for announced in (0, 1, O - 1):
        ymin, xmin, ymax, xmax, label, diff, trunc = f(announced + 8), f(announced + 8), f(announced + 8), f(announced + 8), \
            i(announced), i(announced), i(announced)
        offsets = np.zeros(R + 1, dtype=np.int64)
        rc = _abi.lib.rod_tfrecord_read_gt(data.ctypes.data, data.size, 0, R, announced, ymin.ctypes.data, xmin.ctypes.data,
                                           ymax.ctypes.data, xmax.ctypes.data, label.ctypes.data, diff.ctypes.data,
                                           trunc.ctypes.data, offsets.ctypes.data, None)
        assert rc != 0
        for a in (label, diff, trunc):
            assert (a[announced:] == -7).all()
        with pytest.raises(ValueError):
            _abi.check(rc)


def test_mutated_records_never_crash_the_parser():
    """The reader takes bytes from disk: with the CRC check off, random damage to the protobuf payload (bit flips,
    truncation, inserted and deleted bytes, oversized varints / lengths) must end in a parsed result or a ValueError —
    never in an out-of-bounds access (arrays are sized by the index pass and guarded again by the fill pass)."""
    import struct
    rng = np.random.default_rng(2024)
    data = open(FIXTURE, "rb").read()
    (n0,) = struct.unpack("<Q", data[:8])
    first = data[12:12 + n0]                                            # payload of record 0
    from oracle.tf_shim.example_proto import masked_crc32c

    def frame(payload):
        head = struct.pack("<Q", len(payload))
        return head + struct.pack("<I", masked_crc32c(head)) + payload + struct.pack("<I", masked_crc32c(payload))

    outcomes = {"ok": 0, "rejected": 0}
    for trial in range(400):
        b = bytearray(first)
        kind = trial % 5
        if kind == 0:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        elif kind == 1:
            del b[int(rng.integers(1, len(b))):]
        elif kind == 2:
            at = int(rng.integers(0, len(b)))
            b[at:at] = bytes(rng.integers(0, 256, int(rng.integers(1, 12)), dtype=np.uint8))
        elif kind == 3:
            at = int(rng.integers(0, len(b) - 1))
            del b[at:at + int(rng.integers(1, 8))]
        else:
            at = int(rng.integers(0, len(b)))
            b[at:at + 1] = b"\xff" * 10 + b"\x7f"                      # an 11-byte varint / absurd length
        try:
            gt = _read(frame(bytes(b)) + data, verify_crc=False)         # damaged record in front of the intact file
            assert len(gt) == 13 and gt.offsets[-1] == gt.label.size == gt.ymin.size
            outcomes["ok"] += 1
        except ValueError:
            outcomes["rejected"] += 1
    assert outcomes["rejected"] > 50 and outcomes["ok"] > 20, outcomes
