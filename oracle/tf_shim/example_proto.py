"""tf.train.Example / Features / Feature / *List and tf.python_io.TFRecordWriter for the NumPy `tensorflow` shim.
TEST INFRASTRUCTURE ONLY.

The messages are built with the real `protobuf` runtime from the published schema of
tensorflow/core/example/{feature,example}.proto, so `SerializeToString()` yields genuine protobuf wire format;
the record framing follows tensorflow/core/lib/io/record_writer.cc:
    uint64 length | uint32 masked_crc32c(length) | bytes data | uint32 masked_crc32c(data)      (little endian)
    masked(crc) = ((crc >> 15) | (crc << 17)) + 0xa282ead8   (mod 2^32),  crc = CRC-32C (Castagnoli)."""
import os
import struct
import types

from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto


def _build_messages():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "rodet_oracle/tf_example.proto"
    fd.package = "rodet_oracle_tf"
    fd.syntax = "proto3"

    def lst(name, ftype, packed):
        m = fd.message_type.add()
        m.name = name
        f = m.field.add()
        f.name, f.number, f.label, f.type = "value", 1, _T.LABEL_REPEATED, ftype
        if packed:
            f.options.packed = True
    lst("BytesList", _T.TYPE_BYTES, False)
    lst("FloatList", _T.TYPE_FLOAT, True)
    lst("Int64List", _T.TYPE_INT64, True)

    feat = fd.message_type.add()
    feat.name = "Feature"
    feat.oneof_decl.add().name = "kind"
    for i, (n, t) in enumerate((("bytes_list", "BytesList"), ("float_list", "FloatList"), ("int64_list", "Int64List"))):
        f = feat.field.add()
        f.name, f.number, f.label, f.type = n, i + 1, _T.LABEL_OPTIONAL, _T.TYPE_MESSAGE
        f.type_name = ".rodet_oracle_tf." + t
        f.oneof_index = 0

    feats = fd.message_type.add()
    feats.name = "Features"
    entry = feats.nested_type.add()
    entry.name = "FeatureEntry"
    entry.options.map_entry = True
    k = entry.field.add()
    k.name, k.number, k.label, k.type = "key", 1, _T.LABEL_OPTIONAL, _T.TYPE_STRING
    v = entry.field.add()
    v.name, v.number, v.label, v.type = "value", 2, _T.LABEL_OPTIONAL, _T.TYPE_MESSAGE
    v.type_name = ".rodet_oracle_tf.Feature"
    f = feats.field.add()
    f.name, f.number, f.label, f.type = "feature", 1, _T.LABEL_REPEATED, _T.TYPE_MESSAGE
    f.type_name = ".rodet_oracle_tf.Features.FeatureEntry"

    ex = fd.message_type.add()
    ex.name = "Example"
    f = ex.field.add()
    f.name, f.number, f.label, f.type = "features", 1, _T.LABEL_OPTIONAL, _T.TYPE_MESSAGE
    f.type_name = ".rodet_oracle_tf.Features"

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("rodet_oracle_tf." + n))
    return {n: get(n) for n in ("BytesList", "FloatList", "Int64List", "Feature", "Features", "Example")}


_CRC_TABLE = None


def crc32c(data: bytes) -> int:
    global _CRC_TABLE
    if _CRC_TABLE is None:
        t = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            t.append(c)
        _CRC_TABLE = t
    c = 0xFFFFFFFF
    for b in data:
        c = _CRC_TABLE[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc32c(data: bytes) -> int:
    c = crc32c(data)
    return ((((c >> 15) | (c << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


class TFRecordWriter:
    def __init__(self, path, options=None):
        self._f = open(path, "wb")

    def write(self, record: bytes):
        head = struct.pack("<Q", len(record))
        self._f.write(head + struct.pack("<I", masked_crc32c(head)) + record + struct.pack("<I", masked_crc32c(record)))

    def flush(self):
        self._f.flush()

    def close(self):
        self._f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


class _GFile:
    def __init__(self, name, mode="r"):
        self._f = open(name, mode)

    def read(self, n=-1):
        return self._f.read(n)

    def close(self):
        self._f.close()


def attach(tf):
    """Adds tf.train.{Example,...}, tf.python_io.TFRecordWriter and the few tf.gfile calls of the dataset converter."""
    msgs = _build_messages()
    train = types.ModuleType("tensorflow.train")
    for n, cls in msgs.items():
        setattr(train, n, cls)
    tf.train = train
    pio = types.ModuleType("tensorflow.python_io")
    pio.TFRecordWriter = TFRecordWriter
    tf.python_io = pio
    gfile = types.ModuleType("tensorflow.gfile")
    gfile.FastGFile = _GFile
    gfile.GFile = _GFile
    gfile.Exists = os.path.exists
    gfile.MakeDirs = lambda p: os.makedirs(p, exist_ok=True)
    tf.gfile = gfile
    return tf
