"""AddressSanitizer fuzz of the native TFRecord ground-truth reader (host code of csrc/tfrecord.cu).

Not collected by pytest (needs a sanitizer build):
  nvcc -gencode arch=compute_100a,code=sm_100a -O1 -g -std=c++17 -Xcompiler -fPIC -Xcompiler -fsanitize=address \\
       -Xcompiler -fno-omit-frame-pointer -Iinclude -shared -o /tmp/asan/libtf_asan.so \\
       road-object-detection-for-bdd100k_b200/csrc/tfrecord.cu road-object-detection-for-bdd100k_b200/csrc/common.cu -cudart shared -lasan
  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 python tests/tools/asan_fuzz_tfrecord.py
3 000 damaged records (bit flips, truncation, insertions, deletions, oversized varints, random bytes) in exact-size heap buffers:
every one is either parsed or rejected with an error code; round 2: 307 parsed, 2 693 rejected, no sanitizer report."""
import ctypes, struct, sys
import numpy as np
import os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.tf_shim.example_proto import masked_crc32c
lib = ctypes.CDLL("/tmp/asan/libtf_asan.so")
lib.rod_tfrecord_index.restype = ctypes.c_int
lib.rod_tfrecord_index.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
lib.rod_tfrecord_read_gt.restype = ctypes.c_int
lib.rod_tfrecord_read_gt.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int64, ctypes.c_int64] + [ctypes.c_void_p] * 9
data = open(os.path.join(ROOT, "tests", "golden", "voc_gt_000.tfrecord"), "rb").read()
(n0,) = struct.unpack("<Q", data[:8])
first = data[12:12 + n0]
def frame(p):
    h = struct.pack("<Q", len(p)); return h + struct.pack("<I", masked_crc32c(h)) + p + struct.pack("<I", masked_crc32c(p))
rng = np.random.default_rng(7)
ok = rej = 0
for trial in range(3000):
    b = bytearray(first); kind = trial % 6
    if kind == 0:
        for _ in range(int(rng.integers(1, 6))): b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
    elif kind == 1: del b[int(rng.integers(1, len(b))):]
    elif kind == 2:
        at = int(rng.integers(0, len(b))); b[at:at] = bytes(rng.integers(0, 256, int(rng.integers(1, 12)), dtype=np.uint8))
    elif kind == 3:
        at = int(rng.integers(0, len(b) - 1)); del b[at:at + int(rng.integers(1, 8))]
    elif kind == 4:
        at = int(rng.integers(0, len(b))); b[at:at + 1] = b"\xff" * 10 + b"\x7f"
    else:
        b = bytearray(rng.integers(0, 256, int(rng.integers(0, 200)), dtype=np.uint8).tobytes())
    buf = np.frombuffer(frame(bytes(b)) + data, dtype=np.uint8).copy()      # exact-size heap buffer: ASan sees any overrun
    nr, no = ctypes.c_int64(0), ctypes.c_int64(0)
    rc = lib.rod_tfrecord_index(buf.ctypes.data, buf.size, 0, ctypes.byref(nr), ctypes.byref(no))
    if rc: rej += 1; continue
    R, O = nr.value, no.value
    f = [np.zeros(O, np.float32) for _ in range(4)]; i = [np.zeros(O, np.int64) for _ in range(3)]
    off = np.zeros(R + 1, np.int64); shp = np.zeros((R, 3), np.int64)
    p = lambda a: a.ctypes.data if a.size else None
    rc = lib.rod_tfrecord_read_gt(buf.ctypes.data, buf.size, 0, R, O, *[p(a) for a in f], *[p(a) for a in i], off.ctypes.data, p(shp))
    assert rc == 0, "index accepted what read_gt rejects"
    ok += 1
print("ok", ok, "rejected", rej)
