"""bench.py's reference arm / CPU baseline: runs the oracle port only, in a process that never loads the product
library; its synthetic inputs are the product's generators byte for byte."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_never_loads_the_product_library():
    code = r"""
import sys
sys.path.insert(0, %r)
import bench
cp = bench.CpuPath()
m = cp.match_encode(bench.host_inputs_match(cp.synth, 0, 2))
d = cp.detect(bench.host_inputs_detect(cp.synth, 0, 1, "normal"))
assert m[1].shape == (2, bench.N_ANCHORS) and len(d[0]) == 10
assert "rodet_b200" not in sys.modules, "the reference arm imported the product package"
maps = open("/proc/self/maps").read()
assert "librodet_b200" not in maps, "the reference arm mapped the product's CUDA library"
assert "librodet_oracle" in maps
print("ok")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stderr[-2000:]


def test_oracle_synth_is_the_product_synth():
    from oracle import synth as osynth
    from rodet_b200 import synth
    shapes = [(4, 4, 6), (2, 2, 9)]
    n = 4 * 4 * 6 + 2 * 2 * 9
    for a, b in ((osynth.gt_batch(7, 3), synth.gt_batch(7, 3)),):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    assert np.array_equal(osynth.head_offsets(3, n, 1), synth.head_offsets(3, n, 1))
    assert np.array_equal(osynth.class_probs(3, n), synth.class_probs(3, n))
    assert np.array_equal(osynth.stress_probs(3, n), synth.stress_probs(3, n))
    for mode in ("quadrant", "bumps"):
        p = synth.clustered_probs(5, shapes, mode)
        assert np.array_equal(osynth.clustered_probs(5, shapes, mode), p)
        assert p.shape == (n, 11) and p.dtype == np.float32 and (p >= 0).all() and (p <= 1).all()
        assert (p[:, 1:] >= 0.3).any()


def test_reference_line_shape():
    """--impl reference prints one JSON line with the contract keys (tiny step counts; rank != 0 prints nothing)."""
    import json
    env = dict(os.environ, RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1",
                        "--workload", "match_encode", "--batch", "4"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["config"]["batch_per_gpu"] == 4 and d["cpu_baseline"]["kind"] == "port" and d["gpu_launches"] == 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["value"] > 0
