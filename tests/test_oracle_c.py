"""The C oracle port (multi-threaded CPU baseline) against the NumPy restatement: bit-exact."""
import os

import numpy as np
import pytest

from conftest import golden_anchors
from helpers import bit_equal
from oracle import c_port as C
from oracle import restated as R
from rodet_b200 import synth


@pytest.fixture(scope="module")
def table():
    return R.AnchorTable(golden_anchors("418"))


def test_c_oracle_targets(table):
    B = 3
    corner, labels, counts = synth.gt_batch(40, B)
    center = R.corner_to_center(corner).astype(np.float32)
    gt, cb, lab, pos, idx = C.arm_match_encode(table, center, labels, counts, R.REFINE_POS_JAC)
    ro = np.stack([synth.head_offsets(40 + b, table.n) for b in range(B)])
    ro = np.where((np.arange(table.n)[None, :, None] % 2 == 0) & (pos[..., None] > 0), gt + 0.05 * ro, ro).astype(np.float32)
    for b in range(B):
        o = R.arm_match_encode(table, center[b, :counts[b]], labels[b, :counts[b]])
        assert np.array_equal(pos[b], o[3]) and np.array_equal(idx[b], o[4]) and np.array_equal(lab[b], o[2])
        assert bit_equal(gt[b], o[0]) and bit_equal(cb[b], o[1])
    d = C.odm_target(table, ro, gt, cb, lab, pos, R.DET_POS_JAC)
    o = R.odm_target(table, ro, gt, cb, lab, pos)
    assert np.array_equal(d[1], o[1]) and np.array_equal(d[2], o[2])
    assert bit_equal(d[0], o[0]) and bit_equal(d[3], o[3])
    assert 0 < int(o[1].sum()) < int(pos.sum())


@pytest.mark.parametrize("stress", [False, True])
def test_c_oracle_detect(table, stress):
    B = 2
    mk = synth.stress_probs if stress else synth.class_probs
    probs = np.stack([mk(50 + b, table.n) for b in range(B)])
    if stress:
        probs = (np.round(probs * 64) / 64).astype(np.float32)        # ties
    ro = np.stack([synth.head_offsets(50 + b, table.n) for b in range(B)])
    do = np.stack([synth.head_offsets(50 + b, table.n, 1) for b in range(B)])
    boxes = C.decode_corner(table, ro, do)
    assert bit_equal(boxes, R.decode_corner(table, ro, do))
    s, bx = C.detected_bboxes(probs, boxes, 0.3, 0.45, 400, 200)
    o_s, o_b = R.detected_bboxes(probs, boxes, 0.3, 0.45, None, 400, 200)
    for c in range(1, 11):
        assert bit_equal(s[c], o_s[c]) and bit_equal(bx[c], o_b[c]), c
    assert C.threads() >= 1


def test_bench_reference_arm_runs_on_cpu():
    """`bench.py --impl reference` (the CPU arm the driver times) needs no GPU: one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "images/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
