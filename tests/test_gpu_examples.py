"""The three runnable call sequences under examples/ (train.py:95-152, evaluate.py:128-208, predict.py:67-137 of the
reference, on synthetic tensors) run end to end on the GPU and agree with themselves where two routes exist."""
import importlib
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
EX = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples")


@pytest.fixture(scope="module")
def examples(cuda_device):
    sys.path.insert(0, EX)
    yield lambda name: importlib.import_module(name)
    sys.path.remove(EX)


def test_train_step_example(examples):
    r = examples("train_step").run(batch=3)
    assert r["per_image_equals_batched"] and r["fused_equals_two_calls"]
    assert r["arm_positives"] > 100 and 0 < r["odm_positives"] <= r["arm_positives"]
    assert all(r[k] == r[k] and r[k] > 0 for k in ("refine_loss", "det_loss", "clf_loss", "grad_norm_refine"))


def test_evaluate_example(examples):
    r = examples("evaluate_batch").run(batch=2, num_batches=2)
    assert r["fused_equals_call_sequence"] and r["detections"] > 0
    assert 0.0 <= r["mAP_VOC07"] <= 1.0 and 0.0 <= r["mAP_VOC12"] <= 1.0


def test_predict_example(examples):
    r = examples("predict_image").run()
    assert r["gt_boxes"] >= 1 and r["matched_anchors"] > 0 and r["matched_box_rows"] == 25800
    assert sum(r["detections_per_class"].values()) > 0
