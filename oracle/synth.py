"""Synthetic-input generators for the CPU reference arm (bench.py --impl reference) and the oracle-side tests.
TEST / BENCH INFRASTRUCTURE ONLY.

The generators live in the product package (`road-object-detection-for-bdd100k_b200/synth.py`, NumPy only).
The reference arm must time the CPU path in a process that never imports the product package (importing it
dlopens the CUDA library), so this module executes that one source file by path, without its package."""
import importlib.util
import os

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                     "road-object-detection-for-bdd100k_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("_rodet_synth_standalone", _PATH)
_mod = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mod)

gt_boxes, gt_batch, head_offsets = _mod.gt_boxes, _mod.gt_batch, _mod.head_offsets
class_logits, class_probs, stress_probs, clustered_probs = _mod.class_logits, _mod.class_probs, _mod.stress_probs, _mod.clustered_probs
BASE_SEED, MAX_GT = _mod.BASE_SEED, _mod.MAX_GT
