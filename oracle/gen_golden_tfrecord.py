#!/usr/bin/env python
"""Writes tests/golden/voc_gt_000.tfrecord + voc_gt_expected.npz.  TEST INFRASTRUCTURE ONLY.

The TFRecord file is produced by the UNMODIFIED reference converter, dataset/pascalvoc_to_tfrecords.py::run
(XML annotations -> _process_image -> _convert_to_example -> tf.python_io.TFRecordWriter, :60-230), executed over the
NumPy `tensorflow` shim whose tf.train.* messages are real protobuf messages (oracle/tf_shim/example_proto.py).  The
expected arrays are derived independently from the synthetic annotations: pixel box / image size in Python floats,
rounded to float32 by the FloatList encoding.  Run in the build container (needs /root/reference)."""
import os
import shutil
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

NAMES = ['bus', 'traffic light', 'traffic sign', 'person', 'bike', 'truck', 'motor', 'car', 'train', 'rider']


def main():
    ref_loader.load_reference()
    from dataset import pascalvoc_to_tfrecords as P          # /root/reference/dataset/pascalvoc_to_tfrecords.py
    rng = np.random.default_rng(20261018)
    src, dst = tempfile.mkdtemp(), tempfile.mkdtemp()
    os.makedirs(os.path.join(src, "Annotations"))
    os.makedirs(os.path.join(src, "JPEGImages"))
    n_img = 12
    exp = {"ymin": [], "xmin": [], "ymax": [], "xmax": [], "label": [], "difficult": [], "truncated": [], "offsets": [0], "shape": []}
    for i in range(n_img):
        name = "img_%03d" % i
        h, w = int(rng.integers(300, 900)), int(rng.integers(400, 1400))
        g = 0 if i == 5 else int(rng.integers(1, 40))         # one image without objects
        objs = []
        for _ in range(g):
            y0, x0 = int(rng.integers(0, h - 2)), int(rng.integers(0, w - 2))
            y1, x1 = int(rng.integers(y0 + 1, h)), int(rng.integers(x0 + 1, w))
            lab = NAMES[int(rng.integers(0, len(NAMES)))]
            objs.append((lab, y0, x0, y1, x1))
            exp["ymin"].append(y0 / h); exp["xmin"].append(x0 / w); exp["ymax"].append(y1 / h); exp["xmax"].append(x1 / w)
            exp["label"].append(P.LABELS[lab][0])
            # `if obj.find('difficult'):` is False for a childless element (:107-114): the converter always writes 0
            exp["difficult"].append(0); exp["truncated"].append(0)
        exp["offsets"].append(len(exp["label"]))
        exp["shape"].append([h, w, 3])
        xml = ["<annotation><size><height>%d</height><width>%d</width><depth>3</depth></size>" % (h, w)]
        for lab, y0, x0, y1, x1 in objs:
            xml.append("<object><name>%s</name><difficult>1</difficult><truncated>1</truncated><bndbox><ymin>%d</ymin><xmin>%d</xmin>"
                       "<ymax>%d</ymax><xmax>%d</xmax></bndbox></object>" % (lab, y0, x0, y1, x1))
        xml.append("</annotation>")
        with open(os.path.join(src, "Annotations", name + ".xml"), "w") as f:
            f.write("".join(xml))
        with open(os.path.join(src, "JPEGImages", name + ".jpg"), "wb") as f:
            f.write(bytes(rng.integers(0, 256, size=int(rng.integers(20, 200)), dtype=np.uint8)))   # stands in for the JPEG
    P.run(src + "/", dst, name="bdd100k_train", shuffling=False)
    out = os.path.join(ROOT, "tests", "golden")
    shutil.copy(os.path.join(dst, "bdd100k_train_000.tfrecord"), os.path.join(out, "voc_gt_000.tfrecord"))
    np.savez_compressed(
        os.path.join(out, "voc_gt_expected.npz"),
        ymin=np.asarray(exp["ymin"], np.float64).astype(np.float32), xmin=np.asarray(exp["xmin"], np.float64).astype(np.float32),
        ymax=np.asarray(exp["ymax"], np.float64).astype(np.float32), xmax=np.asarray(exp["xmax"], np.float64).astype(np.float32),
        label=np.asarray(exp["label"], np.int64), difficult=np.asarray(exp["difficult"], np.int64),
        truncated=np.asarray(exp["truncated"], np.int64), offsets=np.asarray(exp["offsets"], np.int64),
        shape=np.asarray(exp["shape"], np.int64))
    print("\nwrote", os.path.getsize(os.path.join(out, "voc_gt_000.tfrecord")), "bytes,", n_img, "records,", len(exp["label"]), "objects")
    shutil.rmtree(src); shutil.rmtree(dst)


if __name__ == "__main__":
    main()
