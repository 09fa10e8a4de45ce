#!/usr/bin/env python
"""bench.py — images/s of the box-level hot path on B200 (BASELINE.json metric: match+encode; decode+NMS).

A "step" is one pass of the hot path over one batch of synthetic BDD100K-shaped input (512x512 anchor layout,
N = 36 852 anchors, 10 classes, <= 100 GT boxes per image):

  match_encode (BASELINE configs[1], primary): training-target generation, ARM matching + encoding followed by ODM
               target generation (reference: refine_groundtruth -> det_groundtruth, train.py:109-113 -> :147-149) as
               ONE fused kernel through `net_tools.target_gen`, batch 32 per GPU; the two-call path is timed beside it.
  decode_nms   (configs[2]): decode + select 0.3 + top-k 400 + NMS 0.45 through `net_tools.decode_detected_bboxes`,
               batch 64 per GPU; nms_stress (configs[4]): every (anchor, class) above threshold;
               decode_nms_clustered: spatially clustered candidates, with the rate of segments that fell back to the
               exact general kernels.
The default run prints ONE compact JSON line: the contract keys describe match_encode; decode_nms, nms_stress and the
clustered workload appear as flat `decode_nms_*` / `nms_stress_*` / `decode_nms_clustered_*` scalars.
`--workload decode_nms` (or nms_stress) makes that workload the primary line instead.

  value        whole-job images/s, inputs resident in HBM, ONE batch in flight (strictly serial steps on one stream);
               `value_overlapped` = consecutive independent batches alternating over `--streams` streams.
  timing       K = --steps steps are captured into one CUDA graph (one host launch per K steps) = one timed loop
               between two CUDA events; loops repeat until >= 50 ms were timed; ms_per_step = median loop / K,
               max over ranks; `loop_spread` = (min, max) of the per-rank medians and of the loops.
  roofline     dominant kernel's algorithmic bytes / its launch duration vs MEASURED_PEAKS.json's HBM copy rate; the
               FP32 non-FMA issue bound of the pair loop is reported next to it (`bound` names the binding one).
  e2e          the same metric through the public API from pinned HOST buffers: H2D of every input and D2H of the
               complete result inside the timed region.
  cpu_baseline the C/OpenMP oracle port (oracle/c) on the host cores, bounded sample (the reference's TF-1 CPU path
               cannot run: TensorFlow is not installable here).
  --impl reference   times only that CPU port, in a process that never imports the product package.
  --global-batch G   strong scaling: G images per step split over the ranks (BASELINE configs[3]: 256).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_CLASSES = 11
IMG, FEATS = (512, 512), [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
SHAPES = [(fh, fw, 6 if i == 0 else 9) for i, (fh, fw) in enumerate(FEATS)]
N_ANCHORS = sum(a * b * c for a, b, c in SHAPES)
SELECT_THR, NMS_THR, TOP_K, KEEP = 0.3, 0.45, 400, 200
FALLBACK_HBM_GBS = 6650.0
MIN_TIMED_MS = 50.0


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=100)
    p.add_argument("--warmup", type=int, default=10)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="all", choices=["all", "match_encode", "decode_nms", "nms_stress"])
    p.add_argument("--batch", type=int, default=32, help="images per GPU per step, match_encode")
    p.add_argument("--batch-detect", type=int, default=64, help="images per GPU per step, decode_nms")
    p.add_argument("--global-batch", type=int, default=0, help="strong scaling: images per step over ALL ranks")
    p.add_argument("--no-graphs", action="store_true", help="launch through Python every step instead of CUDA graphs")
    p.add_argument("--streams", type=int, default=4, help="streams of the `value_overlapped` variant")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--skip-extras", action="store_true", help="only the primary workload (no flat secondary scalars)")
    p.add_argument("--cpu-seconds", type=float, default=10.0)
    return p.parse_args()


# ----------------------------------------------------------------------------------------- helpers
def measured_traffic(key):
    """DRAM bytes per launch / step from the committed ncu capture (profiles/r02_traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            return json.load(f)[key]["bytes"]
    except Exception:
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback 6.65 TB/s (B200_PROFILING.md)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        busy = [x for x in sm if mx and x >= 0.5 * mx] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def split_np(flat, shapes, tail):
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        out.append(np.ascontiguousarray(flat[:, off:off + n]).reshape((flat.shape[0], fh, fw, a) + tail))
        off += n
    return out


def corner_to_center_np(cr):
    """cornerBboxes_2_centerBboxes (utils/common_tools.py:51-54) on the host, float32 op by op."""
    cr = np.asarray(cr, dtype=np.float32)
    return np.stack([(cr[..., 0] + cr[..., 2]) / np.float32(2), (cr[..., 1] + cr[..., 3]) / np.float32(2),
                     cr[..., 2] - cr[..., 0], cr[..., 3] - cr[..., 1]], -1)


def host_inputs_match(synth, first_image, B):
    """Host inputs of one match_encode step: GT in centre form (padded), labels, counts, ARM head output."""
    corner, labels, counts = synth.gt_batch(first_image, B)
    center = corner_to_center_np(corner)
    for b in range(B):
        center[b, counts[b]:] = 0
    ro = np.stack([synth.head_offsets(first_image + b, N_ANCHORS) for b in range(B)])
    return center.astype(np.float32), labels, counts, ro


def host_inputs_detect(synth, first_image, B, kind):
    """kind: normal | stress | quadrant | bumps."""
    if kind == "stress":
        probs = np.stack([synth.stress_probs(first_image + b, N_ANCHORS) for b in range(B)])
    elif kind == "normal":
        probs = np.stack([synth.class_probs(first_image + b, N_ANCHORS) for b in range(B)])
    else:
        probs = np.stack([synth.clustered_probs(first_image + b, SHAPES, kind) for b in range(B)])
    s = (0.05, 0.05) if kind == "stress" else (0.1, 0.2)
    ro = np.stack([synth.head_offsets(first_image + b, N_ANCHORS, 0, *s) for b in range(B)])
    do = np.stack([synth.head_offsets(first_image + b, N_ANCHORS, 1, *s) for b in range(B)])
    return probs, ro, do


# ----------------------------------------------------------------------------------------- CPU side (oracle port)
# The reference's own TF-1 CPU path cannot run (TensorFlow is not installable here), so the CPU side is the oracle
# port: oracle/c (plain C, OpenMP over the host threads), pinned bit-for-bit to oracle/restated.py and through it to
# the fixtures produced by the unmodified reference.  Nothing below imports the product package.
class CpuPath:
    def __init__(self):
        from oracle import c_port, restated, synth
        self.C, self.R, self.synth = c_port, restated, synth
        feats = {"layer_%d" % (i + 1): f for i, f in enumerate(FEATS)}
        self.table = restated.AnchorTable(restated.anchors_all_layer(IMG, feats, restated.init_anchor(len(FEATS), IMG)))
        assert self.table.n == N_ANCHORS

    def match_encode(self, inp):
        c, l, k, ro = inp
        g = self.C.arm_match_encode(self.table, c, l, k, self.R.REFINE_POS_JAC)
        return self.C.odm_target(self.table, ro, g[0], g[1], g[2], g[3], self.R.DET_POS_JAC)

    def detect(self, inp):
        p, ro, do = inp
        return self.C.detected_bboxes(p, self.C.decode_corner(self.table, ro, do), SELECT_THR, NMS_THR, TOP_K, KEEP)

    def time_images(self, fn, make, per, seconds, max_images=4096, n_inputs=3):
        """images/s of fn over batches of `per` images (a few prebuilt batches, rotated) until `seconds` of CPU
        time were spent."""
        inputs = [make(i) for i in range(n_inputs)]
        fn(inputs[0])                                          # warm (page faults, OpenMP pool)
        n_img, t_total, i = 0, 0.0, 1
        while t_total < seconds and n_img < max_images:
            inp = inputs[i % n_inputs]
            t0 = time.perf_counter(); fn(inp); t_total += time.perf_counter() - t0
            n_img += per; i += 1
        return n_img / t_total, n_img


def cpu_baselines(args, B_m, B_d):
    """Rank 0, N = 1: bounded samples of the same workloads on the host cores, plus BASELINE configs[0]
    (decode + NMS of ONE image) on one thread and on all threads."""
    cp = CpuPath()
    threads = os.cpu_count() or 1
    cp.C.set_threads(threads)
    t = cp.C.threads()
    sec = args.cpu_seconds
    v_m, n_m = cp.time_images(cp.match_encode, lambda i: host_inputs_match(cp.synth, 900_000 + i * B_m, B_m), B_m, sec * 0.4)
    v_d, n_d = cp.time_images(cp.detect, lambda i: host_inputs_detect(cp.synth, 910_000 + i * B_d, B_d, "normal"), B_d, sec * 0.3)
    one = lambda i: host_inputs_detect(cp.synth, 920_000 + i, 1, "normal")
    v_1a, _ = cp.time_images(cp.detect, one, 1, sec * 0.1, 200)
    cp.C.set_threads(1)
    v_11, n_11 = cp.time_images(cp.detect, one, 1, sec * 0.1, 200)
    v_m1, _ = cp.time_images(cp.match_encode, lambda i: host_inputs_match(cp.synth, 930_000 + i, 1), 1, sec * 0.1, 200)
    cp.C.set_threads(threads)
    note = "C/OpenMP port (oracle/c); TF-1 reference cannot run here"
    return {
        "match_encode": {"value": v_m, "unit": "images/s", "cores": t, "kind": "port",
                         "sample": "%d images in batches of %d on %d threads; %s" % (n_m, B_m, t, note)},
        "decode_nms": {"value": v_d, "unit": "images/s", "cores": t, "kind": "port",
                       "sample": "%d images in batches of %d on %d threads; %s" % (n_d, B_d, t, note)},
        "config0": {"decode_nms_1image_ms_1thread": 1e3 / v_11, "decode_nms_1image_ms_all_threads": 1e3 / v_1a,
                    "match_encode_1image_ms_1thread": 1e3 / v_m1, "images": n_11},
    }


def run_reference(args):
    """Reference arm: the CPU port on all host threads, rank 0 only, the GPU arm's images per step."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cp = CpuPath()
    cp.C.set_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; this arm owns the host
    t = cp.C.threads()
    world = max(1, args.gpus)
    B_m = args.global_batch // world if args.global_batch else args.batch
    B_d = args.global_batch // world if args.global_batch else args.batch_detect

    def timed(fn, inputs, steps, warmup):
        ts = []
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            fn(inputs[i % len(inputs)])
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
        return float(np.sum(ts)), len(ts)

    # bounded: each step is one batch of the GPU arm's size; cap the step count so the run ends within minutes
    steps_m = max(3, min(args.steps, 40))
    steps_d = max(3, min(args.steps, 20))
    warm = max(1, min(args.warmup, 3))
    res = {}
    if args.workload in ("all", "match_encode"):
        ins = [host_inputs_match(cp.synth, 800_000 + i * B_m, B_m) for i in range(3)]
        tot, n = timed(cp.match_encode, ins, steps_m, warm)
        res["match_encode"] = (B_m * n / tot, 1e3 * tot / n, n, B_m)
    for name, kind in (("decode_nms", "normal"), ("nms_stress", "stress")):
        if args.workload in ("all", name):
            ins = [host_inputs_detect(cp.synth, 810_000 + i * B_d, B_d, kind) for i in range(2)]
            tot, n = timed(cp.detect, ins, steps_d if args.workload != "all" or name == "decode_nms" else max(3, steps_d // 4), warm)
            res[name] = (B_d * n / tot, 1e3 * tot / n, n, B_d)
    primary = args.workload if args.workload != "all" else "match_encode"
    v, ms, n, B = res[primary]
    line = {
        "impl": "reference", "metric": "images/sec (%s)" % ("match+encode" if primary == "match_encode" else "decode+NMS"),
        "value": v, "unit": "images/s", "n_gpus": args.gpus, "steps": n, "warmup": warm, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "strong" if args.global_batch else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(primary), "batch_per_gpu": B, "global_batch": B * world, "image": "512x512", "anchors": N_ANCHORS,
                   "max_gt": 100, "note": "CPU port of the reference path on the host cores; rank 0 only; the "
                                          "step count is capped (%d) so the run stays bounded" % n},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": t, "kind": "port",
                         "sample": "%d steps x %d images, C/OpenMP oracle port on %d threads" % (n, B, t)},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    for name, (v2, ms2, n2, B2) in res.items():
        if name != primary:
            line["%s_value" % name] = v2
            line["%s_e2e_value" % name] = v2
            line["%s_ms_per_step" % name] = ms2
            line["%s_batch_per_gpu" % name] = B2
    emit(line)


def workload_name(w):
    return {"match_encode": "match_encode: ARM refine_groundtruth(JACCARD_BIGGER) + ODM det_groundtruth, BASELINE configs[1]",
            "decode_nms": "decode_nms: decode + select %.2f + top-k %d + NMS %.2f keep %d, BASELINE configs[2]" % (SELECT_THR, TOP_K, NMS_THR, KEEP),
            "nms_stress": "nms_stress: decode_nms with every (anchor, class) above threshold, BASELINE configs[4]"}[w]


# ----------------------------------------------------------------------------------------- GPU arm
class Gpu:
    """Device, process group and the timing discipline shared by the workloads."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.bind_cores()
        self.dev = torch.device("cuda", self.local)
        torch.cuda.set_device(self.dev)
        if self.world > 1:
            os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's banner must not land on stdout (one JSON line)
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream(self.dev)
        self.launches = 0

    def bind_cores(self):
        """One disjoint slice of the host cores per rank, before any pinned buffer is allocated (first touch):
        eight ranks' staging copies otherwise migrate over the same cores."""
        self.cores = None
        lw = int(os.environ.get("LOCAL_WORLD_SIZE", "1"))
        if lw <= 1 or not hasattr(os, "sched_setaffinity"):
            return
        try:
            avail = sorted(os.sched_getaffinity(0))
            per = len(avail) // lw
            if per >= 1:
                mine = avail[self.local * per:(self.local + 1) * per]
                os.sched_setaffinity(0, mine)
                self.cores = [mine[0], mine[-1]]
        except OSError:
            pass

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def gather(self, x):
        """values of x on every rank."""
        if self.world == 1:
            return [float(x)]
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def capture(self, step_fn, K, n_streams=1):
        """One CUDA graph holding K consecutive steps (one host launch per K steps); with n_streams > 1 the
        steps alternate over that many forked / joined streams inside the graph."""
        torch = self.torch
        if self.args.no_graphs:
            return None
        step_fn(0)
        torch.cuda.synchronize(self.dev)
        g = torch.cuda.CUDAGraph()
        lanes = [torch.cuda.Stream(self.dev) for _ in range(n_streams)] if n_streams > 1 else None
        with torch.cuda.graph(g):
            cur = torch.cuda.current_stream(self.dev)
            if lanes:
                for s in lanes:
                    s.wait_stream(cur)
                for i in range(K):
                    with torch.cuda.stream(lanes[i % n_streams]):
                        step_fn(i)
                for s in lanes:
                    cur.wait_stream(s)
            else:
                for i in range(K):
                    step_fn(i)
        return g

    def timed_primary(self, run_loop, K, after_loop=None):
        """time_loops for the headline number, with nvidia-smi clocks / throttle reasons sampled meanwhile."""
        sampler = ClockSampler(self.local)
        if self.rank == 0:
            sampler.start()
            time.sleep(0.05)
        t = self.time_loops(run_loop, K, after_loop, min_ms=4 * MIN_TIMED_MS)
        self.clocks = sampler.stop() if self.rank == 0 else None
        return t

    def time_loops(self, run_loop, K, after_loop=None, min_ms=MIN_TIMED_MS, warm_loops=None, max_loops=400):
        """run_loop() enqueues K steps.  Warm-up loops, barrier + synchronize, then R >= 5 timed loops (each between
        its own pair of CUDA events on the launching stream) so that >= min_ms are timed, barrier + synchronize.
        Returns ms per step from the MAX over ranks of the per-rank median loop, plus the spreads."""
        torch = self.torch
        warm_loops = warm_loops if warm_loops is not None else max(1, -(-self.args.warmup // K))
        for _ in range(warm_loops):
            run_loop()
            if after_loop:
                after_loop()
        torch.cuda.synchronize(self.dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream); run_loop(); e1.record(self.stream)
        if after_loop:
            after_loop()
        torch.cuda.synchronize(self.dev)
        est = max(self.gather(e0.elapsed_time(e1)))
        R = int(min(max_loops, max(5, np.ceil(min_ms / max(est, 1e-3)))))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(R)]
        self.barrier()
        for a, b in evs:
            a.record(self.stream)
            run_loop()
            b.record(self.stream)
            if after_loop:
                after_loop()
        self.barrier()
        ts = [a.elapsed_time(b) for a, b in evs]
        med = float(np.median(ts))
        meds = self.gather(med)
        self.launches = R
        return {"ms_per_step": max(meds) / K, "loops": R, "steps_per_loop": K, "timed_ms": float(np.sum(ts)),
                "rank_spread_ms_per_step": [min(meds) / K, max(meds) / K],
                "loop_spread_ms_per_step": [float(np.min(ts)) / K, float(np.max(ts)) / K]}


def bench_match(G, table, B, extras):
    """match_encode: fused target generation (primary) and the two-call path."""
    torch = G.torch
    from rodet_b200 import _abi, config, synth
    from rodet_b200.utils import net_tools
    args, dev, N = G.args, G.dev, table.n
    JB = config.refine_method.JACCARD_BIGGER
    K = max(1, args.steps)
    n_sets = 8                    # rotating input / output sets: ~100 MB each (19 MB read + 80 MB written) >> 126 MB L2 together
    sets = []
    for s in range(n_sets):
        c, l, k, ro = host_inputs_match(synth, (G.rank * n_sets + s) * B, B)
        sets.append({"center": torch.from_numpy(c).to(dev), "labels": torch.from_numpy(l).to(dev),
                     "counts": torch.from_numpy(k).to(dev),
                     "ro": [torch.from_numpy(a).to(dev) for a in split_np(ro, SHAPES, (4,))],
                     "out": net_tools.target_buffers(table, B, dev), "mean_g": float(k.mean())})

    def fused(i):
        s = sets[i % n_sets]
        return net_tools.target_gen(table, s["center"], s["labels"], s["ro"], gt_counts=s["counts"], out=s["out"])

    def two_call(i):
        s = sets[i % n_sets]
        t = net_tools.refine_groundtruth(table, s["center"], s["labels"], JB, gt_counts=s["counts"])
        return net_tools.det_groundtruth(s["ro"], t[0], t[1], t[2], t[3], table)

    def runner(fn, n_streams=1):
        g = G.capture(fn, K, n_streams)
        if g is None:
            return lambda: [fn(i) for i in range(K)]
        return g.replay

    t1 = G.timed_primary(runner(fused), K)
    launches = t1["loops"] * K
    res = {"timing": t1, "ms_per_step": t1["ms_per_step"], "value": G.world * B / (t1["ms_per_step"] * 1e-3),
           "gpu_launches": launches, "kernels_per_step": ["target_fused_kernel"]}
    mean_g = float(np.mean([s["mean_g"] for s in sets]))
    hbm, peak_src = measured_peaks()
    alg = B * (124 * N + 20 * mean_g)             # ARM 40 N + 20 G, ODM 84 N per image
    real = B * (84 * N + 20 * mean_g)             # what the fused kernel moves: 16 N read + 68 N written
    # FP32 (no FMA) issue-rate probe: the pair loop's compute bound
    import ctypes
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    ops = ctypes.c_double(0)
    for _ in range(2):
        _abi.check(_abi.lib.rod_peak_fp32_nofma(4096, sink.data_ptr(), ctypes.addressof(ops), G.stream.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(G.stream)
    _abi.check(_abi.lib.rod_peak_fp32_nofma(4096, sink.data_ptr(), ctypes.addressof(ops), G.stream.cuda_stream))
    e1.record(G.stream)
    torch.cuda.synchronize(dev)
    fp32_tops = ops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    t_s = t1["ms_per_step"] * 1e-3
    pair_flops = 14.0 * N * mean_g * B
    frac_hbm, frac_fp32 = alg / t_s / 1e9 / hbm, pair_flops / t_s / 1e12 / fp32_tops
    res["roofline"] = {
        "bound": "hbm" if alg / (hbm * 1e9) >= pair_flops / (fp32_tops * 1e12) else "fp32_nofma",
        "kernel": "target_fused_kernel", "achieved": alg / t_s / 1e9, "peak": hbm, "unit": "GB/s", "frac": frac_hbm,
        "traffic": measured_traffic("target_fused_kernel"), "traffic_note": "ncu DRAM bytes end with the kernel: most of the 80 MB of stores are still dirty in L2",
        "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "moved_bytes_per_launch": real,
        "launch_ms": t1["ms_per_step"], "frac_fp32_nofma": frac_fp32, "fp32_nofma_tops_measured": fp32_tops,
        "pair_flops_per_launch": pair_flops,
        "note": "1 launch = 1 step; algorithmic = SURVEY 8d M (124N+20G B/img), moved = 84N+20G; FP32 pair bound = 14NG flops, no FMA"}
    if extras:
        t4 = G.time_loops(runner(fused, args.streams), K)
        t2 = G.time_loops(runner(two_call), K)
        res["value_overlapped"] = G.world * B / (t4["ms_per_step"] * 1e-3)
        res["ms_per_step_overlapped"] = t4["ms_per_step"]
        res["two_call_ms_per_step"] = t2["ms_per_step"]
        res["two_call_value"] = G.world * B / (t2["ms_per_step"] * 1e-3)
    res["e2e"] = e2e_match(G, table, B, sets[0], extras)
    res["config"] = {"workload": workload_name("match_encode"), "batch_per_gpu": B, "global_batch": B * G.world,
                     "image": "512x512", "anchors": N, "max_gt": 100, "mean_gt": mean_g,
                     "api": "net_tools.target_gen (fused ARM+ODM)", "l2": "inputs > L2: %d rotating sets of ~100 MB" % n_sets,
                     "launch": "python launches" if args.no_graphs else "CUDA graph of %d serial steps, one batch in flight" % K,
                     "parallelism": "image-sharded, no collective"}
    return res


def e2e_match(G, table, B, s0, extras):
    """Public API from pinned host buffers: ONE staging copy H2D (head output | GT | labels | counts), the fused
    kernel, ONE copy D2H of all eight outputs (68 B / anchor); two batches in flight."""
    torch = G.torch
    from rodet_b200 import _abi
    from rodet_b200.utils import net_tools
    dev, N = G.dev, table.n
    ro_host = torch.cat([t.reshape(B, -1, 4) for t in s0["ro"]], 1).cpu()
    parts = [ro_host.numpy().view(np.uint8).reshape(-1), s0["center"].cpu().numpy().view(np.uint8).reshape(-1),
             s0["labels"].cpu().numpy().view(np.uint8).reshape(-1), s0["counts"].cpu().numpy().view(np.uint8).reshape(-1)]
    sizes = [p.size for p in parts]
    h_in = torch.from_numpy(np.concatenate(parts)).pin_memory()
    gmax = s0["center"].shape[1]
    lanes = []
    for _ in range(2):
        d_in = torch.empty_like(h_in, device=dev)
        o0, o1, o2, o3 = np.cumsum([0] + sizes)[:4]
        ro = d_in[o0:o0 + sizes[0]].view(torch.float32).view(B, N, 4)
        out = net_tools.target_buffers(table, B, dev, flat=True)
        lanes.append({"d_in": d_in, "ro": _abi.LayerList(ro, table, True, False),
                      "center": d_in[o1:o1 + sizes[1]].view(torch.float32).view(B, gmax, 4),
                      "labels": d_in[o2:o2 + sizes[2]].view(torch.int64).view(B, gmax),
                      "counts": d_in[o3:o3 + sizes[3]].view(torch.int32), "out": out,
                      "h_out": torch.empty_like(out["_flat"], device="cpu").pin_memory(),
                      "h_mask": torch.empty((B, N), dtype=torch.int32).pin_memory(), "evt": None})

    def make_step(full):
        def step(i):
            L = lanes[i % 2]
            if L["evt"] is not None:
                L["evt"].synchronize()                     # the caller consumes this lane's previous result
            L["d_in"].copy_(h_in, non_blocking=True)
            net_tools.target_gen(table, L["center"], L["labels"], L["ro"], gt_counts=L["counts"], out=L["out"])
            if full:
                L["h_out"].copy_(L["out"]["_flat"], non_blocking=True)
            else:
                L["h_mask"].copy_(L["out"]["mask"], non_blocking=True)
            L["evt"] = torch.cuda.Event()
            L["evt"].record(torch.cuda.current_stream(dev))
        return step

    def run(step, K):
        fork = torch.cuda.Event()
        fork.record(G.stream)
        for s in side:
            s.wait_event(fork)
        for i in range(K):
            with torch.cuda.stream(side[i % 2]):
                step(i)
        for s in side:
            j = torch.cuda.Event()
            j.record(s)
            G.stream.wait_event(j)

    side = [torch.cuda.Stream(dev) for _ in range(2)]
    K = 8
    full = make_step(True)
    t = G.time_loops(lambda: run(full, K), K, min_ms=60.0, warm_loops=1, max_loops=20)
    h2d, d2h = int(h_in.numel()), int(lanes[0]["h_out"].numel())
    res = {"value": G.world * B / (t["ms_per_step"] * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": t["ms_per_step"],
           "h2d_GBps_per_rank": h2d / (t["ms_per_step"] * 1e-3) / 1e9, "d2h_GBps_per_rank": d2h / (t["ms_per_step"] * 1e-3) / 1e9,
           "api": "target_gen: pinned host inputs (1 H2D copy), all 8 output lists read back (1 D2H copy), 2 batches in flight"}
    if extras:
        mask = make_step(False)
        t2 = G.time_loops(lambda: run(mask, K), K, min_ms=40.0, warm_loops=1, max_loops=20)
        res["mask_only_value"] = G.world * B / (t2["ms_per_step"] * 1e-3)
        res["mask_only_d2h_bytes_per_step"] = B * N * 4
    return res


def bench_detect(G, table, B, kind, extras, primary):
    """decode_nms on `kind` scores.  Multi-GPU: every step's detection counts [11, B] land in a staging buffer
    inside the graph and ONE NCCL all-gather per loop (K batches) ships them."""
    torch = G.torch
    from rodet_b200 import synth
    from rodet_b200.dist import allgather_counts
    from rodet_b200.utils import net_tools
    args, dev, N = G.args, G.dev, table.n
    K = max(1, args.steps if primary else min(args.steps, 32))
    n_sets = max(3, args.streams) if extras else 3       # >= 3 x 123 MB of inputs > 126 MB L2; one set (and workspace) per overlapped stream
    sets = []
    for s in range(n_sets):
        p, ro, do = host_inputs_detect(synth, 500_000 + (G.rank * n_sets + s) * B, B, kind)
        sets.append({"p": [torch.from_numpy(a).to(dev) for a in split_np(p, SHAPES, (N_CLASSES,))],
                     "ro": [torch.from_numpy(a).to(dev) for a in split_np(ro, SHAPES, (4,))],
                     "do": [torch.from_numpy(a).to(dev) for a in split_np(do, SHAPES, (4,))],
                     "ws": net_tools.detect_workspace(table, B, TOP_K, dev), "host": (p, ro, do) if s == 0 else None})
    stage = torch.zeros((K, N_CLASSES, B), dtype=torch.int32, device=dev)
    kw = dict(select_threshold=SELECT_THR, nms_threshold=NMS_THR, top_k=TOP_K, keep_top_k=KEEP)

    def step(i):
        s = sets[i % n_sets]
        s["out"] = net_tools.decode_detected_bboxes(table, s["ro"], s["do"], s["p"], workspace=s["ws"],
                                                    counts_out=stage[i % K], **kw)

    pending = [None]

    def ship_counts():
        if G.world > 1:
            if pending[0] is not None:
                pending[0].result()                        # consumed one gather later
            pending[0] = allgather_counts(stage.view(K * N_CLASSES, B), B * G.world, async_op=True)

    def runner(n_streams=1):
        g = G.capture(step, K, n_streams)
        if g is None:
            return lambda: [step(i) for i in range(K)]
        return g.replay

    if G.world > 1:                                       # communicator / channels for this message size
        for _ in range(2):
            allgather_counts(stage.view(K * N_CLASSES, B), B * G.world)
    t1 = (G.timed_primary if primary else G.time_loops)(runner(), K, after_loop=ship_counts)
    if pending[0] is not None:
        pending[0].result(); pending[0] = None
    flags = float(np.mean([net_tools.detect_fallback_flags(s["ws"])[1:].float().mean().item() for s in sets]))
    hbm, peak_src = measured_peaks()
    alg = B * (76 * N + (N_CLASSES - 1) * KEEP * 20)
    t_s = t1["ms_per_step"] * 1e-3
    res = {"timing": t1, "ms_per_step": t1["ms_per_step"], "value": G.world * B / t_s, "over_rate": flags,
           "gpu_launches": 4 * t1["loops"] * K,
           "kernels_per_step": ["sample_kernel", "scan_kernel", "segment_kernel", "nms_kernel (flagged segments only: top-k + NMS)"],
           "detections_per_image": float(stage[0, 1:].sum().item()) / B,
           "roofline": {"bound": "hbm", "kernel": "whole step (sample + scan + segment kernels)", "achieved": alg / t_s / 1e9,
                        "peak": hbm, "unit": "GB/s", "frac": alg / t_s / 1e9 / hbm, "traffic": measured_traffic("decode_nms_step") if kind == "normal" else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": alg, "launch_ms": t1["ms_per_step"]},
           "config": {"workload": workload_name("nms_stress" if kind == "stress" else "decode_nms") + ("" if kind in ("normal", "stress") else " [%s-clustered scores]" % kind),
                      "batch_per_gpu": B, "global_batch": B * G.world, "image": "512x512", "anchors": N,
                      "api": "net_tools.decode_detected_bboxes", "l2": "inputs larger than L2: %d rotating sets of 123 MB" % n_sets,
                      "launch": "python launches" if args.no_graphs else "CUDA graph of %d serial steps per host launch, one batch in flight" % K,
                      "collective": ("one NCCL all_gather of the [%d x 11, B] int32 detection counts per %d batches" % (K, K)) if G.world > 1 else "none (1 GPU)"}}
    if extras:
        t4 = G.time_loops(runner(args.streams), K, after_loop=ship_counts)
        if pending[0] is not None:
            pending[0].result(); pending[0] = None
        res["value_overlapped"] = G.world * B / (t4["ms_per_step"] * 1e-3)
        res["ms_per_step_overlapped"] = t4["ms_per_step"]
    if primary or kind == "normal":
        res["e2e"] = e2e_detect(G, table, B, sets[0]["host"], kw)
    return res


def e2e_detect(G, table, B, host, kw):
    torch = G.torch
    from rodet_b200 import _abi
    from rodet_b200.utils import net_tools
    dev, N = G.dev, table.n
    p, ro, do = host
    parts = [np.ascontiguousarray(a).view(np.uint8).reshape(-1) for a in (p, ro, do)]
    sizes = [a.size for a in parts]
    h_in = torch.from_numpy(np.concatenate(parts)).pin_memory()
    lanes = []
    for _ in range(2):
        d_in = torch.empty_like(h_in, device=dev)
        v = lambda k, inner: d_in[sum(sizes[:k]):sum(sizes[:k + 1])].view(torch.float32).view(B, N, inner)
        lanes.append({"d_in": d_in, "p": table.split(v(0, N_CLASSES)), "ro": table.split(v(1, 4)), "do": table.split(v(2, 4)),
                      "ws": net_tools.detect_workspace(table, B, TOP_K, dev),
                      "s": torch.empty((N_CLASSES, B, KEEP), dtype=torch.float32).pin_memory(),
                      "b": torch.empty((N_CLASSES, B, KEEP, 4), dtype=torch.float32).pin_memory(), "evt": None})
    side = [torch.cuda.Stream(dev) for _ in range(2)]

    def step(i):
        L = lanes[i % 2]
        if L["evt"] is not None:
            L["evt"].synchronize()
        L["d_in"].copy_(h_in, non_blocking=True)
        rs, rb = net_tools.decode_detected_bboxes(table, L["ro"], L["do"], L["p"], workspace=L["ws"], **kw)
        L["s"][1:].copy_(rs[1]._base[1:], non_blocking=True)   # class-major [C,B,keep] buffers behind the dicts
        L["b"][1:].copy_(rb[1]._base[1:], non_blocking=True)
        L["evt"] = torch.cuda.Event()
        L["evt"].record(torch.cuda.current_stream(dev))

    def run(K):
        fork = torch.cuda.Event()
        fork.record(G.stream)
        for s in side:
            s.wait_event(fork)
        for i in range(K):
            with torch.cuda.stream(side[i % 2]):
                step(i)
        for s in side:
            j = torch.cuda.Event()
            j.record(s)
            G.stream.wait_event(j)

    K = 4
    t = G.time_loops(lambda: run(K), K, min_ms=60.0, warm_loops=1, max_loops=12)
    h2d = int(h_in.numel())
    d2h = int((lanes[0]["s"][1:].numel() + lanes[0]["b"][1:].numel()) * 4)
    return {"value": G.world * B / (t["ms_per_step"] * 1e-3), "unit": "images/s", "ms_per_step": t["ms_per_step"],
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "h2d_GBps_per_rank": h2d / (t["ms_per_step"] * 1e-3) / 1e9,
            "api": "net_tools.decode_detected_bboxes on pinned host inputs (1 H2D copy); scores + boxes read back; 2 batches in flight"}


def bench_next_rows(G, table, B_m, B_d):
    """SURVEY section 8 'next' rows, two scalars each (N = 1 only): f-1 decode_nms fed with class logits (softmax fused
    into the scan pass), f-3 targets + refine_loss + det_clf_loss forward as a graph replay and forward + backward eagerly."""
    torch = G.torch
    from rodet_b200 import synth
    from rodet_b200.utils import net_tools
    dev, N, out = G.dev, table.n, {}
    lay = lambda flat_, tail: [torch.from_numpy(a).to(dev) for a in split_np(flat_, SHAPES, tail)]
    # f-1
    sets = []
    for s in range(3):
        first = 700_000 + s * B_d
        z = np.stack([synth.class_logits(first + b, N_ANCHORS) for b in range(B_d)])
        _, ro, do = host_inputs_detect(synth, first, 1, "normal")
        ro = np.stack([synth.head_offsets(first + b, N_ANCHORS, 0, 0.1, 0.2) for b in range(B_d)])
        do = np.stack([synth.head_offsets(first + b, N_ANCHORS, 1, 0.1, 0.2) for b in range(B_d)])
        sets.append({"z": lay(z, (N_CLASSES,)), "ro": lay(ro, (4,)), "do": lay(do, (4,)), "ws": net_tools.detect_workspace(table, B_d, TOP_K, dev)})
    kw = dict(select_threshold=SELECT_THR, nms_threshold=NMS_THR, top_k=TOP_K, keep_top_k=KEEP, from_logits=True)

    def step_logits(i):
        s = sets[i % 3]
        s["out"] = net_tools.decode_detected_bboxes(table, s["ro"], s["do"], s["z"], workspace=s["ws"], **kw)
    K = 24
    g = G.capture(step_logits, K)
    t = G.time_loops(g.replay if g is not None else (lambda: [step_logits(i) for i in range(K)]), K)
    out["decode_nms_logits_ms_per_step"] = t["ms_per_step"]
    out["decode_nms_logits_value"] = B_d / (t["ms_per_step"] * 1e-3)
    del sets
    # f-3
    c, l, k, ro = host_inputs_match(synth, 800_000, B_m)
    cen, lab, cnt = torch.from_numpy(c).to(dev), torch.from_numpy(l).to(dev), torch.from_numpy(k).to(dev)
    ro_l = lay(ro, (4,))
    do_l = lay(np.stack([synth.head_offsets(800_000 + b, N_ANCHORS, 1) for b in range(B_m)]), (4,))
    clf_l = lay(np.stack([synth.class_logits(800_000 + b, N_ANCHORS) for b in range(B_m)]), (N_CLASSES,))
    buf = net_tools.target_buffers(table, B_m, dev)

    def forward(ro_, do_, clf_):
        arm, det = net_tools.target_gen(table, cen, lab, [t_.detach() for t_ in ro_], gt_counts=cnt, out=buf)
        rl = net_tools.refine_loss(ro_, arm[0], arm[3])
        dl, cl = net_tools.det_clf_loss(ro_, clf_, do_, det[0], det[1], det[2], det[3])
        return rl + dl + cl
    keep = {}

    def step_fwd(i):
        with torch.no_grad():
            keep["loss"] = forward(ro_l, do_l, clf_l)
    K = 8
    g = G.capture(step_fwd, K)
    t = G.time_loops(g.replay if g is not None else (lambda: [step_fwd(i) for i in range(K)]), K)
    out["targets_losses_forward_graph_ms"] = t["ms_per_step"]
    leaves = [[x.clone().requires_grad_(True) for x in ts] for ts in (ro_l, do_l, clf_l)]

    def step_bwd():
        for ts in leaves:
            for x in ts:
                x.grad = None
        forward(*leaves).backward()
    t = G.time_loops(step_bwd, 1, min_ms=30.0, warm_loops=3, max_loops=40)
    out["targets_losses_forward_backward_eager_ms"] = t["ms_per_step"]
    return out


def flat(line, prefix, r, keys=("value", "ms_per_step", "value_overlapped", "ms_per_step_overlapped", "over_rate")):
    for k in keys:
        if k in r:
            line["%s_%s" % (prefix, k)] = r[k]
    if "roofline" in r:
        line["%s_roofline_frac" % prefix] = r["roofline"]["frac"]
    if "e2e" in r:
        line["%s_e2e_value" % prefix] = r["e2e"]["value"]
    if "config" in r and prefix == "decode_nms":
        line["%s_batch_per_gpu" % prefix] = r["config"]["batch_per_gpu"]


_REAL_STDOUT = None


def claim_stdout():
    """Everything any library prints to fd 1 (NCCL's version banner, ...) goes to stderr; the one JSON line is written
    to the original stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    G = Gpu(args)
    from rodet_b200 import config
    from rodet_b200.anchor_table import AnchorTable
    from rodet_b200.utils import net_tools
    config.img_size = IMG
    anchors = net_tools.anchors_all_layer(IMG, {"layer_%d" % (i + 1): f for i, f in enumerate(FEATS)},
                                          net_tools.init_anchor(len(FEATS)))
    config.img_size = (418, 418)
    table = AnchorTable.from_anchors(anchors, G.dev)
    strong = args.global_batch > 0
    if strong and args.global_batch % G.world:
        raise SystemExit("--global-batch must be divisible by the number of ranks")
    B_m = args.global_batch // G.world if strong else args.batch
    B_d = args.global_batch // G.world if strong else args.batch_detect
    extras = not args.skip_extras
    primary = "match_encode" if args.workload == "all" else args.workload

    if primary == "match_encode":
        r = bench_match(G, table, B_m, extras)
        metric = "images/sec (match+encode)"
    else:
        r = bench_detect(G, table, B_d, "stress" if primary == "nms_stress" else "normal", extras, True)
        metric = "images/sec (decode+NMS)"
    line = {"metric": metric, "value": r["value"], "unit": "images/s", "n_gpus": G.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": r["config"], "roofline": r["roofline"], "e2e": r["e2e"], "gpu_launches": r["gpu_launches"],
            "kernels_per_step": r["kernels_per_step"], "timing": r["timing"], "clocks": getattr(G, "clocks", None)}
    for k in ("value_overlapped", "ms_per_step_overlapped", "two_call_ms_per_step", "two_call_value", "over_rate", "detections_per_image"):
        if k in r:
            line[k] = r[k]
    if G.cores:
        line["host_cores_of_rank0"] = G.cores

    if args.workload == "all" and extras:
        flat(line, "decode_nms", bench_detect(G, table, B_d, "normal", True, False))
        flat(line, "nms_stress", bench_detect(G, table, B_d, "stress", False, False))
        flat(line, "decode_nms_clustered", bench_detect(G, table, B_d, "quadrant", False, False))
        flat(line, "decode_nms_bumps", bench_detect(G, table, B_d, "bumps", False, False))
        if G.world == 1:
            line.update(bench_next_rows(G, table, B_m, B_d))

    if G.rank == 0 and not args.skip_cpu and G.world == 1:
        cb = cpu_baselines(args, B_m, B_d)
        line["cpu_baseline"] = cb["match_encode" if primary == "match_encode" else "decode_nms"]
        if primary == "match_encode":
            line["decode_nms_cpu_value"] = cb["decode_nms"]["value"]
        else:
            line["match_encode_cpu_value"] = cb["match_encode"]["value"]
        line["cpu_config0"] = cb["config0"]
    if G.rank == 0:
        emit(line)
    if G.world > 1:
        G.dist.destroy_process_group()


if __name__ == "__main__":
    main()
