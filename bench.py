#!/usr/bin/env python
"""bench.py — images/s of the box-level hot path on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic BDD100K-shaped input
(512x512 anchor layout, N = 36 852 anchors, 10 classes, <= 100 GT boxes per image):

  primary   match_encode  (BASELINE configs[1]): ARM matching + encoding (refine_groundtruth)
            followed by ODM target generation (det_groundtruth), batch 32 per GPU;
  secondary decode_nms    (configs[2]): decode + select + top-k 400 + NMS 0.45, batch 64 per GPU;
            nms_stress    (configs[4]): the same with every (anchor, class) above threshold.

`value` = whole-job images/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e`   = the same metric through the public Python API with HOST (pinned) buffers, the
          host->device copies of the inputs and the device->host read of the result inside
          the timed region;
`roofline` = algorithmic bytes of the dominant kernel / its measured launch time, against
          MEASURED_PEAKS.json's HBM copy bandwidth;
`cpu_baseline` = the CPU oracle (a port of the reference; the reference's own TF-1 path cannot
          run here) on a bounded sample, rank 0 only.
Multi-GPU: one process per GPU (torchrun), images sharded by rank, no data-path collective
for match_encode; decode_nms adds one NCCL all-gather of per-rank detection counts.
`--impl reference` times the CPU oracle as the reference arm.
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
SELECT_THR, NMS_THR, TOP_K, KEEP = 0.3, 0.45, 400, 200
FALLBACK_HBM_GBS = 6650.0


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--batch", type=int, default=32, help="images per GPU per step, match_encode")
    p.add_argument("--batch-detect", type=int, default=64, help="images per GPU per step, decode_nms")
    p.add_argument("--no-graphs", action="store_true", help="launch through Python every step instead of CUDA graphs")
    p.add_argument("--streams", type=int, default=4,
                   help="CUDA streams that consecutive (independent) batches alternate on; 1 = strictly serial steps")
    p.add_argument("--skip-secondary", action="store_true")
    p.add_argument("--skip-cpu", action="store_true")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    return p.parse_args()


# ----------------------------------------------------------------------------------------- helpers
def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as f:
                d = json.load(f)
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


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
        time.sleep(0.15)
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_anchors():
    from rodet_b200 import config
    from rodet_b200.utils import net_tools
    config.img_size = IMG
    sizes = net_tools.init_anchor(len(FEATS))
    feats = {"layer_%d" % (i + 1): f for i, f in enumerate(FEATS)}
    anchors = net_tools.anchors_all_layer(IMG, feats, sizes)
    config.img_size = (418, 418)
    return anchors


def split_np(flat, shapes, tail):
    out, off = [], 0
    for fh, fw, a in shapes:
        n = fh * fw * a
        out.append(np.ascontiguousarray(flat[:, off:off + n]).reshape((flat.shape[0], fh, fw, a) + tail))
        off += n
    return out


SHAPES = [(fh, fw, 6 if i == 0 else 9) for i, (fh, fw) in enumerate(FEATS)]
N_ANCHORS = sum(a * b * c for a, b, c in SHAPES)


def host_inputs_match(first_image, B):
    """Host (NumPy) inputs of one match_encode step: GT in centre form + ARM head output."""
    from rodet_b200 import synth
    from oracle_free_math import corner_to_center_np
    corner, labels, counts = synth.gt_batch(first_image, B)
    center = corner_to_center_np(corner)
    for b in range(B):
        center[b, counts[b]:] = 0
    ro = np.stack([synth.head_offsets(first_image + b, N_ANCHORS) for b in range(B)])
    return center.astype(np.float32), labels, counts, ro


def host_inputs_detect(first_image, B, stress):
    from rodet_b200 import synth
    mk = synth.stress_probs if stress else synth.class_probs
    s = (0.05, 0.05) if stress else (0.1, 0.2)
    probs = np.stack([mk(first_image + b, N_ANCHORS) for b in range(B)])
    ro = np.stack([synth.head_offsets(first_image + b, N_ANCHORS, 0, *s) for b in range(B)])
    do = np.stack([synth.head_offsets(first_image + b, N_ANCHORS, 1, *s) for b in range(B)])
    return probs, ro, do


# ----------------------------------------------------------------------------------------- reference arm / CPU baseline
# The reference's own TF-1 CPU path cannot run (TensorFlow is not installable here), so the CPU side
# is the oracle port: oracle/c (plain C, OpenMP over all host threads), pinned bit-for-bit to
# oracle/restated.py and through it to the fixtures produced by the unmodified reference.
def cpu_match_encode(table, center, labels, counts, ro):
    from oracle import c_port as C
    from oracle import restated as R
    g = C.arm_match_encode(table, center, labels, counts, R.REFINE_POS_JAC)
    return C.odm_target(table, ro, g[0], g[1], g[2], g[3], R.DET_POS_JAC)


def cpu_detect(table, probs, ro, do):
    from oracle import c_port as C
    return C.detected_bboxes(probs, C.decode_corner(table, ro, do), SELECT_THR, NMS_THR, TOP_K, KEEP)


def cpu_baseline(workload, seconds):
    """Times the CPU oracle port on a bounded sample of the same workload (rank 0, N=1)."""
    from oracle import c_port as C
    from oracle import restated as R
    table = R.AnchorTable(make_anchors())
    per = 8
    n_img, t_total = 0, 0.0
    while t_total < seconds and n_img < 512:
        if workload == "match_encode":
            c, l, k, ro = host_inputs_match(900_000 + n_img, per)
            t0 = time.perf_counter(); cpu_match_encode(table, c, l, k, ro); t_total += time.perf_counter() - t0
        else:
            p, ro, do = host_inputs_detect(900_000 + n_img, per, workload == "nms_stress")
            t0 = time.perf_counter(); cpu_detect(table, p, ro, do); t_total += time.perf_counter() - t0
        n_img += per
    return {"value": n_img / t_total, "unit": "images/s", "cores": C.threads(), "kind": "port",
            "sample": "%d images of the %s workload in batches of %d, C/OpenMP oracle port (oracle/c) on %d threads; "
                      "the reference's TF-1 CPU path cannot run here (no TensorFlow)" % (n_img, workload, per, C.threads())}


def run_reference(args):
    """Reference arm: the CPU oracle port with all host threads, rank 0 only, 8 images per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import c_port as C
    from oracle import restated as R
    C.set_threads(os.cpu_count() or 1)          # torchrun exports OMP_NUM_THREADS=1; this arm owns the host
    table = R.AnchorTable(make_anchors())
    per_step = 8
    inputs = [host_inputs_match(800_000 + i * per_step, per_step) for i in range(4)]
    times = []
    for i in range(args.warmup + args.steps):
        c, l, k, ro = inputs[i % len(inputs)]
        t0 = time.perf_counter()
        cpu_match_encode(table, c, l, k, ro)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
    total = float(np.sum(times))
    v = per_step * len(times) / total
    line = {
        "impl": "reference", "metric": "images/sec (match+encode)", "value": v, "unit": "images/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "match_encode: ARM refine_groundtruth(JACCARD_BIGGER) + ODM det_groundtruth, "
                               "BASELINE configs[1]", "image": "512x512", "anchors": N_ANCHORS, "max_gt": 100,
                   "images_per_step": per_step,
                   "note": "CPU port of the reference path (the TF-1 reference cannot run: no TensorFlow); "
                           "%d images per step so the run stays bounded" % per_step},
        "cpu_baseline": {"value": v, "unit": "images/s", "cores": C.threads(), "kind": "port",
                         "sample": "%d steps x %d images, C/OpenMP oracle port on %d threads" % (len(times), per_step, C.threads())},
        "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    from rodet_b200 import _abi, config
    from rodet_b200.anchor_table import AnchorTable
    from rodet_b200.utils import net_tools

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version banner must not land on stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    JB = config.refine_method.JACCARD_BIGGER
    anchors = make_anchors()
    table = AnchorTable.from_anchors(anchors, dev)
    hbm_gbs, peak_src = measured_peaks()
    N = table.n
    stream = torch.cuda.current_stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    side_streams = [torch.cuda.Stream(dev) for _ in range(max(1, args.streams))]

    def time_loop(step_fn, steps, warmup, n_streams=1):
        """K steps between two CUDA events on the launching stream.  With n_streams > 1 consecutive
        steps (independent batches, disjoint buffers) alternate over that many streams, all forked
        from / joined to the timing stream, so a batch's memory-bound kernels overlap the next
        batch's issue-bound ones."""
        lanes = side_streams[:n_streams] if n_streams > 1 else [stream]

        def run(first, count):
            if n_streams > 1:
                fork = torch.cuda.Event()
                fork.record(stream)
                for s in lanes:
                    s.wait_event(fork)
            for i in range(count):
                with torch.cuda.stream(lanes[i % len(lanes)]):
                    step_fn(first + i)
            if n_streams > 1:
                for s in lanes:
                    j = torch.cuda.Event()
                    j.record(s)
                    stream.wait_event(j)
        run(0, warmup)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run(warmup + (warmup % 2), steps)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    def to_dev_list(flat, tail):
        return [torch.from_numpy(a).to(dev) for a in split_np(flat, SHAPES, tail)]

    # =================================================================== match_encode (primary)
    B = args.batch
    # several independent input/output sets, rotated every step, so the working set of
    # consecutive steps (~150 MB each: 66 MB read + 80 MB written) exceeds the 126 MB L2
    n_sets = max(4, 2 * args.streams)
    sets = []
    for s in range(n_sets):
        c, l, k, ro = host_inputs_match((rank * n_sets + s) * B, B)
        sets.append({"center": torch.from_numpy(c).to(dev), "labels": torch.from_numpy(l).to(dev),
                     "counts": torch.from_numpy(k).to(dev), "ro": to_dev_list(ro, (4,)), "host": (c, l, k, ro)})

    def arm(s):
        return net_tools.refine_groundtruth(table, s["center"], s["labels"], JB, gt_counts=s["counts"])

    def odm(s, t):
        return net_tools.det_groundtruth(s["ro"], t[0], t[1], t[2], t[3], table)

    for s in sets:                       # persistent outputs per set (also the ODM inputs)
        s["arm_out"] = arm(s)
        s["odm_out"] = odm(s, s["arm_out"])
    torch.cuda.synchronize(dev)

    use_graphs = not args.no_graphs
    launches_per_step = 2
    if use_graphs:
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                t = arm(s)
                s["graph_out"] = odm(s, t)
            s["graph"] = g
            ga, go = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(ga):
                s["ga_out"] = arm(s)
            with torch.cuda.graph(go):
                s["go_out"] = odm(s, s["arm_out"])
            s["g_arm"], s["g_odm"] = ga, go
        step_m = lambda i: sets[i % n_sets]["graph"].replay()
        step_arm = lambda i: sets[i % n_sets]["g_arm"].replay()
        step_odm = lambda i: sets[i % n_sets]["g_odm"].replay()
    else:
        step_m = lambda i: odm(sets[i % n_sets], arm(sets[i % n_sets]))
        step_arm = lambda i: arm(sets[i % n_sets])
        step_odm = lambda i: odm(sets[i % n_sets], sets[i % n_sets]["arm_out"])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = time_loop(step_m, args.steps, args.warmup, args.streams if use_graphs else 1)
    ms_step = ms_total / args.steps
    ms_step_serial = time_loop(step_m, args.steps, args.warmup, 1) / args.steps if args.streams > 1 else ms_step
    value = world * B * args.steps / (ms_total * 1e-3)

    # per-kernel launch times for the roofline (same stream, CUDA events, rotating sets)
    ms_arm = time_loop(step_arm, args.steps, max(3, args.warmup // 4)) / args.steps
    ms_odm = time_loop(step_odm, args.steps, max(3, args.warmup // 4)) / args.steps
    clocks = sampler.stop() if rank == 0 else None      # sampled across the timed region and the per-kernel loops
    mean_g = float(np.mean([s["host"][2].mean() for s in sets]))
    bytes_arm = B * (40 * N + 20 * mean_g)                     # write 40 N, read 20 G
    bytes_odm = B * 84 * N                                     # read 56 N + write 28 N
    dom = "odm_target_kernel" if ms_odm >= ms_arm else "arm_jaccard_bigger_kernel"
    dom_bytes, dom_ms = (bytes_odm, ms_odm) if ms_odm >= ms_arm else (bytes_arm, ms_arm)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            traffic = json.load(f).get(dom)
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": hbm_gbs, "unit": "GB/s",
                "frac": achieved / hbm_gbs, "traffic": traffic,
                "traffic_source": "profiles/r01_traffic.json (ncu dram bytes per launch)", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": dom_ms,
                "kernels": {"arm_jaccard_bigger_kernel": {"ms": ms_arm, "GBps": bytes_arm / (ms_arm * 1e-3) / 1e9,
                                                          "frac": bytes_arm / (ms_arm * 1e-3) / 1e9 / hbm_gbs},
                            "odm_target_kernel": {"ms": ms_odm, "GBps": bytes_odm / (ms_odm * 1e-3) / 1e9,
                                                  "frac": bytes_odm / (ms_odm * 1e-3) / 1e9 / hbm_gbs}},
                "step_frac_of_hbm_roofline": (bytes_arm + bytes_odm) / (ms_step * 1e-3) / 1e9 / hbm_gbs}

    # FP32 (no-FMA) issue-rate probe: the ARM loop's compute roofline denominator
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    import ctypes
    ops = ctypes.c_double(0)
    for _ in range(2):
        _abi.check(_abi.lib.rod_peak_fp32_nofma(4096, sink.data_ptr(), ctypes.addressof(ops), stream.cuda_stream))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record(stream)
    _abi.check(_abi.lib.rod_peak_fp32_nofma(4096, sink.data_ptr(), ctypes.addressof(ops), stream.cuda_stream))
    e1.record(stream)
    torch.cuda.synchronize(dev)
    fp32_nofma_tops = ops.value / (e0.elapsed_time(e1) * 1e-3) / 1e12
    roofline["note"] = ("ARM is bound by FP32 non-FMA issue, not HBM: evaluating all 14*N*G pair flops at the measured "
                        "FADD/FMUL rate would take %.1f us per launch; the step as a whole is graded against HBM "
                        "(step_frac_of_hbm_roofline)" % (14.0 * N * mean_g * B / (fp32_nofma_tops * 1e12) * 1e6))
    roofline["fp32_nofma_tops_measured"] = fp32_nofma_tops
    roofline["arm_pair_flops_frac"] = (14.0 * N * mean_g * B / (ms_arm * 1e-3)) / (fp32_nofma_tops * 1e12)

    # ---- e2e: host (pinned) buffers, H2D of the step's inputs + D2H of the result inside the timed region
    hs = sets[0]["host"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_center, h_labels, h_counts = pin(hs[0]), pin(hs[1]), pin(hs[2])
    h_ro = [pin(a) for a in split_np(hs[3], SHAPES, (4,))]
    E2E_LANES = 2        # double buffering: a batch's copies overlap the other lane's kernels / copies
    lanes_m = []
    for _ in range(E2E_LANES):
        lanes_m.append({"center": torch.empty_like(h_center, device=dev), "labels": torch.empty_like(h_labels, device=dev),
                        "counts": torch.empty_like(h_counts, device=dev),
                        "ro": [torch.empty_like(x, device=dev) for x in h_ro],
                        "mask": torch.empty((B, N), dtype=torch.int32).pin_memory(), "evt": None})
    h2d = sum(x.numel() * x.element_size() for x in [h_center, h_labels, h_counts] + h_ro)
    d2h = B * N * 4

    def step_e2e(i):
        L = lanes_m[i % E2E_LANES]
        if L["evt"] is not None:
            L["evt"].synchronize()                         # the caller consumes this lane's previous result
        L["center"].copy_(h_center, non_blocking=True)
        L["labels"].copy_(h_labels, non_blocking=True)
        L["counts"].copy_(h_counts, non_blocking=True)
        for d, h in zip(L["ro"], h_ro):
            d.copy_(h, non_blocking=True)
        t = net_tools.refine_groundtruth(table, L["center"], L["labels"], JB, gt_counts=L["counts"])
        o = net_tools.det_groundtruth(L["ro"], t[0], t[1], t[2], t[3], table)
        L["mask"].copy_(o[1].flat, non_blocking=True)      # ODM positive mask, flat [B,N]
        L["evt"] = torch.cuda.Event()
        L["evt"].record(torch.cuda.current_stream(dev))

    e2e_steps = max(6, min(args.steps, 50))
    ms_e2e = float(np.median([time_loop(step_e2e, e2e_steps, 4, E2E_LANES) for _ in range(3)]))   # host / PCIe variance
    e2e = {"value": world * B * e2e_steps / (ms_e2e * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / e2e_steps,
           "api": "net_tools.refine_groundtruth + net_tools.det_groundtruth on pinned host inputs; "
                  "result read back = ODM positive mask [B,N] int32; two batches in flight (double-buffered "
                  "device inputs on 2 streams), every batch's H2D + kernels + D2H inside the timed region"}

    line = {
        "metric": "images/sec (match+encode)", "value": value, "unit": "images/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "match_encode: ARM refine_groundtruth(JACCARD_BIGGER) + ODM det_groundtruth, "
                               "BASELINE configs[1]", "batch_per_gpu": B, "global_batch": B * world,
                   "image": "512x512", "anchors": N, "max_gt": 100, "mean_gt": mean_g,
                   "l2": "inputs larger than L2: %d rotating input/output sets (~150 MB each)" % n_sets,
                   "launch": ("CUDA graph replay; consecutive batches alternate over %d streams" % args.streams)
                   if use_graphs else "python launches", "ms_per_step_1stream": ms_step_serial,
                   "parallelism": "image-sharded, no collective"},
        "roofline": roofline, "e2e": e2e, "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
    }

    # =================================================================== decode_nms (secondary)
    if not args.skip_secondary:
        for name, stress in (("decode_nms", False), ("nms_stress", True)):
            line[name] = bench_detect(args, name, stress, dev, rank, world, table, to_dev_list, time_loop, hbm_gbs, N)

        if world == 1:
            line["decode_nms_from_logits"] = bench_detect_logits(args, dev, table, to_dev_list, time_loop, hbm_gbs, N)
            line["targets_and_losses"] = bench_losses(args, dev, table, sets, arm, odm, to_dev_list, time_loop, hbm_gbs, N)

    if rank == 0 and not args.skip_cpu and world == 1:
        line["cpu_baseline"] = cpu_baseline("match_encode", args.cpu_seconds)
        if not args.skip_secondary:
            line["decode_nms"]["cpu_baseline"] = cpu_baseline("decode_nms", args.cpu_seconds)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_detect(args, name, stress, dev, rank, world, table, to_dev_list, time_loop, hbm_gbs, N):
    import torch
    import torch.distributed as dist
    from rodet_b200.dist import allgather_counts
    from rodet_b200.utils import net_tools
    B = args.batch_detect
    n_sets = max(2, args.streams)                   # >= 2 x 180 MB of inputs > 126 MB L2; one set per stream
    sets = []
    for s in range(n_sets):
        p, ro, do = host_inputs_detect(500_000 + (rank * n_sets + s) * B, B, stress)
        sets.append({"probs": to_dev_list(p, (N_CLASSES,)), "ro": to_dev_list(ro, (4,)), "do": to_dev_list(do, (4,)),
                     "host": (p, ro, do) if s == 0 else None})

    def run(s):
        rs, rb, cnt = net_tools.decode_detected_bboxes(table, s["ro"], s["do"], s["probs"], select_threshold=SELECT_THR,
                                                       nms_threshold=NMS_THR, top_k=TOP_K, keep_top_k=KEEP,
                                                       return_counts=True)
        return rs, rb, cnt

    for s in sets:
        s["out"] = run(s)
    torch.cuda.synchronize(dev)
    use_graphs = not args.no_graphs
    if use_graphs:
        for s in sets:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                s["gout"] = run(s)
            s["graph"] = g

    pending = [None]
    ROUNDS = 4                                            # rounds (of n_sets batches) per count all-gather
    stage = torch.zeros((ROUNDS, N_CLASSES, n_sets * B), dtype=torch.int32, device=dev) if world > 1 else None

    def step(i):
        s = sets[i % n_sets]
        if use_graphs:
            s["graph"].replay()
            s["cnt"] = s["gout"][2]
        else:
            s["cnt"] = run(s)[2]
        if world > 1:
            # One NCCL all-gather of detection counts per ROUNDS * n_sets batches (the counts
            # only size the evaluation arrays, so the collective is amortised and its result is consumed one
            # gather later).  Every round's counts are packed into a staging buffer on the device.
            cur = torch.cuda.current_stream(dev)
            s["evt"] = torch.cuda.Event()
            s["evt"].record(cur)
            if (i + 1) % n_sets == 0:
                for o in sets:
                    if o.get("evt") is not None:
                        cur.wait_event(o["evt"])
                r = ((i + 1) // n_sets - 1) % ROUNDS
                torch.cat([o["cnt"] for o in sets], dim=1, out=stage[r])      # [C, n_sets * B]
                if r == ROUNDS - 1:
                    if pending[0] is not None:
                        pending[0].result()
                    pending[0] = allgather_counts(stage.view(ROUNDS * N_CLASSES, n_sets * B), n_sets * B * world, async_op=True)

    steps = max(10, args.steps // 4)
    ns = args.streams if use_graphs else 1
    if world > 1:      # establish the NCCL communicator / channels for this message size outside the timed region
        for _ in range(2):
            allgather_counts(stage.view(ROUNDS * N_CLASSES, n_sets * B), n_sets * B * world)
        torch.cuda.synchronize(dev)
    wu = max(2 * n_sets, args.warmup // 4)
    if world > 1:
        wu = max(wu, 2 * ROUNDS * n_sets)               # two full gather periods before the clock starts
    ms = time_loop(step, steps, wu, ns)
    if pending[0] is not None:
        pending[0].result()
        pending[0] = None
    ms_serial = time_loop(step, steps, wu, 1) if ns > 1 else ms
    if pending[0] is not None:
        pending[0].result()
        pending[0] = None
    value = world * B * steps / (ms * 1e-3)
    alg_bytes = B * (76 * N + (N_CLASSES - 1) * KEEP * 20)
    res = {"metric": "images/sec (decode+NMS)", "value": value, "unit": "images/s", "ms_per_step": ms / steps,
           "ms_per_step_1stream": ms_serial / steps, "streams": ns,
           "steps": steps, "batch_per_gpu": B,
           "config": {"workload": "%s: decode + select %.2f + top-k %d + NMS %.2f keep %d, BASELINE configs[%d]" % (
               name, SELECT_THR, TOP_K, NMS_THR, KEEP, 4 if stress else 2),
               "collective": ("one NCCL all_gather of [%d*11, %d*B] int32 detection counts per %d batches, consumed one gather later" % (ROUNDS, n_sets, ROUNDS * n_sets)) if world > 1 else "none (1 GPU)"},
           "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms / steps * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                        "frac": alg_bytes / (ms / steps * 1e-3) / 1e9 / hbm_gbs, "algorithmic_bytes_per_step": alg_bytes,
                        "scope": "whole step (scan + segment kernels)"},
           "gpu_launches": 5 * steps, "kernels_per_step": ["sample_kernel", "scan_kernel", "segment_kernel", "topk_segment_kernel (flagged segments only)", "nms_kernel (flagged segments only)"], "detections_per_image": float(sets[0]["out"][2].sum().item()) / B}

    # e2e through the public API with pinned host inputs and results read back
    p, ro, do = sets[0]["host"]
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_p = [pin(a) for a in split_np(p, SHAPES, (N_CLASSES,))]
    h_ro = [pin(a) for a in split_np(ro, SHAPES, (4,))]
    h_do = [pin(a) for a in split_np(do, SHAPES, (4,))]
    lanes_d = []
    for _ in range(2):
        lanes_d.append({"p": [torch.empty_like(x, device=dev) for x in h_p], "ro": [torch.empty_like(x, device=dev) for x in h_ro],
                        "do": [torch.empty_like(x, device=dev) for x in h_do],
                        "s": torch.empty((N_CLASSES, B, KEEP), dtype=torch.float32).pin_memory(),
                        "b": torch.empty((N_CLASSES, B, KEEP, 4), dtype=torch.float32).pin_memory(), "evt": None})
    h_s, h_b = lanes_d[0]["s"], lanes_d[0]["b"]

    def step_e2e(i):
        L = lanes_d[i % 2]
        if L["evt"] is not None:
            L["evt"].synchronize()                         # the caller consumes this lane's previous result
        for dl, hl in ((L["p"], h_p), (L["ro"], h_ro), (L["do"], h_do)):
            for d, h in zip(dl, hl):
                d.copy_(h, non_blocking=True)
        rs, rb, cnt = net_tools.decode_detected_bboxes(table, L["ro"], L["do"], L["p"], select_threshold=SELECT_THR,
                                                       nms_threshold=NMS_THR, top_k=TOP_K, keep_top_k=KEEP,
                                                       return_counts=True)
        L["s"][1:].copy_(rs[1]._base[1:], non_blocking=True)   # class-major [C,B,keep] buffers behind the dicts
        L["b"][1:].copy_(rb[1]._base[1:], non_blocking=True)
        L["evt"] = torch.cuda.Event()
        L["evt"].record(torch.cuda.current_stream(dev))

    e2e_steps = max(4, min(steps, 10))
    ms_e = float(np.median([time_loop(step_e2e, e2e_steps, 2, 2) for _ in range(3)]))
    res["e2e"] = {"value": world * B * e2e_steps / (ms_e * 1e-3), "unit": "images/s", "ms_per_step": ms_e / e2e_steps,
                  "h2d_bytes_per_step": sum(x.numel() * 4 for x in h_p + h_ro + h_do),
                  "d2h_bytes_per_step": (h_s[1:].numel() + h_b[1:].numel()) * 4,
                  "api": "net_tools.decode_detected_bboxes on pinned host inputs; scores+boxes read back; two batches in flight"}
    return res


def bench_losses(args, dev, table, sets, arm, odm, to_dev_list, time_loop, hbm_gbs, N):
    """the training epilogue ARM + ODM targets -> refine_loss + det_clf_loss (forward), and the
    same with backward() to the head outputs, on the match_encode batch (B = 32)."""
    import torch
    from rodet_b200 import synth
    from rodet_b200.utils import net_tools
    B = args.batch
    s = sets[0]
    first = 800_000
    s["do"] = to_dev_list(np.stack([synth.head_offsets(first + b, N_ANCHORS, 1) for b in range(B)]), (4,))
    s["clf"] = to_dev_list(np.stack([synth.class_logits(first + b, N_ANCHORS) for b in range(B)]), (N_CLASSES,))

    def forward(ro, do, clf):
        t = arm(s)
        d = net_tools.det_groundtruth(ro, t[0], t[1], t[2], t[3], table)
        rl = net_tools.refine_loss(ro, t[0], t[3])
        dl, cl = net_tools.det_clf_loss(ro, clf, do, d[0], d[1], d[2], d[3])
        return rl + dl + cl

    def step_fwd(i):
        with torch.no_grad():
            s["loss"] = forward(s["ro"], s["do"], s["clf"])

    leaves = [[t.clone().requires_grad_(True) for t in s[k]] for k in ("ro", "do", "clf")]

    def step_bwd(i):
        for ts in leaves:
            for t in ts:
                t.grad = None
        forward(*leaves).backward()

    steps = max(10, args.steps // 8)
    res = {"config": {"workload": "ARM + ODM targets + refine_loss + det_clf_loss (hard-negative mining), B = %d; "
                                  "forward / forward_backward are eager launches through autograd, forward_graph is a CUDA-graph replay" % B}}
    variants = [("forward", step_fwd), ("forward_backward", step_bwd)]
    if not args.no_graphs:
        step_fwd(0)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g), torch.no_grad():
            s["gloss"] = forward(s["ro"], s["do"], s["clf"])
        variants.append(("forward_graph", lambda i: g.replay()))
    for name, fn in variants:
        ms = time_loop(fn, steps, 3, 1)
        res["ms_per_step_" + name] = ms / steps
    res.update({"metric": "images/sec (targets + losses, forward)", "value": B / (res["ms_per_step_forward"] * 1e-3), "unit": "images/s",
                "loss": float(s["loss"]), "batch_per_gpu": B})
    return res


def bench_detect_logits(args, dev, table, to_dev_list, time_loop, hbm_gbs, N):
    """the decode_nms workload fed with class LOGITS.  Fused = softmax inside the select
    pass (no probability tensor in HBM); unfused = net_tools.softmax (writes [B,N,11]) + the same detect."""
    import torch
    from rodet_b200 import synth
    from rodet_b200.utils import net_tools
    B = args.batch_detect
    n_sets = max(2, args.streams)
    sets = []
    for s in range(n_sets):
        first = 700_000 + s * B
        z = np.stack([synth.class_logits(first + b, N_ANCHORS) for b in range(B)])
        ro = np.stack([synth.head_offsets(first + b, N_ANCHORS, 0, 0.1, 0.2) for b in range(B)])
        do = np.stack([synth.head_offsets(first + b, N_ANCHORS, 1, 0.1, 0.2) for b in range(B)])
        d = {"z": to_dev_list(z, (N_CLASSES,)), "ro": to_dev_list(ro, (4,)), "do": to_dev_list(do, (4,))}
        d["p"] = [torch.empty_like(t) for t in d["z"]]
        sets.append(d)
    kw = dict(select_threshold=SELECT_THR, nms_threshold=NMS_THR, top_k=TOP_K, keep_top_k=KEEP, return_counts=True)

    def fused(s):
        return net_tools.decode_detected_bboxes(table, s["ro"], s["do"], s["z"], from_logits=True, **kw)

    def unfused(s):
        net_tools.softmax(s["z"], out=s["p"])
        return net_tools.decode_detected_bboxes(table, s["ro"], s["do"], s["p"], **kw)

    res = {}
    steps = max(10, args.steps // 4)
    ns = 1 if args.no_graphs else args.streams
    for name, fn in (("fused", fused), ("softmax_then_detect", unfused)):
        for s in sets:
            s["out_" + name] = fn(s)
        torch.cuda.synchronize(dev)
        if not args.no_graphs:
            for s in sets:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    s["gout_" + name] = fn(s)
                s["graph_" + name] = g
        step = (lambda i, name=name: sets[i % n_sets]["graph_" + name].replay()) if not args.no_graphs else \
               (lambda i, fn=fn: fn(sets[i % n_sets]))
        ms = time_loop(step, steps, 2 * n_sets, ns)
        res["ms_per_step_" + name] = ms / steps
    same = all(torch.equal(s["out_fused"][0][c], s["out_softmax_then_detect"][0][c]) for s in sets for c in range(1, N_CLASSES))
    alg_bytes = B * (76 * N + (N_CLASSES - 1) * KEEP * 20)
    ms = res["ms_per_step_fused"]
    res.update({"metric": "images/sec (softmax+decode+NMS)", "value": B / (ms * 1e-3), "unit": "images/s", "streams": ns,
                "batch_per_gpu": B, "fused_equals_unfused": bool(same),
                "config": {"workload": "decode_nms fed with class logits (SURVEY.md section 8 f-1): softmax fused into the select pass"},
                "roofline": {"bound": "hbm", "achieved": alg_bytes / (ms * 1e-3) / 1e9, "peak": hbm_gbs, "unit": "GB/s",
                             "frac": alg_bytes / (ms * 1e-3) / 1e9 / hbm_gbs, "algorithmic_bytes_per_step": alg_bytes},
                "gpu_launches": 5 * steps,
                "detections_per_image": float(sets[0]["out_fused"][2].sum().item()) / B})
    return res


# tiny NumPy helper kept outside oracle/ so the GPU arm never imports the oracle
class _M:
    @staticmethod
    def corner_to_center_np(cr):
        cr = np.asarray(cr, dtype=np.float32)
        return np.stack([(cr[..., 0] + cr[..., 2]) / np.float32(2), (cr[..., 1] + cr[..., 3]) / np.float32(2),
                         cr[..., 2] - cr[..., 0], cr[..., 3] - cr[..., 1]], -1)


sys.modules["oracle_free_math"] = _M


if __name__ == "__main__":
    main()
