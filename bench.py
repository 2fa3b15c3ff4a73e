#!/usr/bin/env python
"""bench.py — the retrieval hot path (calc_map_k: pack -> hist -> scan -> rank -> mAP) on N GPUs of one node.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C3|C4-64|...]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
         bench.py --gpus N --steps K --warmup W

One JSON line on stdout (rank 0).  A "step" is one full evaluation of the workload: the reference-format
inputs (+-1 fp32 codes, int64 multi-hot labels) are bit-packed, every query is ranked against the whole
gallery and the mAP scalar is produced.  metric = query x gallery pairs per second.

  value   inputs resident in HBM when the clock starts (CUDA events, max over ranks, L2 flushed between steps)
  e2e     the same through the reference-facing call calc_utils.calc_map_k(host tensors): pinned host buffers,
          H2D copies and the D2H of the result inside the timed region
  N > 1   weak scaling: every rank holds one gallery shard of the workload's size (total gallery = N x shard),
          queries replicated; exchange (NCCL) = all-gather of per-shard bucket totals + AP partials (mAP) or bucket totals +
          one all-reduce(MAX) of the [Q, k] key buffer (top-k; --topk-exchange allgather_merge = BASELINE's literal all-gather).

--impl reference times the reference's own CPU evaluator (oracle/calc_utils_port.py: the same ATen CPU ops as
common/calc_utils.py, the Python reference itself cannot travel to the GPU box) on a bounded query sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from clip_based_cross_modal_hash_b200 import synth  # noqa: E402

METRIC = "hamming_retrieval_query_x_gallery_pairs_per_sec"
UNIT = "pairs/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback (MEASURED_PEAKS.json absent)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            with open(path) as f:
                d = json.load(f)
            if "hbm_gbs" in d:
                return float(d["hbm_gbs"]), "measured"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def make_inputs(cfg, seed, n_items=None):
    Q, N, K, C = cfg["Q"], n_items or cfg["N"], cfg["K"], cfg["C"]
    return (synth.random_codes(Q, K, seed), synth.random_codes(N, K, seed + 1),
            synth.random_labels(Q, C, seed + 2), synth.random_labels(N, C, seed + 3))


# ---------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU evaluator on a bounded query sample
# ---------------------------------------------------------------------------------------------------------
def cpu_reference_pairs_per_sec(cfg, sample_q, steps, warmup):
    from oracle import calc_utils_port as port

    qB, rB, qL, rL = make_inputs(cfg, 1234)
    qB, qL = qB[:sample_q], qL[:sample_q]
    k = cfg["k"]
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        port.calc_map_k(qB, rB, qL, rL, k, stable=False, query_chunk=100)  # as shipped: unstable sort
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    per = sum(times) / len(times)
    return sample_q * cfg["N"] / per, per


def run_reference(args, cfg, name):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    sample_q = max(50, min(cfg["Q"], int(5e7 // cfg["N"])))   # ~1.5 s of CPU work per step on 16 threads
    v, per = cpu_reference_pairs_per_sec(cfg, sample_q, args.steps, args.warmup)
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": name, "Q": cfg["Q"], "N": cfg["N"], "bits": cfg["K"], "classes": cfg["C"], "k": cfg["k"],
                   "note": "reference CPU evaluator (torch CPU ops of common/calc_utils.py:58-92, restated in "
                           "oracle/calc_utils_port.py), one step = %d queries x full gallery" % sample_q},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d of %d queries x %d gallery items per step" % (sample_q, cfg["Q"], cfg["N"])},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if not args.no_encode:
        iv, idt = cpu_encode_images_per_sec()
        line["encode"] = {"metric": "clip_encode_images_per_sec", "value": iv, "unit": "img/s", "impl": "reference",
                          "cpu_baseline": {"value": iv, "unit": "img/s", "cores": cores, "kind": "port",
                                           "sample": "512 images (batches of 32) through the fp32 CPU restatement of encode_image, %.1f s" % idt}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
# CLIP ViT-B/32 encode (second half of BASELINE.json's metric: imgs/sec), batch 256 per GPU
# ---------------------------------------------------------------------------------------------------------
ENCODE_BATCH = 256


def tensor_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["bf16_tflops_sustained"]), float(d["bf16_tflops"]), "measured"
    except Exception:
        return 1400.0, 1590.0, "fallback"


def cpu_encode_images_per_sec(n_images=512, chunk=32):
    """The reference's fp32 CPU path (oracle/clip_port.py restates models/CLIP/model.py:232-268) on a bounded sample:
    `n_images` images in batches of `chunk` (a few seconds on 16 host threads)."""
    from oracle import clip_port as port

    sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
    image = synth.random_images(chunk, seed=1)
    with torch.no_grad():
        port.encode_image(sd, image[:4])
        t0 = time.perf_counter()
        for _ in range(n_images // chunk):
            port.encode_image(sd, image)
        dt = time.perf_counter() - t0
    return n_images / dt, dt


def bench_encode(args, dev, world, rank, barrier, all_max):
    """get_code step of BASELINE's C2 method (DCMHT, 64 bit) on random-init ViT-B/32: images -> CLIP tower -> hash head ->
    packed 64-bit codes."""
    from clip_based_cross_modal_hash_b200 import models
    from oracle import clip_port as port

    B = ENCODE_BATCH
    sd = synth.clip_state_dict(synth.VIT_B32, seed=0)
    model = models.DCMHT(sd, synth.dcmht_head_state_dict(512, 64, seed=1), device=dev)
    nbuf = 3  # 3 x 154 MB of images > 126 MB L2: every step reads its batch from HBM
    host_img = [synth.random_images(B, seed=10 + i + 100 * rank).pin_memory() for i in range(nbuf)]
    text, _ = synth.random_captions(B, seed=20 + rank)
    host_txt = text.pin_memory()
    d_img = [t.to(dev) for t in host_img]
    d_txt = host_txt.to(dev)
    steps, warm = args.steps, max(args.warmup, 3)

    def timed(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    for i in range(warm):
        model.encode_image_packed(d_img[i % nbuf])
        model.encode_text_packed(d_txt)
    img_ms = all_max(timed(lambda i: model.encode_image_packed(d_img[i % nbuf]), steps)) / steps
    txt_ms = all_max(timed(lambda i: model.encode_text_packed(d_txt), steps)) / steps
    # the image tower alone (without head/pack) for the tensor-pipe roofline
    tower_ms = all_max(timed(lambda i: model.backbone.encode_image(d_img[i % nbuf]), steps)) / steps
    # end to end: models.get_code over host (pinned) batches, H2D inside the timed region, packed codes read back
    loader = [(host_img[i % nbuf], host_txt, None, None, torch.arange(B) + B * i) for i in range(steps)]
    models.get_code(model, loader[:2], 2 * B, dev)
    e2e_runs = []
    for _ in range(2):   # two passes over the same `steps` batches, the faster one is reported (host-side jitter on shared boxes)
        barrier()
        t0 = time.perf_counter()
        ci, ct = models.get_code(model, loader, steps * B, dev)
        codes_host = ci.cpu()
        e2e_runs.append(all_max((time.perf_counter() - t0) * 1e3) / steps)
    e2e_ms = min(e2e_runs)
    sustained, burst, kind = tensor_peak()
    fl = port.flops_image()
    ach = fl * B / (tower_ms * 1e-3) / 1e12
    out = {
        "metric": "clip_encode_images_per_sec", "value": B * world / (img_ms * 1e-3), "unit": "img/s", "batch_per_gpu": B,
        "ms_per_batch": img_ms, "what": "DCMHT get_code step: fp32 NCHW images (resident in HBM) -> ViT-B/32 tower (bf16 tcgen05 GEMMs, "
        "fp32 residual stream) -> DCMHT head (fp32) -> pair argmax -> 64-bit packed codes; random-init weights, synthetic images",
        "text": {"value": B * world / (txt_ms * 1e-3), "unit": "captions/s", "ms_per_batch": txt_ms, "tokens": 32},
        "e2e": {"value": B * world / (e2e_ms * 1e-3), "unit": "image+caption pairs/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": B * 3 * 224 * 224 * 4 + B * 32 * 8 + B * 8, "d2h_bytes_per_step": int(codes_host.numel() * 4 // steps),
                "what": "models.get_code over pinned host batches (image + caption), copies overlapped on a side stream; best of 2 passes",
                "ms_per_step_runs": e2e_runs},
        "roofline": {"bound": "tensor", "kernel": "image tower (12 blocks, 50 tokens)", "achieved": ach, "peak": sustained,
                     "unit": "TFLOP/s", "frac": ach / sustained, "frac_of_burst_peak": ach / burst, "peak_kind": kind + " cuBLAS bf16, sustained",
                     "traffic": None, "algorithmic_flops_per_image": fl, "tower_ms": tower_ms},
        "gpu_launches_per_step": 12 * 7 + 10,
        "l2": "3 image batches of 154 MB are cycled (462 MB > 126 MB L2): every step reads its images from HBM",
        "gpu_launches": (12 * 7 + 10) * steps * 2 + (12 * 7 + 10) * steps,
    }
    return out


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def run_ours(args, cfg, name):
    import torch.distributed as dist

    from clip_based_cross_modal_hash_b200 import calc_utils, retrieval as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Q, N, K, C, k = cfg["Q"], cfg["N"], cfg["K"], cfg["C"], cfg["k"]

    # reference-format inputs: this rank's gallery shard (weak scaling: one workload-sized shard per rank)
    qB, rB, qL, rL = make_inputs(cfg, 1234 + 7 * rank)
    if world > 1:
        qB, _, qL, _ = make_inputs(cfg, 1234)  # queries replicated
    host = [t.pin_memory() for t in (qB, rB, qL, rL)]
    d_qB, d_rB, d_qL, d_rL = (t.to(dev) for t in host)
    if k is None and args.op == "topk":
        raise SystemExit("top-k needs a workload with k")
    if args.op == "topk":
        host = host[:2]
    h2d = sum(t.numel() * t.element_size() for t in host)
    pinned_keys = torch.empty((Q, k), dtype=torch.int64).pin_memory() if args.op == "topk" else None

    st = R.CudaStages()
    ev = R.ShardedEvaluator(stages=st) if world > 1 else None
    plan = st.make_plan(Q, N, K, C)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    keys_buf = torch.empty((Q, k), dtype=torch.int64, device=dev) if args.op == "topk" else None

    def step(events=None):
        """pack + evaluate; returns the fp64 mAP (device).  events: optional list collecting stage boundaries."""
        def mark():
            if events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                events.append(e)
        mark()
        bad = R.new_bad_counter(dev)
        qp, gp = R.pack_codes(d_qB, bad), R.pack_codes(d_rB, bad)
        if args.op == "topk":
            mark()
            if world > 1:
                keys = ev.topk(qp, gp, K, k, rank * N, n_geom=N, method=args.topk_exchange)
                mark()
                return keys
            pl = st.make_plan(Q, N, K, 0)
            hist = st.hist(pl, qp, None, gp, None)
            mark()
            sc = st.scan(pl, hist, 1, 0, k, with_rel=False)
            mark()
            keys = st.rank_topk(pl, qp, gp, sc, k, 0, keys=keys_buf)
            mark()
            mark()
            return keys
        qlp, glp = R.pack_labels(d_qL, bad), R.pack_labels(d_rL, bad)
        mark()
        if world > 1:
            res = ev.map_k(qp, qlp, gp, glp, K, C, k, n_geom=N)
            mark()
            return res.map
        hist = st.hist(plan, qp, qlp, gp, glp)
        mark()
        sc = st.scan(plan, hist, 1, 0, k)
        mark()
        app = st.rank_map(plan, qp, qlp, gp, glp, sc)
        mark()
        _, m = st.map_finish(plan, app, sc["total"])
        mark()
        return m

    def e2e_step():
        if args.op == "topk":
            qp = R.pack_codes(host[0].to(dev, non_blocking=True))
            gp = R.pack_codes(host[1].to(dev, non_blocking=True))
            keys = ev.topk(qp, gp, K, k, rank * N, n_geom=N, method=args.topk_exchange) if world > 1 else R.topk(qp, gp, K, k)
            return pinned_keys.copy_(keys)
        if world == 1:
            return calc_utils.calc_map_k(host[0], host[1], host[2], host[3], k)
        qp = R.pack_codes(host[0].to(dev, non_blocking=True))
        gp = R.pack_codes(host[1].to(dev, non_blocking=True))
        qlp = R.pack_labels(host[2].to(dev, non_blocking=True))
        glp = R.pack_labels(host[3].to(dev, non_blocking=True))
        return ev.map_k(qp, qlp, gp, glp, K, C, k, n_geom=N).map.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
        flush.zero_()
    barrier()

    stage_ms = None
    per_step = []
    with ClockSampler(local_rank) as clocks:
        barrier()
        for _ in range(args.steps):
            flush.zero_()
            evs = []
            step(evs)
            torch.cuda.synchronize()
            per_step.append(evs[0].elapsed_time(evs[-1]))
            if world == 1:
                d = [evs[i].elapsed_time(evs[i + 1]) for i in range(len(evs) - 1)]
                stage_ms = d if stage_ms is None else [a + b for a, b in zip(stage_ms, d)]
        barrier()
        total_ms = sum(per_step)
        # end-to-end through the public call, host buffers, copies inside the timed region
        e2e_ms = []
        for it in range(2 + args.steps):
            flush.zero_()
            barrier()
            t0 = time.perf_counter()
            out = e2e_step()
            torch.cuda.synchronize()
            if it >= 2:
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
        barrier()
    t = torch.tensor([total_ms, sum(e2e_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_total_ms = t.tolist()
    pairs_per_step = Q * N * world
    value = pairs_per_step * args.steps / (total_ms * 1e-3)
    e2e_value = pairs_per_step * len(e2e_ms) / (e2e_total_ms * 1e-3)

    def all_max(v):
        tt = torch.tensor([v], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    encode = None
    if not args.no_encode:
        del d_qB, d_rB, d_qL, d_rL, flush
        torch.cuda.empty_cache()
        with ClockSampler(local_rank) as enc_clocks:
            encode = bench_encode(args, dev, world, rank, barrier, all_max)
        encode["clocks"] = enc_clocks.summary()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_kind = peaks()
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": name, "Q": Q, "N_per_gpu": N, "N_total": N * world, "bits": K, "classes": C, "k": k,
                   "step": "pack(+-1 fp32 codes, int64 labels) -> hist -> scan -> rank/AP -> mAP",
                   "l2": "256 MiB flush write between timed steps", "op": args.op,
                   "topk_exchange": args.topk_exchange if (args.op == "topk" and world > 1) else None, "map": float(out.item()) if (out is not None and args.op == "map") else None},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16 if args.op == "map" else Q * k * 8,
                "ms_per_step": e2e_total_ms / len(e2e_ms)},
        "gpu_launches": args.steps * 10,
        "clocks": clocks.summary(),
    }
    if encode is not None:
        line["encode"] = encode
        line["gpu_launches"] += encode["gpu_launches"]
    if world == 1:
        names = (["pack", "hist_kernel", "scan", "rank_map_kernel", "map_finish"] if args.op == "map"
                 else ["pack", "hist_kernel", "scan", "rank_topk_kernel", "none"])
        stage = {n: v / args.steps for n, v in zip(names, stage_ms)}
        W, LW = plan.W, plan.LW
        # algorithmic bytes of the dominant kernel (rank_map): gallery codes+labels once, query codes+labels,
        # rank bases in (within + below, all + rel), AP partials out   
        alg = (N * (W + LW) * 4 + Q * (W + LW) * 4 + 2 * plan.within_elems * 4 + 2 * plan.below_elems * 4
               + Q * 4 + plan.ap_elems * 8)
        dom = "rank_map_kernel"
        if args.op == "topk":  # dominant = pass 1; compulsory bytes: gallery + query codes in, histograms out
            dom = "hist_kernel"
            alg = N * W * 4 + Q * W * 4 + plan.Qpad * (K + 1) * 4 * st.make_plan(Q, N, K, 0).nchunks
        achieved = alg / (stage[dom] * 1e-3) / 1e9
        line["roofline"] = {
            "bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": None, "peak_kind": peak_kind, "algorithmic_bytes_per_launch": alg,
            "kernel_ms": stage[dom], "share_of_step": stage[dom] / (total_ms / args.steps),
            "note": "the ranking kernels keep the gallery in L2/shared memory and never materialise Q x N; they are "
                    "bound by the integer/LSU issue rate (XOR+POPC+2 shared-memory counter updates per pair), not by HBM",
            "pairs_per_sec_kernel": Q * N / (stage[dom] * 1e-3),
        }
        # the bound SURVEY §8(d) names for this path: the POPC pipe (16 lanes/clk/SM), one popc.b32 per code word per pair and
        # per ranking pass (hist + rank = 2 passes)
        sm_clock = (line["clocks"].get("sm_mhz") or 1965.0) * 1e6
        popc_peak = 16.0 * 148 * sm_clock   # measured 15.8 popc/clk/SM on this pool (scripts/micro/popc_peak.cu, profiles/README.md)
        passes_ms = stage["hist_kernel"] + stage[dom] if dom != "hist_kernel" else stage["hist_kernel"] + stage["rank_topk_kernel"]
        line["roofline"]["popc_bound"] = {"popc_per_step": 2 * Q * N * W, "peak_popc_per_s": popc_peak,
                                          "frac": (2.0 * Q * N * W / popc_peak) / (passes_ms * 1e-3),
                                          "note": "fraction of the POPC-pipe bound reached by the two ranking passes together"}
        line["stage_ms"] = stage
        torch.set_num_threads(os.cpu_count() or 1)
        sample_q = max(50, min(Q, int(2e8 // N)))   # ~6 s of CPU work per pass (one warm-up pass, one timed)
        v, per = cpu_reference_pairs_per_sec(cfg, sample_q, 1, 1)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": "%d of %d queries x %d gallery items, %.1f s" % (sample_q, Q, N, per)}
        if encode is not None:
            iv, idt = cpu_encode_images_per_sec()
            encode["cpu_baseline"] = {"value": iv, "unit": "img/s", "cores": torch.get_num_threads(), "kind": "port",
                                      "sample": "512 images (batches of 32) through the fp32 CPU restatement of encode_image, %.1f s" % idt}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(synth.CONFIGS))
    ap.add_argument("--topk-exchange", default="rank_scatter", choices=["rank_scatter", "allgather_merge"])
    ap.add_argument("--no-encode", action="store_true", help="skip the CLIP encode section of the line")
    ap.add_argument("--op", default=None, choices=["map", "topk"],
                    help="map = calc_map_k (default for C1-C3); topk = Hamming + per-query top-k (default for C4-*)")
    args = ap.parse_args()
    if args.op is None:
        args.op = "topk" if args.workload.startswith("C4") else "map"
    cfg = synth.CONFIGS[args.workload]
    if args.impl == "reference":
        run_reference(args, cfg, args.workload)
    else:
        run_ours(args, cfg, args.workload)


if __name__ == "__main__":
    main()
