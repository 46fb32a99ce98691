#!/usr/bin/env python
"""Benchmark of the descriptor-extraction hot path (BASELINE.json configs[1]) + the 1M-row similarity/top-k search.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port), rank 0 only

Headline metric: descriptor images/sec.  A "step" = one pass of the hot path over one batch: ResNet101-GeM
descriptors of 64 synthetic 1024x1024 RGB images per GPU (random-init weights of that architecture, inputs resident
in HBM).  `value` is device-timed (CUDA events, max over ranks); `e2e` is the same metric through the C-ABI host
entry point (pinned host images -> H2D -> forward -> D2H descriptors) with the copies inside the timed region.
The second half of BASELINE's metric (1M-DB queries/sec: 1000 queries x 1M x 2048, k = 100, database sharded
row-wise over the GPUs, one all-gather of the per-shard top-k) is reported under the "search" key of the same line.
Prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

ARCH = "resnet101_rmac"
BATCH, SIZE = 64, 1024
SEARCH_N, SEARCH_Q, SEARCH_D, SEARCH_K = 1_000_000, 1000, 2048, 100
CPU_THREADS = 16          # measured on the GPU host: torch CPU conv is fastest at 16 threads (128 logical cores)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--size", type=int, default=SIZE)
    ap.add_argument("--no-search", action="store_true", help="skip the 1M-row search measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--host-chunk", type=int, default=0)
    ap.add_argument("--search-n", type=int, default=SEARCH_N)
    ap.add_argument("--search-q", type=int, default=SEARCH_Q)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU while the timed region runs: NVML (a few ms per sample) when
    the bindings are importable, else nvidia-smi (tens of ms per sample)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    BITS = [0x8, 0x40, 0x20, 0x4]     # nvmlClocksEventReason{HwSlowdown, HwThermalSlowdown, SwThermalSlowdown, SwPowerCap}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.source = index, [], False, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.nvml, self.source = pynvml, "nvml"
        except Exception:
            self.nvml = None

    def _reasons_mask(self):
        for fn in ("nvmlDeviceGetCurrentClocksEventReasons", "nvmlDeviceGetCurrentClocksThrottleReasons"):
            f = getattr(self.nvml, fn, None)
            if f is not None:
                try:
                    return int(f(self.handle))
                except Exception:
                    pass
        return 0

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    mask = self._reasons_mask()
                    self.rows.append([mhz, self.max_mhz] + [bool(mask & b) for b in self.BITS])
                    time.sleep(0.004)
                    continue
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) == 6:
                    self.rows.append([float(parts[0]), float(parts[1])] + [p.lower().startswith("active") for p in parts[2:]])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = [n for j, n in enumerate(self.NAMES) if any(r[2 + j] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.rows[0][1], "reasons": reasons,
                "samples": len(self.rows), "source": self.source}


# --------------------------------------------------------------------------------------------- CPU reference arm
def cpu_extract_rate(n_img, size, repeats):
    """The reference's CPU path for extraction = oracle port of net(imgs) (torch CPU fp32)."""
    import torch
    import synthdata as synth
    from oracle import dir_oracle as O
    torch.set_num_threads(min(CPU_THREADS, os.cpu_count() or 1))
    sd = synth.make_state_dict(ARCH, seed=0)
    x = synth.make_images(n_img, size, size, seed=1234, smooth=False)
    O.extract(x[:1, :, :256, :256], sd, ARCH)                 # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    for _ in range(repeats):
        O.extract(x, sd, ARCH)
    dt = time.perf_counter() - t0
    return n_img * repeats / dt, dt / repeats, torch.get_num_threads()


def cpu_search_rate(n_db, n_q, dim, k):
    """The reference's CPU path for retrieval: np.dot scores (common.py:33) + per-query argsort (generic.py:207)."""
    import numpy as np
    r = np.random.RandomState(0)
    db = r.standard_normal((n_db, dim)).astype(np.float32)
    q = r.standard_normal((n_q, dim)).astype(np.float32)
    t0 = time.perf_counter()
    sc = np.dot(q, db.T)
    for i in range(n_q):
        np.argsort(sc[i])[::-1][:k]
    dt = time.perf_counter() - t0
    return n_q / dt, dt


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    steps = max(1, min(args.steps, 3))
    sample = 2
    rate, per_step, threads = cpu_extract_rate(sample, args.size, steps)
    qrate, qdt = cpu_search_rate(100_000, 16, SEARCH_D, SEARCH_K)
    line = {
        "impl": "reference", "metric": "descriptor images/sec", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": steps, "warmup": 1, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Resnet101-GeM descriptor extraction, %dx%d synthetic RGB (BASELINE configs[1])" % (args.size, args.size),
                   "arch": ARCH, "images_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": "%d images of %dx%d per step through oracle/dir_oracle.py (torch CPU fp32, %d threads of %d logical cores)"
                                   % (sample, args.size, args.size, threads, os.cpu_count())},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "search": {"metric": "queries/sec", "value": qrate * 100_000 / SEARCH_N, "unit": "queries/s",
                   "sample": "16 queries x 100k x 2048 np.dot + argsort (%.2f s), scaled linearly to the 1M-row database" % qdt},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import synthdata as synth
    from dirb200 import nets, ops
    from dirb200.dist import ShardedIndex, shard_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_tf_burst = peaks.get("bf16_tflops", 1590.0)
    peak_gbs = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback"

    ops.require_gpu(local)
    B, S = args.batch, args.size
    net = nets.create_model(ARCH)
    net.load_state_dict(synth.make_state_dict(ARCH, seed=0))
    net.eval()
    if args.chunk:
        net.set_backend_option("chunk", args.chunk)
    if args.host_chunk:
        net.set_backend_option("host_chunk", args.host_chunk)

    # synthetic normalised images, generated on the device (rank-dependent seed): 64 x 3 x 1024 x 1024 fp32 = 805 MB
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    u8 = torch.randint(0, 256, (B, 3, S, S), generator=g, device="cuda", dtype=torch.uint8)
    mean = torch.tensor(synth.RGB_MEANS, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(synth.RGB_STDS, device="cuda").view(1, 3, 1, 1)
    imgs = ((u8.float() / 255.0 - mean) / std).contiguous()
    del u8

    warmup = max(3, args.warmup)
    for _ in range(warmup):
        d = net.forward(imgs, want_f16=True)[0]
    torch.cuda.synchronize()
    assert bool(torch.isfinite(d).all()), "non-finite descriptors"
    launches_per_step, flops_per_step = net.last_launch_stats()

    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        net.forward(imgs, want_f16=True)
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    sampler.stop_flag = True
    sampler.join(timeout=2)
    ms_per_step = ms_total / args.steps
    value = world * B * args.steps / (ms_total * 1e-3)

    # ---- end to end through the C-ABI host entry point: pinned host images in, host descriptors out
    host = torch.empty((B, 3, S, S), dtype=torch.float32).pin_memory()
    host.copy_(imgs)
    e2e_steps = max(1, min(args.steps, 3))
    net.forward_host(host.numpy(), device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        net.forward_host(host.numpy(), device=local)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_s
    # same, from uint8 HWC host images (what a decoder yields): ToTensor + Normalize run inside the stem, 4x fewer bytes
    host8 = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8).pin_memory()
    net.forward_host_u8(host8.numpy(), device=local)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        net.forward_host_u8(host8.numpy(), device=local)
    barrier()
    e2e_u8_value = world * B * e2e_steps / max_over_ranks(time.perf_counter() - t0)
    del host8

    # ---- roofline of the dominant kernel (the persistent tcgen05 convolution), timed live with CUDA events on the
    #      launch stream during one extra, instrumented step
    net.set_backend_option("profile", 1)
    net.forward(imgs, want_f16=True)
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    net.forward(imgs, want_f16=True)
    p1.record()
    torch.cuda.synchronize()
    prof_step_ms = p0.elapsed_time(p1)
    prof = net.profile()
    net.set_backend_option("profile", 0)
    conv = prof["conv_tcgen05"]
    conv_tf = conv["flops"] / (conv["ms"] * 1e-3) / 1e12 if conv["ms"] else 0.0
    tot_ms = sum(v["ms"] for v in prof.values())
    roofline = {"bound": "tensor", "kernel": "conv_pers_kernel + conv_halo_kernel (all %d Bottleneck convolution launches of the step)" % conv["launches"],
                "achieved": conv_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": conv_tf / peak_tf, "traffic": None,
                "peak_source": peak_src + ", sustained dense 16-bit",
                "avg_launch_ms": conv["ms"] / max(1, conv["launches"]), "flops_per_launch": conv["flops"] / max(1, conv["launches"]),
                "share_of_step": conv["ms"] / tot_ms if tot_ms else None,
                "hbm_view": {"achieved_gbs": conv["bytes"] / (conv["ms"] * 1e-3) / 1e9 if conv["ms"] else 0.0, "peak_gbs": peak_gbs,
                             "note": "algorithmic activation+weight bytes of the same launches / same time"},
                "classes_ms": {k: round(v["ms"], 3) for k, v in prof.items()},
                "instrumented_step_ms": prof_step_ms}

    line = {
        "metric": "descriptor images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": "Resnet101-GeM descriptor extraction, batch %d x %dx%d synthetic RGB per GPU (BASELINE configs[1])" % (B, S, S),
                   "arch": ARCH, "global_batch": world * B, "parallelism": "image shards, dp%d, no collective" % world,
                   "l2": "inputs (%.0f MB/step) and activations exceed the 126 MB L2" % (B * 3 * S * S * 4 / 1e6),
                   "weights": "random init (synth.make_state_dict seed 0)"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S * 4,
                "d2h_bytes_per_step": B * net.descriptor_dim * 4, "steps": e2e_steps,
                "api": "dirb200_net_forward_host (pinned host buffers, H2D of chunk i+1 overlaps compute of chunk i)",
                "uint8_input": {"value": e2e_u8_value, "unit": "images/s", "h2d_bytes_per_step": B * 3 * S * S,
                                "api": "dirb200_net_forward_host_u8 (uint8 HWC pixels, normalisation fused into the stem)"}},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(),
        "roofline": roofline,
        "step_tflops": flops_per_step / (ms_per_step * 1e-3) / 1e12,
    }

    # ---- second half of the metric: queries/sec on the 1M x 2048 database, sharded row-wise over the ranks
    if not args.no_search:
        del imgs, host
        torch.cuda.empty_cache()
        N, Q, D, K = args.search_n, args.search_q, SEARCH_D, SEARCH_K
        s0, s1 = shard_rows(N, world, rank)
        gen = torch.Generator(device="cuda").manual_seed(99 + rank)
        db = torch.randn((s1 - s0, D), generator=gen, device="cuda", dtype=torch.float32)
        db, db16 = ops.l2_normalize(db, want_f16=True)
        gq = torch.Generator(device="cuda").manual_seed(7)
        q = ops.l2_normalize(torch.randn((Q, D), generator=gq, device="cuda", dtype=torch.float32))
        index = ShardedIndex(db, row_offset=s0, db16_local=db16)
        for _ in range(2):
            index.search(q, K)
        barrier()
        ssteps = max(3, args.steps)
        launches0 = index.local.stats()["launches"]
        s_e0, s_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_e0.record()
        for _ in range(ssteps):
            sc, ix = index.search(q, K)
        s_e1.record()
        barrier()
        s_ms = max_over_ranks(s_e0.elapsed_time(s_e1)) / ssteps
        qh = q.cpu().pin_memory()
        barrier()
        t0 = time.perf_counter()
        for _ in range(ssteps):
            sc, ix = index.search(qh.cuda(non_blocking=True), K)
            sc_h, ix_h = sc.cpu(), ix.cpu()
        barrier()
        s_e2e = max_over_ranks(time.perf_counter() - t0) / ssteps
        # PCA-whitening of a database block on the tensor cores (common.whiten_features, a8): 131072 x 2048 -> 2048
        wn = min(131072, s1 - s0)
        comp = torch.randn((D, D), generator=gen, device="cuda") / 45.0
        wmean = torch.zeros(D, device="cuda")
        wcs = torch.ones(D, device="cuda")
        ops.whiten(db[:wn].contiguous(), comp, wmean, wcs)
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        ops.whiten(db[:wn].contiguous(), comp, wmean, wcs)
        w1.record()
        torch.cuda.synchronize()
        w_ms = w0.elapsed_time(w1)
        st = index.local.stats()
        flops = 2.0 * Q * (N + st["dense_rows"] * world) * D
        line["search"] = {
            "metric": "1M-DB queries/sec", "value": Q / (s_ms * 1e-3), "unit": "queries/s", "ms_per_step": s_ms,
            "config": {"workload": "%d queries x %d x %d fp16 database, k=%d, exact fp64 re-scoring" % (Q, N, D, K),
                       "sharding": "rows / %d ranks, one all-gather of (score,index)[Q][k], %d B per rank" % (world, 16 * Q * K)},
            "e2e": {"value": Q / s_e2e, "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4, "d2h_bytes_per_step": Q * K * 16},
            "roofline": {"bound": "tensor", "kernel": "conv_pers_kernel<256,4,PERS_EPI_SIM_*> (seed + filter passes)",
                         "achieved": flops / world / (s_ms * 1e-3) / 1e12, "peak": peak_tf_burst, "unit": "TFLOP/s",
                         "frac": flops / world / (s_ms * 1e-3) / 1e12 / peak_tf_burst,
                         "note": "whole search step per GPU (GEMM passes + selection + re-scoring) vs burst dense 16-bit peak",
                         "hbm_gbs": (s1 - s0) * D * 2 / (s_ms * 1e-3) / 1e9},
            "stats": st, "gpu_launches": st["launches"] * ssteps,
            "whiten": {"rows": wn, "ms": w_ms, "rows_per_s": wn / (w_ms * 1e-3),
                       "tflops_algorithmic": 2.0 * wn * D * D / (w_ms * 1e-3) / 1e12,
                       "note": "x-mean -> fp16 hi/lo split, 3 tcgen05 GEMM passes, column scale, row L2 (fp32-level accuracy)"},
        }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, per_step, threads = cpu_extract_rate(2, S, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                                "sample": "2 images of %dx%d through oracle/dir_oracle.py (torch CPU fp32, %d threads of %d logical cores)"
                                          % (S, S, threads, os.cpu_count())}
        if not args.no_search:
            qrate, qdt = cpu_search_rate(100_000, 16, SEARCH_D, SEARCH_K)
            line["search"]["cpu_baseline"] = {"value": qrate * 100_000 / args.search_n, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                              "sample": "16 queries x 100k x 2048 np.dot + argsort (%.2f s), scaled to %d rows" % (qdt, args.search_n)}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
