#!/usr/bin/env python
"""Benchmark of the descriptor-extraction hot path (BASELINE.json configs[1]) + the 1M-row search.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port), rank 0 only

A "step" = one pass of the hot path over one batch: ResNet101-GeM descriptors of 64 synthetic 1024x1024 RGB
images per GPU (random-init weights of that architecture, inputs resident in HBM).  Prints ONE JSON line.
`value` is device-timed (CUDA events, max over ranks); `e2e` is the same metric through the C-ABI host entry
point (pinned host images -> H2D -> forward -> D2H descriptors) with the copies inside the timed region.
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
BATCH, HEIGHT, WIDTH = 64, 1024, 1024
FLOPS_PER_IMG = 325.99e9          # 2*MAC over conv+fc, ResNet-101 @ 1024^2
SEARCH_N, SEARCH_Q, SEARCH_D, SEARCH_K = 1_000_000, 1000, 2048, 100


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--size", type=int, default=HEIGHT)
    ap.add_argument("--no-search", action="store_true", help="skip the secondary 1M-row search measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--chunk", type=int, default=0)
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with nvidia-smi while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) == 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for j, n in enumerate(names) if any(r[2 + j].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


def cpu_reference_rate(batch, size, steps, warmup):
    """The reference's CPU path for this workload = oracle port of net(imgs) (torch CPU fp32, all host cores)."""
    import torch
    import dirb200.synth as synth
    from oracle import dir_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = synth.make_state_dict(ARCH, seed=0)
    x = synth.make_images(batch, size, size, seed=1234, smooth=False)
    for _ in range(warmup):
        O.extract(x[:1], sd, ARCH)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.extract(x, sd, ARCH)
    dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 2
    rate, per_step = cpu_reference_rate(sample, args.size, max(1, min(args.steps, 3)), 1)
    cores = os.cpu_count()
    line = {
        "impl": "reference", "metric": "descriptor images/sec", "value": rate, "unit": "images/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Resnet101-GeM descriptor extraction, %dx%d synthetic RGB (BASELINE configs[1])" % (args.size, args.size),
                   "arch": ARCH, "images_per_step": sample},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d images of %dx%d per step through oracle/dir_oracle.py (torch CPU fp32, %d threads)" % (sample, args.size, args.size, cores)},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import dirb200.synth as synth
    from dirb200 import nets, ops

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

    ops.require_gpu(local)
    B, S = args.batch, args.size
    net = nets.create_model(ARCH)
    net.load_state_dict(synth.make_state_dict(ARCH, seed=0))
    net.eval()
    if args.chunk:
        net.set_backend_option("chunk", args.chunk)

    # synthetic normalised images, generated on the device (rank-dependent seed): 64 x 3 x 1024 x 1024 fp32 = 805 MB
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    u8 = torch.randint(0, 256, (B, 3, S, S), generator=g, device="cuda", dtype=torch.uint8)
    mean = torch.tensor(synth.RGB_MEANS, device="cuda").view(1, 3, 1, 1)
    std = torch.tensor(synth.RGB_STDS, device="cuda").view(1, 3, 1, 1)
    imgs = ((u8.float() / 255.0 - mean) / std).contiguous()
    del u8

    for _ in range(max(3, args.warmup)):
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
        dh = net.forward_host(host.numpy(), device=local)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    e2e_value = world * B * e2e_steps / e2e_s
    h2d = B * 3 * S * S * 4
    d2h = B * net.descriptor_dim * 4

    line = {
        "metric": "descriptor images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16 operands, f32 accumulate", "data": "synthetic",
        "config": {"workload": "Resnet101-GeM descriptor extraction, batch %d x %dx%d synthetic RGB per GPU (BASELINE configs[1])" % (B, S, S),
                   "arch": ARCH, "global_batch": world * B, "parallelism": "image shards, dp%d, no collective" % world,
                   "l2": "inputs (805 MB/step) and activations exceed the 126 MB L2",
                   "weights": "random init (synth.make_state_dict seed 0)"},
        "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "api": "dirb200_net_forward_host (pinned host buffers)"},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": sampler.summary(),
    }
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
    achieved_tf = flops_per_step / (ms_per_step * 1e-3) / 1e12
    line["roofline"] = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                        "frac": achieved_tf / peak_tf, "traffic": None,
                        "note": "whole step (all kernels) vs %s dense 16-bit peak; algorithmic conv+fc FLOPs = %.1f GFLOP/img"
                                % ("measured sustained" if peaks else "fallback", flops_per_step / B / 1e9)}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, per_step = cpu_reference_rate(2, S, 1, 1)
        line["cpu_baseline"] = {"value": rate, "unit": "images/s", "cores": os.cpu_count(), "kind": "port",
                                "sample": "2 images of %dx%d, oracle/dir_oracle.py (torch CPU fp32)" % (S, S)}
    if rank == 0:
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
