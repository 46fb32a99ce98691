"""In-situ A/B of the convolution tuning knobs on ResNet-101 64 x 1024^2: for every option set, a sustained region of
forwards timed with CUDA events (img/s) and the per launch-type table of the same region (profile=2), so a change shows up
both in the step time and in the launch type it targets.  Usage: python tools/conv_sweep.py "name:k=v,k=v" ...
(no arguments = the default list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import synthdata as synth
from dirb200 import nets

B = int(os.environ.get("SWEEP_B", "64"))
STEPS = int(os.environ.get("SWEEP_STEPS", "6"))
net = nets.create_model("resnet101_rmac")
net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
x = torch.randn((B, 3, 1024, 1024), device="cuda")
net.forward(x)

sets = sys.argv[1:] or [
    "1g-late:epi_mode=0", "2g-late:epi_mode=1", "1g-early:epi_mode=2", "2g-early:epi_mode=3",
    "1g-late+l2pf:epi_mode=0,l2_prefetch=1", "2g-late+l2pf:epi_mode=1,l2_prefetch=1",
    "1g-early+l2pf:epi_mode=2,l2_prefetch=1", "2g-early+l2pf:epi_mode=3,l2_prefetch=1",
    "2g-early rv1<256,2,6>:epi_mode=3,res_variant=1", "2g-early rv1+l2pf:epi_mode=3,res_variant=1,l2_prefetch=1",
    "2g-early rv2<128,4,6>:epi_mode=3,res_variant=2", "2g-early rv2+l2pf:epi_mode=3,res_variant=2,l2_prefetch=1",
    "2g-early rv3<256,2,8>:epi_mode=3,res_variant=3", "2g-early rv3+l2pf:epi_mode=3,res_variant=3,l2_prefetch=1",
    "1g-early rv3+l2pf:epi_mode=2,res_variant=3,l2_prefetch=1", "1g-late again:epi_mode=0"]
defaults = {}
ref = None
for spec in sets:
    name, _, kv = spec.partition(":")
    opts = dict((k, float(v)) for k, v in (p.split("=") for p in kv.split(",") if p))
    for k in defaults:
        net.set_backend_option_live(k, defaults[k])
    for k, v in opts.items():
        defaults.setdefault(k, 0.0)
        net.set_backend_option_live(k, v)
    for _ in range(3):
        d = net.forward(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        d = net.forward(x)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    if ref is None:
        ref = d.clone()
    net.set_backend_option_live("profile", 2)
    for _ in range(3):
        net.forward(x)
    torch.cuda.synchronize()
    table = net.profile_table()
    net.set_backend_option_live("profile", 0)
    conv = [r for r in table if r["cls"] == 0]
    cms = sum(r["ms"] for r in conv) / 3
    cfl = sum(r["flops"] for r in conv) / 3
    print("== %-22s %.2f ms/step %.1f img/s | conv %.2f ms %.0f TFLOP/s | identical=%s maxdiff=%.2e" % (
        name, ms, B / ms * 1e3, cms, cfl / cms / 1e9, bool(torch.equal(ref, d)), float((ref - d).abs().max())), flush=True)
    top = sorted(table, key=lambda r: -r["ms"])[:int(os.environ.get("SWEEP_TOP", "8"))]
    for r in top:
        n = r["launches"]
        print("     %-52s x%-3d %7.1f us/launch  %6.0f TFLOP/s %6.0f GB/s" % (
            r["tag"][:52], n // 3, r["ms"] / n * 1e3, r["flops"] / (r["ms"] * 1e-3) / 1e12 if r["ms"] else 0,
            r["bytes"] / (r["ms"] * 1e-3) / 1e9 if r["ms"] else 0), flush=True)
