#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_extract.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_ext.log 2>&1
echo "== pytest rc=$?"; tail -6 gpurun_out/pytest_ext.log
for f in 1 0; do
cat > /tmp/ab.py <<PY
import sys; sys.path.insert(0, "/root/repo")
import torch, dirb200.synth as synth
from dirb200 import nets
net = nets.create_model("resnet101_rmac"); net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
net.set_backend_option("fuse_ds", $f)
x = torch.randn((64, 3, 1024, 1024), device="cuda")
for _ in range(3): net.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): net.forward(x)
e1.record(); torch.cuda.synchronize()
print("fuse_ds=$f ms/step %.2f img/s %.1f" % (e0.elapsed_time(e1) / 5, 64 * 5 / e0.elapsed_time(e1) * 1e3))
PY
timeout 300 python /tmp/ab.py 2>&1 | tail -1
done
timeout 1200 python bench.py > gpurun_out/bench_full.log 2>&1
echo "== bench rc=$?"; tail -1 gpurun_out/bench_full.log | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print("extract", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "conv TF", d["roofline"]["achieved"], d["roofline"]["classes_ms"], d["clocks"]); print("search", d["search"]["value"], d["search"]["ms_per_step"], "cpu", d.get("cpu_baseline"))'
