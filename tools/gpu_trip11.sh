#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gpu_diag.py > gpurun_out/diag.log 2>&1; grep -E "tap_3x3|random_3x3|multiimg" gpurun_out/diag_conv.txt | cut -c1-330
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_extract.py -m gpu -q -p no:cacheprovider > gpurun_out/pytest_conv.log 2>&1
echo "== pytest rc=$?"; tail -6 gpurun_out/pytest_conv.log
for h in 1 0; do
cat > /tmp/halo_ab.py <<PY
import sys; sys.path.insert(0, "/root/repo")
import torch, dirb200.synth as synth
from dirb200 import nets
net = nets.create_model("resnet101_rmac"); net.load_state_dict(synth.make_state_dict("resnet101_rmac", seed=0))
net.set_backend_option("halo", $h)
x = torch.randn((64, 3, 1024, 1024), device="cuda")
for _ in range(3): net.forward(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): net.forward(x)
e1.record(); torch.cuda.synchronize()
print("halo=$h ms/step %.2f img/s %.1f" % (e0.elapsed_time(e1) / 5, 64 * 5 / e0.elapsed_time(e1) * 1e3))
PY
timeout 300 python /tmp/halo_ab.py 2>&1 | tail -1
done
