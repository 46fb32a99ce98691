#!/bin/bash
# compute-sanitizer passes over the operator-level GPU tests (small shapes; the sanitizer slows kernels 10-100x).
# Usage on the GPU box:  bash tools/gpu_sanitize.sh [memcheck|racecheck|synccheck|initcheck]   (default: memcheck)
# Output: gpurun_out/sanitize_<tool>.log ; exit code 0 iff the sanitizer reported no error.
# Not part of the default test run: budget ~5-10 GPU-minutes per tool.
TOOL=${1:-memcheck}
mkdir -p gpurun_out
SEL=${SANITIZE_SEL:-'test_conv_bn_act or test_stem or test_head or test_pool_scales or test_topk_small_db or test_topk_filtered or test_topk_ties or test_aqe or test_resize or test_peer_exchange or test_conv_epilogue'}
timeout 420 /usr/local/cuda/bin/compute-sanitizer --tool "$TOOL" --error-exitcode 86 --target-processes all \
  python -m pytest tests/test_gpu_ops.py tests/test_gpu_search.py -m gpu -q -x -p no:cacheprovider -k "$SEL" \
  > "gpurun_out/sanitize_${TOOL}.log" 2>&1
rc=$?
grep -E "ERROR SUMMARY|passed|failed|error" "gpurun_out/sanitize_${TOOL}.log" | tail -5
echo "== sanitize ${TOOL} rc=${rc}"
exit $rc
