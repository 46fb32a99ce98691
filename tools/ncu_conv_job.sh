# ncu --set full captures of single convolution kernels (one launch each; numbers printed under ncu are never bench values)
mkdir -p gpurun_out
python tools/one_conv.py 256 1024 1 1 > gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 256 1024 1 1 epi_warps=8 >> gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 1024 256 1 0 >> gpurun_out/one_conv.log 2>&1
cat gpurun_out/one_conv.log
# layer3 1x1 expand + residual, final epilogue (16 warps)
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_pers_kernel -s 3 -c 1 -f -o gpurun_out/r2c_c3res python tools/one_conv.py 256 1024 1 1 > gpurun_out/ncu1.log 2>&1; echo rc=$?
# layer1 fused projection shortcut (K = [64 | 64] -> 256 @256x256): the first launch of that template instance in a forward
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled \
  -k 'regex:conv_pers_kernel<\(int\)256, \(int\)4, \(int\)0, \(int\)2, \(int\)16>' -s 2 -c 1 -f -o gpurun_out/r2c_fused_ds \
  python bench.py --steps 1 --warmup 3 --no-search --no-cpu-baseline --no-latency > gpurun_out/ncu2.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
