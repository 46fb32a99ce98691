mkdir -p gpurun_out
python tools/one_conv.py 256 1024 1 1 epi_mode=0 > gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 256 1024 1 1 epi_mode=0 res_variant=1 >> gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 256 1024 1 1 epi_mode=0 res_variant=2 >> gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 64 256 1 1 64 256 256 epi_mode=0 >> gpurun_out/one_conv.log 2>&1
python tools/one_conv.py 1024 256 1 0 >> gpurun_out/one_conv.log 2>&1
cat gpurun_out/one_conv.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_pers_kernel -s 3 -c 1 -f -o gpurun_out/r2b_c3res_1g python tools/one_conv.py 256 1024 1 1 epi_mode=0 > gpurun_out/ncu1.log 2>&1; echo rc=$?
ls -la gpurun_out/*.ncu-rep
