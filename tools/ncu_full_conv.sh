set -x
timeout 1200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_conv_tc -o gpurun_out/conv_tc_final python tools/tc_one.py 64,56,56,48,48,3,1 64,28,28,96,96,3,1 64,56,56,64,256,1,1 64,14,14,192,192,3,1 > gpurun_out/ncu_full.log 2>&1; echo rc=$?
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out/*.ncu-rep
