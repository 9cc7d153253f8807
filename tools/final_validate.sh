set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_conv_tc|k_conv_simt|k_fuse|k_stn|k_iuv|k_gcn|k_maxpool|k_smpl|k_faces|k_project|k_resolve" -c 1300 --csv --log-file gpurun_out/ncu_final.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name regex:k_conv_tc -o gpurun_out/conv_tc_final python tools/tc_one.py 64,56,56,48,48,3,1 64,28,28,96,96,3,1 64,14,14,192,192,3,1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-400 gpurun_out/bench_final.json; cut -c1-300 gpurun_out/bench_ref.json
