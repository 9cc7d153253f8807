# Full validation bundle on the GPU box (about 8 GPU-minutes): bash tools/final_validate.sh
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
timeout 200 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
# launch list + DRAM bytes of one eager step (cold-cache, serialised: shares, not absolutes)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:"k_conv_tc|k_conv_simt|k_fuse|k_stn|k_iuv|k_gcn|k_maxpool|k_global|k_linear|k_nchw|k_smpl|k_faces|k_project|k_resolve|k_img2map" -c 1300 --csv --log-file gpurun_out/ncu_final.csv python bench.py --steps 1 --warmup 1 --no-graph --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-600 gpurun_out/bench_final.json; cut -c1-400 gpurun_out/bench_ref.json
