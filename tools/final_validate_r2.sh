# Final validation of round 2 on the GPU box (about 7 GPU-minutes): bash tools/final_validate_r2.sh
# full GPU suite, smoke, the default bench line, and the A/B of the exact-mode MMA issue order (DANET_TC_MMAORDER).
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 150 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 400 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_final.json
DANET_TC_MMAORDER=0 timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_order0.json 2> gpurun_out/bench_order0.err; echo "bench order0 rc=$?"; cut -c1-400 gpurun_out/bench_order0.json
timeout 120 python tools/tc_layers.py order1 exact s1 > gpurun_out/layers_order1.log 2>&1; cat gpurun_out/layers_order1.log
DANET_TC_MMAORDER=0 timeout 120 python tools/tc_layers.py order0 exact s1 > gpurun_out/layers_order0.log 2>&1; cat gpurun_out/layers_order0.log
timeout 60 python tools/losses_bench.py > gpurun_out/losses_bench2.log 2>&1; tail -22 gpurun_out/losses_bench2.log
