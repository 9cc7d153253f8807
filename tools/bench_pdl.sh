for v in 0 1; do
  DANET_TC_PDL=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_pdl$v.json 2> gpurun_out/bench_pdl$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pdl$v.json"))
print("PDL=$v", round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "conv_ms", round(d["roofline"]["conv_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4))
PY
done
