timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu ${BENCH_EXTRA} > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_quick.json"))
r=d["roofline"]
print(round(d["value"],1), "img/s", round(d["ms_per_step"],3), "ms  e2e", round(d["e2e"]["value"],1), "conv_ms", round(r["conv_ms_per_step"],3), "frac", round(r["frac"],4))
print({k: round(v,3) for k,v in r["other_ms"].items()})
PY
