import json, sys
d = json.load(open("gpurun_out/tc_debug.json"))
bad = [k for k, v in d.items() if not (isinstance(v, dict) and v.get("max_err") is not None and v["max_err"] < 1e-2 and v["nan_frac"] == 0)]
bad = [k for k in bad if d[k] != "unsupported"]
print("bad cases:", bad)
sys.exit(1 if bad or len(d) < 20 else 0)
