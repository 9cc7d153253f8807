"""Sums dram bytes / duration of the conv launches of one bench step from an ncu CSV
(metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum) -> profiles/conv_traffic.json"""
import collections, csv, json, sys
path, width, batch, convs_per_step = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])   # conv LAUNCHES per step (bench line: config.conv_path)
lines = [l for l in open(path) if not l.startswith("==")]
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
for row in csv.DictReader(lines):
    name = row["Kernel Name"].split("(")[0]
    try:
        v = float(row["Metric Value"].replace(",", ""))
    except ValueError:
        continue
    u = row["Metric Unit"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1)
    per[name][row["Metric Name"]] += v * scale
    if row["Metric Name"] == "gpu__time_duration.sum":
        cnt[name] += 1
conv = [k for k in per if "k_conv" in k]
nsteps = sum(cnt[k] for k in conv) / float(convs_per_step)      # the capture spans warm-up + timed + e2e + profile steps
rd = sum(per[k]["dram__bytes_read.sum"] for k in conv) / nsteps
wr = sum(per[k]["dram__bytes_write.sum"] for k in conv) / nsteps
us = sum(per[k]["gpu__time_duration.sum"] for k in conv) / nsteps
out = {"width": width, "batch": batch, "dram_bytes_per_step": rd + wr, "dram_read_bytes_per_step": rd,
       "dram_write_bytes_per_step": wr, "conv_us_per_step_under_ncu": us, "conv_launches_per_step": sum(cnt[k] for k in conv) / nsteps,
       "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over the conv launches of %.2f bench steps, averaged per step (%s)" % (nsteps, path)}
json.dump(out, open("profiles/conv_traffic.json", "w"), indent=1)
print(out)
for k in sorted(per, key=lambda k: -per[k]["gpu__time_duration.sum"])[:12]:
    print("%-50s n=%4d %10.1f us  rd %8.1f MB wr %8.1f MB" % (k[:50], cnt[k], per[k]["gpu__time_duration.sum"], per[k]["dram__bytes_read.sum"] / 1e6, per[k]["dram__bytes_write.sum"] / 1e6))
