"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share)."""
import collections, csv, sys

def main(path, out=None):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for row in csv.DictReader(lines):
        name = row["Kernel Name"].split("(")[0]
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = row["Metric Unit"]
        v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v * 1e6 if u in ("s", "second") else v
        a = agg[name]; a[0] += 1; a[1] += v; a[2] = max(a[2], v)
    tot = sum(v[1] for v in agg.values())
    txt = ["# per-kernel device time from %s (cold-cache, serialised: compare SHARES)" % path,
           "%-64s %6s %12s %10s %7s" % ("kernel", "n", "total_us", "max_us", "share")]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        txt.append("%-64s %6d %12.1f %10.1f %6.2f%%" % (k[:64], v[0], v[1], v[2], 100 * v[1] / tot))
    txt.append("total_us %.1f" % tot)
    s = "\n".join(txt) + "\n"
    if out:
        open(out, "w").write(s)
    print(s)

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
