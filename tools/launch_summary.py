"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and, with
--step N, the N-th head step launch by launch."""
import collections
import csv
import sys


def load(path):
    rows, hdr, out = list(csv.reader(l for l in open(path) if not l.startswith("=="))), None, []
    for r in rows:
        if len(r) < 5:
            continue
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(d["Metric Value"].replace(",", ""))
        u = d["Metric Unit"]
        v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
        out.append((d["Kernel Name"], d.get("Grid Size", ""), v))
    return out


def main():
    path = sys.argv[1]
    out = load(path)
    if "--step" in sys.argv:
        k = int(sys.argv[sys.argv.index("--step") + 1])
        idx = [i for i, o in enumerate(out) if "roi_align_fwd" in o[0]]
        start = idx[k] - 1
        end = idx[k + 1] - 1 if k + 1 < len(idx) else len(out)
        tot = 0
        for o in out[start:end]:
            print("%-58s %-16s %9.1f us" % (o[0][:58], o[1], o[2]))
            tot += o[2]
        print("sum %.1f us" % tot)
        return
    agg, tot = collections.OrderedDict(), 0.0
    for name, _, v in out:
        a = agg.setdefault(name[:70], [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-72s n=%4d %10.1f us %5.1f%%" % (k, n, t, 100 * t / tot))
    print("total %.1f us over %d launches" % (tot, len(out)))


if __name__ == "__main__":
    main()
