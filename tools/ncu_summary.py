"""Summarise an `ncu --csv --log-file` launch list (profiles/*.csv) per kernel: launches, total / average duration and
share, and -- when captured -- DRAM and L2 bytes per launch.  python tools/ncu_summary.py file.csv [file2.csv ...]"""
import collections
import csv
import sys

TIME = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def summarise(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    if not rows:
        return []
    ci = {k: i for i, k in enumerate(rows[0])}
    agg = collections.defaultdict(lambda: collections.defaultdict(float))
    cnt = collections.Counter()
    for r in rows[1:]:
        name = r[ci["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        m, unit = r[ci["Metric Name"]], r[ci["Metric Unit"]]
        try:
            v = float(r[ci["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        if unit in TIME:
            v *= TIME[unit]
        elif unit in BYTES:
            v *= BYTES[unit]
        agg[name][m] += v
        if m == "gpu__time_duration.sum":
            cnt[name] += 1
    total = sum(a["gpu__time_duration.sum"] for a in agg.values())
    out = []
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"]):
        n = max(cnt[name], 1)
        dram = (a.get("dram__bytes_read.sum", 0.0) + a.get("dram__bytes_write.sum", 0.0)) / n
        out.append((name, n, a["gpu__time_duration.sum"], a["gpu__time_duration.sum"] / n, 100 * a["gpu__time_duration.sum"] / total,
                    dram / 1e6 if dram else None, a.get("lts__t_bytes.sum", 0.0) / n / 1e6 or None))
    return out, total


def main():
    for path in sys.argv[1:]:
        res, total = summarise(path)
        print("### %s  (total %.1f us)\n" % (path, total))
        print("| kernel | launches | total us | avg us | share | DRAM MB / launch | L2 MB / launch |")
        print("|---|---|---|---|---|---|---|")
        for name, n, t, avg, share, dram, l2 in res:
            print("| `%s` | %d | %.1f | %.1f | %.1f %% | %s | %s |" % (name[:60], n, t, avg, share,
                  "%.1f" % dram if dram is not None else "-", "%.1f" % l2 if l2 is not None else "-"))
        print()


if __name__ == "__main__":
    main()
