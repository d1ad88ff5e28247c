"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum --csv` launch list (plain or .gz), the summary quoted in
profiles/README.md:   python tools/launch_summary.py profiles/r1_launches_final_eager_full_step.csv.gz"""
import collections
import csv
import gzip
import re
import sys


def main(path: str, top: int = 25) -> None:
    fh = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(fh))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    H, rows = rows[hdr], rows[hdr + 1:]
    kn, mv = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if len(r) <= mv:
            continue
        name = re.sub(r"\(.*", "", r[kn]).replace("<unnamed>::", "").replace("void ", "")
        if "base_convert" in name:
            name = re.sub(r"<\d+>", "<NS>", name)
        agg[name][0] += 1
        agg[name][1] += float(r[mv]) / 1e3
    total = sum(v[1] for v in agg.values())
    print(f"{sum(v[0] for v in agg.values())} launches, {total / 1e3:.2f} ms of kernel time (profiler-serialised)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{k:34s} n={v[0]:6d}  {v[1]:10.1f} us  {100 * v[1] / total:5.1f} %  avg {v[1] / v[0]:6.1f} us")
    ntt = sum(v[1] for k, v in agg.items() if k.startswith("ntt_"))
    print(f"NTT passes together: {100 * ntt / total:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
