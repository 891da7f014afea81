#!/usr/bin/env python
"""Turn the raw ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.
  python tools/summarize_profiles.py launches gpurun_out/r01_launches_bench.csv profiles/r01_launch_shares.md
  python tools/summarize_profiles.py full     gpurun_out/stream_raw.csv          profiles/r01_ncu_stream_kernels.md
(the second input is `ncu -i X.ncu-rep --page raw --csv`)"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"b200amg::", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(src, dst):
    rows = [r for r in csv.reader(open(src, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        try:
            v = float(r[iv].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[iu] in ("ns", "nsecond") else v * 1e3 if r[iu] in ("ms", "msecond") else v * 1e6 if r[iu] in ("s", "second") else v
        k = short(r[ik])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src})\n\nper-launch times are cold-cache and serialised: compare SHARES.  total {total/1e3:.2f} ms over "
                f"{sum(a[0] for a in agg.values())} launches\n\n| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {n} | {t:.1f} | {100*t/total:.1f}% | {t/n:.1f} |\n")


def full(src, dst):
    rows = list(csv.reader(open(src, errors="replace")))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    idx = [hdr.index(w) for w in want if w in hdr]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src}); cold cache, one replayed launch each\n\n| " + " | ".join(hdr[i] + (f" [{units[i]}]" if units[i] else "") for i in idx) + " |\n")
        f.write("|" + "---|" * len(idx) + "\n")
        for r in rows[2:]:
            f.write("| " + " | ".join(short(r[i]) if hdr[i] == "Kernel Name" else r[i][:12] for i in idx) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
