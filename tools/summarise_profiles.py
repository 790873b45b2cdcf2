"""Turn the raw captures in gpurun_out/ into the tracked summaries under profiles/ (per round tag)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def launch_list(tag, out):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)
        a = agg.setdefault(row["Kernel Name"][:90], [0, 0.0])
        a[0] += 1
        a[1] += ms
        total += ms
    with open(out, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none, command: python bench.py --steps 2 "
                f"--warmup 3 --no-cpu-baseline --no-clocks  (tag {tag}; per-launch times are cold-cache and serialised)\n")
        f.write(f"# total kernel time {total:.2f} ms over {sum(a[0] for a in agg.values())} launches\n")
        f.write("ms_total,launches,ms_per_launch,share,kernel\n")
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{t:.3f},{c},{t / c:.3f},{t / total:.4f},{k}\n")


def main(tag):
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launch_list(tag, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    rep = os.path.join(ROOT, "gpurun_out", f"prof_tc_{tag}.ncu-rep")
    with open(os.path.join(ROOT, "profiles", f"{tag}_tc_kernels_ncu.txt"), "w") as f:
        sys.stdout = f
        print(f"# ncu --set full --clock-control none -k regex:k_tc_(bitmap|merge|gallop) -s 3 -c 3, scale 24, tag {tag}")
        ncu_summary.main(rep)
        sys.stdout = sys.__stdout__
    for name in (f"bench_{tag}.json", f"sweep_variants_{tag}.jsonl"):
        src = os.path.join(ROOT, "gpurun_out", name)
        if os.path.exists(src):
            with open(src) as fi, open(os.path.join(ROOT, "profiles", f"{tag}_{name.replace('_' + tag, '')}"), "w") as fo:
                fo.write(fi.read())
    print("profiles written for", tag)


if __name__ == "__main__":
    main(sys.argv[1])
