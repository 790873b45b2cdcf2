"""Turn the raw captures in gpurun_out/ into the tracked summaries under profiles/ (per round tag).

    python tools/summarise_profiles.py <tag>

Inputs (written by tools/gpu_profile.sh on the GPU box): launches_<tag>.csv, prof_tc_<tag>.ncu-rep,
bench_<tag>.json, sweep_variants_<tag>.jsonl.  Outputs: profiles/<tag>_*.{csv,txt,json,jsonl} and
profiles/traffic.json (DRAM bytes per launch of each triangle kernel; bench.py's roofline.traffic reads it).
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402


def to_ms(value, unit):
    v = float(value.replace(",", ""))
    return v / 1e6 if unit.startswith("n") else (v / 1e3 if unit.startswith("u") else v)


def launch_list(tag, out):
    path = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    agg, total = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        ms = to_ms(row["Metric Value"], row["Metric Unit"])
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


def traffic(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    kn, rd, wr, du = (hdr.index(x) for x in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum",
                                             "gpu__time_duration.sum"))
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    out = {}
    for r in rows[2:]:
        for name in ("k_tc_bitmap", "k_tc_merge", "k_tc_gallop"):
            if name in r[kn]:
                try:
                    b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
                except (ValueError, KeyError):
                    continue
                # the capture holds ONE counting step; the bitmap kernel runs as two launches (small / wide windows)
                e = out.setdefault(name, {"dram_bytes_per_launch": 0.0, "ncu_duration_ms": 0.0, "launches_summed": 0,
                                          "capture": f"profiles/{tag}_tc_kernels_ncu.txt",
                                          "workload": "kronecker-24, one GPU, one counting step"})
                e["dram_bytes_per_launch"] += b
                e["ncu_duration_ms"] += to_ms(r[du], units[du])
                e["launches_summed"] += 1
    with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
        json.dump(out, f, indent=1)


def main(tag):
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    launch_list(tag, os.path.join(ROOT, "profiles", f"{tag}_launches.csv"))
    rep = os.path.join(ROOT, "gpurun_out", f"prof_tc_{tag}.ncu-rep")
    with open(os.path.join(ROOT, "profiles", f"{tag}_tc_kernels_ncu.txt"), "w") as f:
        sys.stdout = f
        print(f"# ncu --set full --clock-control none -k regex:k_tc_(bitmap|merge|gallop) -s 3 -c 3, scale 24, tag {tag}")
        ncu_summary.main(rep)
        sys.stdout = sys.__stdout__
    traffic(rep, tag)
    forced = os.path.join(ROOT, "gpurun_out", f"prof_forced_{tag}.ncu-rep")
    if os.path.exists(forced):
        with open(os.path.join(ROOT, "profiles", f"{tag}_forced_variants_ncu.txt"), "w") as f:
            sys.stdout = f
            print(f"# ncu --set full, scale 22, every oriented edge forced through ONE light kernel "
                  f"(variant=merge, then variant=gallop); tag {tag}")
            ncu_summary.main(forced)
            sys.stdout = sys.__stdout__
    for name in (f"bench_{tag}.json", f"sweep_variants_{tag}.jsonl", f"configs_{tag}.jsonl"):
        src = os.path.join(ROOT, "gpurun_out", name)
        if os.path.exists(src):
            with open(src) as fi, open(os.path.join(ROOT, "profiles", f"{tag}_{name.replace('_' + tag, '')}"), "w") as fo:
                fo.write(fi.read())
    print("profiles written for", tag)


if __name__ == "__main__":
    main(sys.argv[1])
