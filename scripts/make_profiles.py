#!/usr/bin/env python
"""Build profiles/<tag>_* summaries from an evidence run (scripts/gpu_baseline.sh) in gpurun_out/.

  python scripts/make_profiles.py r40 r1
copies the small logs, writes per-kernel ncu metric tables (markdown) and the roofline traffic JSON bench.py reads.
"""
import collections, csv, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
tag, out = sys.argv[1], sys.argv[2]
os.makedirs(DST, exist_ok=True)


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("b200dm::", "").replace("__nv_bfloat16", "bf16").replace("(int)", "").replace("(bool)", "")


METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma pipe %"),
    ("dram__bytes_read.sum", "DRAM read MB"),
    ("dram__bytes_write.sum", "DRAM write MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def raw_table(name):
    p = os.path.join(SRC, f"{tag}_{name}_raw.csv")
    if not os.path.isfile(p):
        return [], []
    rows = list(csv.reader(open(p, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m, _ in METRICS if m in hdr}
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    recs = []
    for r in rows[2:]:
        rec = {"kernel": short(r[kn]), "grid": r[gs]}
        for m, label in METRICS:
            if m not in idx:
                continue
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                continue
            u = units[idx[m]]
            if u in UNIT:
                v *= UNIT[u]
            elif u in ("ns", "nsecond"):
                v /= 1e3
            elif u in ("ms", "msecond"):
                v *= 1e3
            rec[label] = round(v, 2)
        recs.append(rec)
    return recs, [l for m, l in METRICS if m in idx]


lines = [f"# Round-1 profiles (evidence run `{tag}`, scripts/gpu_baseline.sh)\n"]
traffic = {}
for name, title in (("conv", "tcgen05 forward / data-gradient convs (first 12 launches of a training step)"),
                    ("wgrad", "weight-gradient kernels (first 8 launches of the backward pass)"),
                    ("hbmk", "HBM-bound kernels (first 14 launches)"),
                    ("hbmk_ddim", "HBM-bound kernels at the DDIM shape (B=256, 64x64; first level of an evaluation)")):
    recs, labels = raw_table(name)
    if not recs:
        continue
    lines.append(f"\n## ncu --set full: {title}\n")
    lines.append("| kernel | grid | " + " | ".join(labels) + " |")
    lines.append("|---|---|" + "---|" * len(labels))
    for r in recs:
        lines.append(f"| `{r['kernel']}` | {r['grid']} | " + " | ".join(str(r.get(l, "")) for l in labels) + " |")
        if name == "hbmk_ddim":
            continue                                # the roofline traffic JSON describes the training step
        k = r["kernel"].split("<")[0]
        t = traffic.setdefault(k, {"launches": 0, "dram_MB": 0.0, "us": 0.0})
        t["launches"] += 1
        t["dram_MB"] += r.get("DRAM read MB", 0.0) + r.get("DRAM write MB", 0.0)
        t["us"] += r.get("us", 0.0)
for k, t in traffic.items():
    t["dram_bytes_per_launch"] = round(t["dram_MB"] * 1e6 / t["launches"])
json.dump(traffic, open(os.path.join(DST, f"{out}_roofline_traffic.json"), "w"), indent=1)

# launch list summary
p = os.path.join(SRC, f"{tag}_launches.csv")
if os.path.isfile(p):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_launch_summary.py"), p, "--md"],
                         capture_output=True, text=True).stdout
    lines.append("\n## ncu launch list of one eager training step (`--metrics gpu__time_duration.sum`)\n")
    lines.append("Serialised and cold-cache under ncu: compare SHARES with bench.py's CUDA-event table, not absolutes.\n")
    lines.append(txt)
    shutil.copy(p, os.path.join(DST, f"{out}_train_step_launches_ncu.csv"))

for f, note in ((f"{tag}_bench_train.log", "bench.py (train, N=1)"), (f"{tag}_bench_ddim.log", "bench.py --workload ddim"),
                (f"{tag}_bench_reference.log", "bench.py --impl reference"), (f"{tag}_hbm.log", "scripts/hbm_microbench.py"),
                (f"{tag}_umma_rate.log", "scripts/umma_rate.py"), (f"{tag}_conv_micro.log", "scripts/conv_microbench.py"),
                (f"{tag}_kernels.json", "bench.py --profile-out (per-launch CUDA-event times, training step)"),
                (f"{tag}_kernels_ddim.json", "bench.py --profile-out (DDIM evaluation)"),
                (f"{tag}_bench_ddpm.log", "bench.py --workload ddpm (configs[3]: 1000-step ancestral sampling, batch 1024)"),
                (f"{tag}_bench_train64.log", "bench.py --workload train64 (configs[4] per GPU: 3x64x64, 64 images)"),
                (f"{tag}_kernels_train64.json", "bench.py --profile-out (training step at 3x64x64)"),
                (f"{tag}_opt_cost.log", "scripts/opt_cost.py (optimiser tail inside the loop)")):
    if os.path.isfile(os.path.join(SRC, f)):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f.replace(tag, out)))
        lines.append(f"* `{f.replace(tag, out)}` — {note}")
for name in ("conv", "wgrad", "hbmk", "hbmk_ddim"):
    f = os.path.join(SRC, f"{tag}_{name}_raw.csv")
    if os.path.isfile(f):
        shutil.copy(f, os.path.join(DST, f"{out}_{name}_ncu_raw.csv"))
open(os.path.join(DST, f"{out}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:6000])
