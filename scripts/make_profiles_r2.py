#!/usr/bin/env python
"""Build the committed round-2 profile summaries (profiles/r2_*) from the evidence runs in gpurun_out/:

  python scripts/make_profiles_r2.py <profile tag (scripts/gpu_r2_profile.sh)> <evidence tag (scripts/gpu_r2.sh)>

* `ncu --set full` captures (raw pages) -> metric tables + profiles/r2_roofline_traffic.json (what bench.py reports as
  `roofline.traffic`);
* per-launch metric passes (tensor-pipe %, DRAM bytes) of every conv launch of a training step and of every launch of a
  DDIM evaluation -> per-kernel tables;
* the ncu launch list of an eager training step; the bench / micro-benchmark logs.
"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
ptag, etag = sys.argv[1], sys.argv[2]
out = "r2"


def short(n):
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    return n.replace("b200dm::", "").replace("__nv_bfloat16", "bf16")


METRICS = [
    ("gpu__time_duration.sum", "us"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe % (active)"),
    ("dram__bytes_read.sum", "DRAM read MB"),
    ("dram__bytes_write.sum", "DRAM write MB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2->SM MB"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
    ("launch__registers_per_thread", "regs"),
]
UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}


def conv(v, u):
    if u in UNIT:
        return v * UNIT[u]
    if u in ("ns", "nsecond"):
        return v / 1e3
    if u in ("ms", "msecond"):
        return v * 1e3
    return v


def raw_table(path):
    if not os.path.isfile(path):
        return [], []
    rows = list(csv.reader(open(path, errors="replace")))
    hdr, units = rows[0], rows[1]
    idx = {m: hdr.index(m) for m, _ in METRICS if m in hdr}
    kn, gs = hdr.index("Kernel Name"), hdr.index("Grid Size")
    recs = []
    for r in rows[2:]:
        rec = {"kernel": short(r[kn]), "grid": r[gs]}
        for m, label in METRICS:
            if m in idx:
                try:
                    rec[label] = round(conv(float(r[idx[m]].replace(",", "")), units[idx[m]]), 2)
                except ValueError:
                    pass
        recs.append(rec)
    return recs, [l for m, l in METRICS if m in idx]


def long_table(path):
    """`ncu --metrics ... --csv --log-file`: one row per (launch, metric)."""
    if not os.path.isfile(path):
        return []
    rows = list(csv.reader(open(path, errors="replace")))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Name" in r]
    if not h:
        return []
    hdr = rows[h[0]]
    kn, mn, mu, mv, idc = (hdr.index(c) for c in ("Kernel Name", "Metric Name", "Metric Unit", "Metric Value", "ID"))
    per = collections.OrderedDict()
    for r in rows[h[0] + 1:]:
        if len(r) <= mv:
            continue
        d = per.setdefault(r[idc], {"kernel": short(r[kn])})
        try:
            d[r[mn]] = conv(float(r[mv].replace(",", "")), r[mu])
        except ValueError:
            pass
    return list(per.values())


def family_table(recs, title):
    agg = collections.OrderedDict()
    for d in recs:
        agg.setdefault(d["kernel"], []).append(d)
    lines = [f"\n## {title}\n", "| kernel | launches | total us | tensor pipe % (time-weighted / max) | DRAM MB per launch (read+write) | DRAM % of peak (time-weighted) |",
             "|---|---|---|---|---|---|"]
    tot = sum(x.get("gpu__time_duration.sum", 0) for x in recs)
    for k, L in sorted(agg.items(), key=lambda kv: -sum(x.get("gpu__time_duration.sum", 0) for x in kv[1])):
        t = sum(x.get("gpu__time_duration.sum", 0) for x in L)
        tp = sum(x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) * x.get("gpu__time_duration.sum", 0) for x in L) / max(t, 1e-9)
        tmax = max(x.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0) for x in L)
        dm = sum(x.get("dram__bytes_read.sum", 0) + x.get("dram__bytes_write.sum", 0) for x in L) / len(L)
        dp = sum(x.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0) * x.get("gpu__time_duration.sum", 0) for x in L) / max(t, 1e-9)
        lines.append(f"| `{k}` | {len(L)} | {t:.1f} ({100 * t / max(tot, 1e-9):.1f} %) | {tp:.1f} / {tmax:.1f} | {dm:.1f} | {dp:.1f} |")
    lines.append(f"\nTotal {tot:.1f} us over {len(recs)} launches (serialised under ncu).")
    return lines


lines = [f"# Round-2 profiles (ncu runs `{ptag}`, evidence run `{etag}`; scripts/gpu_r2_profile.sh, scripts/gpu_r2.sh)\n"]
traffic = {}
for name, title in (("conv_gn_ddim", "fused conv3x3 + GroupNorm + FiLM + SiLU at the DDIM shape (B=256, 64x64, level 0)"),
                    ("halo_train", "3x3 data-gradient / weight-gradient halo kernels of a training step (B=128, 32x32)"),
                    ("hbmk", "HBM-bound kernels of a training step"),
                    ("lablock", "fused LinearAttention block (inference) at the DDIM shape (B=256, 64x64, level 0, C=64)")):
    recs, labels = raw_table(os.path.join(SRC, f"{ptag}_{name}_raw.csv"))
    if not recs:
        continue
    lines.append(f"\n## ncu --set full: {title}\n")
    lines.append("| kernel | grid | " + " | ".join(labels) + " |")
    lines.append("|---|---|" + "---|" * len(labels))
    for r in recs:
        lines.append(f"| `{r['kernel']}` | {r['grid']} | " + " | ".join(str(r.get(l, "")) for l in labels) + " |")
        k = r["kernel"].split("<")[0]
        t = traffic.setdefault(k, {"launches": 0, "dram_MB": 0.0, "us": 0.0, "shape": title})
        t["launches"] += 1
        t["dram_MB"] += r.get("DRAM read MB", 0.0) + r.get("DRAM write MB", 0.0)
        t["us"] += r.get("us", 0.0)
    shutil.copy(os.path.join(SRC, f"{ptag}_{name}_raw.csv"), os.path.join(DST, f"{out}_{name}_ncu_raw.csv"))
for k, t in traffic.items():
    t["dram_bytes_per_launch"] = round(t["dram_MB"] * 1e6 / t["launches"])
json.dump(traffic, open(os.path.join(DST, f"{out}_roofline_traffic.json"), "w"), indent=1)

for f, title in ((f"{ptag}_conv_train_metrics.csv", "every conv / weight-gradient launch of one training step (B=128, 32x32): ncu metrics pass"),
                 (f"{etag}_eval_ddim_metrics.csv", "every launch of one DDIM UNet evaluation (B=256, 64x64): ncu metrics pass")):
    recs = long_table(os.path.join(SRC, f))
    if recs:
        lines += family_table(recs, title)
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f.replace(ptag, out).replace(etag, out)))

p = os.path.join(SRC, f"{ptag}_launches.csv")
if os.path.isfile(p):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_launch_summary.py"), p, "--md"],
                         capture_output=True, text=True).stdout
    lines.append("\n## ncu launch list of one eager training step (`--metrics gpu__time_duration.sum`)\n")
    lines.append("Serialised and cold-cache under ncu: compare SHARES with bench.py's CUDA-event table, not absolutes.\n")
    lines.append(txt)
    shutil.copy(p, os.path.join(DST, f"{out}_train_step_launches_ncu.csv"))

lines.append("\n## Files\n")
for f, note in ((f"{etag}_bench.log", "bench.py as the driver runs it (train + DDIM-50 secondary, CPU and eager-GPU baselines)"),
                (f"{etag}_bench_reference.log", "bench.py --impl reference (the unmodified reference module on the host cores)"),
                (f"{etag}_bench_train64.log", "bench.py --workload train64 (configs[4] per GPU)"),
                (f"{etag}_bench_ddpm.log", "bench.py --workload ddpm (configs[3])"),
                (f"{etag}_bench_nofuse.log", "bench.py with B200DM_FUSE_GN=0 (conv + one-pass norm instead of the fused launch)"),
                (f"{etag}_kernels.json", "per-launch CUDA-event times of one training step"),
                (f"{etag}_kernels_ddim.json", "per-launch CUDA-event times of one DDIM evaluation"),
                (f"{etag}_kernels_train64.json", "per-launch CUDA-event times, training step at 3x64x64"),
                (f"{etag}_hbm.log", "scripts/hbm_microbench.py"), (f"{etag}_umma_rate.log", "scripts/umma_rate.py"),
                (f"{etag}_conv_gn_micro.log", "scripts/conv_microbench.py --what gn (fused launch vs conv + norm)"),
                (f"{etag}_conv_micro.log", "scripts/conv_microbench.py (forward / weight gradient per layer)"),
                (f"{etag}_phase_halo.log", "scripts/phase_timing.py (SM-clock timeline of one conv3x3_halo CTA)"),
                (f"{etag}_phase_gn.log", "scripts/phase_timing_gn.py (timeline of one fused conv+GroupNorm CTA)"),
                (f"{etag}_phase_linattn.log", "scripts/phase_timing_linattn.py (timelines of the fused LinearAttention passes)"),
                (f"{etag}_side_cost.log", "scripts/side_cost.py (what the second-stream kernels cost the training step)"),
                (f"{etag}_bench_ddim_nofuse_linattn.log", "bench.py --workload ddim with B200DM_FUSE_LINATTN=0 (unfused attention chain)"),
                (f"{etag}_parity_report.jsonl", "measured distances of every GPU parity test"),
                (f"{etag}_pytest.log", "pytest -m gpu")):
    if os.path.isfile(os.path.join(SRC, f)):
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f.replace(etag, out)))
        lines.append(f"* `{f.replace(etag, out)}` — {note}")
open(os.path.join(DST, f"{out}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
print("\n".join(lines)[:5000])
