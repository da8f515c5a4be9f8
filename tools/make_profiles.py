#!/usr/bin/env python
"""Turns the artefacts of the last measurement run (tools_gpu_profiles_r02.sh -> gpurun_out/) into the tracked summaries
under profiles/ (TAG = r02):
  profiles/TAG_launches*.md   per-kernel share of one step (ncu --metrics gpu__time_duration.sum launch lists)
  profiles/TAG_kernels.md     ncu --set full counters of the dominant kernels (in-step and at the stress shape)
  profiles/TAG_sass.txt       per-kernel counts of the tcgen05 / TMEM / TMA SASS mnemonics (cuobjdump of the shipped .so)
  profiles/traffic.json       per-launch DRAM bytes (read + write) that bench.py reports as roofline.traffic
  + the bench lines, the GPU test log, the micro-benchmark log and the bf16 parity record, copied as they are
Usage: python tools/make_profiles.py [TAG]"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    res = []
    for r in rows[2:]:
        d = {"name": r[idx["Kernel Name"]]}
        for w in WANT:
            if w in idx:
                d[w] = (r[idx[w]], units[idx[w]])
        res.append(d)
    return res


def to_bytes(v, unit):
    f = float(v.replace(",", ""))
    return f * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def launches_md(src="launches.csv", dst=TAG + "_launches.md", title="phase-2 step (current kernels)",
                cmd="python bench.py --steps 2 --warmup 1 --no-graph --no-cpu-baseline --no-stress --no-train"):
    if not os.path.exists(os.path.join(G, src)):
        return
    rows = [r for r in csv.reader(open(os.path.join(G, src))) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[i_val].replace(",", ""))
        except ValueError:
            continue
        a = agg.setdefault(r[i_name][:100], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    lines = [f"# {TAG} - ncu launch list of the {title}", "",
             f"Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv {cmd}` (cold-cache, "
             "serialised launches: compare SHARES, not absolutes; the capture spans several eager steps).", "",
             "| kernel | launches | total ns | avg ns | share |", "|---|---:|---:|---:|---:|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {a[0]} | {a[1]:.0f} | {a[1] / a[0]:.0f} | {100 * a[1] / tot:.1f}% |")
    open(os.path.join(P, dst), "w").write("\n".join(lines) + "\n")


def kernels_md():
    traffic = {}
    lines = [f"# {TAG} - ncu --set full counters of the dominant kernels", "",
             "`--clock-control none --import-source on`; per-launch values.  In-step captures come from "
             "`bench.py --no-graph` (K = 5000 / 5400 RoIs), the stress capture from `tools/prof_roi.py` "
             "(96 000 RoIs = config #4, one image).", ""]
    for title, rep, keys in [("In-step kernels (bench workload)", "prof_step.ncu-rep", None),
                             ("RoIAlign at the stress shape (96 000 RoIs, 2.4 GB bf16 out)", "prof_roi_stress.ncu-rep", None),
                             ("RoIAlignRotated on the TMA + mma path (OBB config #3, 5000 / 5400 RoIs)", "prof_obb_roi.ncu-rep", None),
                             ("RoIAlign backward (training step, 5000 RoIs)", "prof_roi_bwd.ncu-rep", None)]:
        path = os.path.join(G, rep)
        if not os.path.exists(path):
            continue
        lines += [f"## {title}", ""]
        for d in raw(path):
            nm = d["name"][:110]
            lines += [f"### `{nm}`", "", "| counter | value |", "|---|---:|"]
            for w in WANT:
                if w in d:
                    lines.append(f"| {w} | {d[w][0]} {d[w][1]} |")
            lines.append("")
            rd = to_bytes(*d["dram__bytes_read.sum"])
            wr = to_bytes(*d["dram__bytes_write.sum"])
            dur = float(d["gpu__time_duration.sum"][0].replace(",", ""))
            if "fc_gemm" in nm and dur > 80:
                traffic.setdefault("fc_gemm_kernel@fc1", rd + wr)
            if "roi_align_mma" in nm and "obb" in rep:
                traffic.setdefault("roi_align_mma_kernel_rot@5000", rd + wr)
            elif "roi_align_mma" in nm:
                traffic.setdefault("roi_align_mma_kernel@96000" if "stress" in rep else "roi_align_mma_kernel@5000", rd + wr)
    open(os.path.join(P, TAG + "_kernels.md"), "w").write("\n".join(lines) + "\n")
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    return traffic


def sass_txt():
    """Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA (B200_PROFILING.md)."""
    so = os.path.join(ROOT, "point_teacher_b200", "_C", "libptb200.so")
    out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    pats = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "HMMA.16816", "SYNCS", "REDG", "RED.E"]
    cur, table = None, collections.OrderedDict()
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()[:110]
            table[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in pats:
            if re.search(r"\b" + re.escape(p) + r"(\.|\b)", line):
                if p == "UTCHMMA" and "UTCHMMA.2CTA" in line:
                    continue
                table[cur][p] += 1
    lines = [f"# {TAG} - SASS mnemonic counts per kernel (cuobjdump -sass point_teacher_b200/_C/libptb200.so, sm_100a)",
             "# UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld (TMEM), UTMALDG / UTMASTG = TMA tensor load / store, UTCBAR = tcgen05.commit,",
             "# UTCATOMSWS = TMEM allocation, HMMA.16816 = mma.sync.m16n8k16 (RoIAlign interpolation, small heads)", ""]
    for k, c in table.items():
        if sum(c.values()):
            lines.append(f"{k}\n    " + "  ".join(f"{p}={c[p]}" for p in pats if c[p]))
    open(os.path.join(P, TAG + "_sass.txt"), "w").write("\n".join(lines) + "\n")


def copy_logs():
    for f in os.listdir(G):
        if f.startswith(TAG + "_") and f.endswith((".json", ".log")):
            shutil.copy(os.path.join(G, f), os.path.join(P, f))
    if os.path.exists(os.path.join(G, "parity_r02.json")):
        shutil.copy(os.path.join(G, "parity_r02.json"), os.path.join(P, TAG + "_parity.json"))


if __name__ == "__main__":
    os.makedirs(P, exist_ok=True)
    launches_md()
    launches_md("obb_launches.csv", TAG + "_launches_obb.md", "OBB (config #3) phase-2 step", "python tools/prof_obb.py 3")
    launches_md("train_launches.csv", TAG + "_launches_train.md", "HBB training step (forward + backward)",
                "python tools/prof_train.py 3")
    sass_txt()
    copy_logs()
    print(kernels_md())
