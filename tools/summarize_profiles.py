#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the tracked summaries under profiles/.
    python tools/summarize_profiles.py r1 gpurun_out/r1_launches.csv gpurun_out/prof_r1d.ncu-rep"""
import collections, csv, json, re, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
prof = ROOT / "profiles"; prof.mkdir(exist_ok=True)

# ---- launch list (ncu --metrics gpu__time_duration.sum): per-kernel totals and SHARES of the step
rows = [r for r in csv.reader(open(launches)) if len(r) > 5]
hdr, data = None, []
for r in rows:
    if "Kernel Name" in r: hdr = r; continue
    if hdr and len(r) == len(hdr): data.append(dict(zip(hdr, r)))
agg = collections.OrderedDict()
for d in data:
    short = re.sub(r"\(.*", "", d["Kernel Name"]); short = re.sub(r"void |b2s::|\(anonymous namespace\)::", "", short)
    v = float(d["Metric Value"].replace(",", "")); unit = d["Metric Unit"]
    ns = v * 1e3 if unit.startswith("us") else (v if unit.startswith("ns") else v * 1e6)
    a = agg.setdefault(short, [0, 0.0]); a[0] += 1; a[1] += ns
tot = sum(a[1] for a in agg.values())
lines = ["kernel,launches,total_us,share_pct,avg_us"]
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    lines.append(f"\"{k}\",{n},{ns / 1e3:.1f},{100 * ns / tot:.1f},{ns / n / 1e3:.1f}")
(prof / f"{tag}_launch_list_summary.csv").write_text("\n".join(lines) + "\n")
(prof / f"{tag}_launch_list_raw.csv").write_text(open(launches).read())

# ---- full capture: key metrics per kernel + DRAM traffic
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
nb_ = int(sys.argv[4]) if len(sys.argv) > 4 else 4
md = [f"# ncu --set full summary, {tag} (B200, tools/prof_target.py: b{nb_} t15 c10 200x200, K = {48 * nb_} MB)\n",
      "`ncu --set full --clock-control none --import-source on -k regex:fft2_ -s 6 -c 3` (cold-cache, serialised: durations are a little",
      "longer than the CUDA-event numbers of bench.py).  The raw .ncu-rep stays in gpurun_out/ (scratch).\n"]
traffic = {}
f = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    md.append(f"\n## {name[:160]}\n\n| metric | value | unit |\n|---|---|---|")
    for w in want:
        if w in hdr:
            i = hdr.index(w); md.append(f"| {w} | {r[i]} | {units[i]} |")
    t = float(r[hdr.index("dram__bytes_read.sum")]) * f[units[hdr.index("dram__bytes_read.sum")]] + \
        float(r[hdr.index("dram__bytes_write.sum")]) * f[units[hdr.index("dram__bytes_write.sum")]]
    md.append(f"| **dram read + write** | {t / 1e6:.1f} | MB per launch |")
    nb = int(sys.argv[4]) if len(sys.argv) > 4 else 4          # slices per launch of the capture (tools/prof_target.py argument)
    if "EpiDCStage" in name or "EpiKspace<200, 200, 2>" in name:     # one sens_expand + soft-DC call = packed whole-image kernel + half-split tail
        traffic["sens_expand_dc_dram_bytes_per_launch"] = traffic.get("sens_expand_dc_dram_bytes_per_launch", 0.0) + t
        traffic["sens_expand_dc_algorithmic_bytes"] = 104000000 * nb
    if "EpiReduce" in name: traffic.update(sens_reduce_dram_bytes_per_launch=t, sens_reduce_algorithmic_bytes=56000000 * nb)
    if "EpiPlain" in name: traffic.update(fft2c_dram_bytes_per_launch=t, fft2c_algorithmic_bytes=96000000 * nb)
    if "normal_warp" in name: traffic.update(normal_op_dram_bytes_per_launch=t, normal_op_algorithmic_bytes=12800000 * nb)
traffic["source"] = f"ncu --set full, profiles/{tag}_ncu_full_summary.md"
(prof / f"{tag}_ncu_full_summary.md").write_text("\n".join(md) + "\n")
json.dump(traffic, open(prof / f"{tag}_traffic.json", "w"), indent=1)
print("\n".join(lines[:8])); print(json.dumps(traffic))
