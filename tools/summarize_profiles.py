"""Turns the captures of tools/make_profiles.sh (gpurun_out/) into the tracked summaries under profiles/.
    python tools/summarize_profiles.py [tag]      (default tag: r01)
"""
import collections, csv, json, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out", ROOT / "profiles"
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"


def ncu(*args):
    return subprocess.run(["ncu", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


# ---- launch list -> shares
src = OUT / "launches_hd1080_60frames.csv"
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
h = rows[0]
ik, iv = h.index("Kernel Name"), h.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    name = r[ik].split("(")[0].replace("dmf::", "").replace("void ", "").split("<")[0]
    if not name or "render_kernel" in r[ik]:
        continue
    per.setdefault(name, []).append(float(r[iv].replace(",", "")) / 1e3)
(PROF / f"{tag}_launches_hd1080_60frames.csv").write_text(src.read_text())
tot = sum(sum(v) for k, v in per.items() if k.endswith("_kernel") and k not in ("ref_stats_kernel", "ref_expand_kernel", "fill_state_kernel"))
lines = [f"# ncu launch list summary — {tag}, hd_1920x1080, frames 1..60 (ncu --metrics gpu__time_duration.sum --clock-control none)",
         "# per-launch times are cold-cache and serialised; compare SHARES with bench.py's kernel_ms_per_step",
         "kernel,launches,total_us,share,first8_us,last4_us"]
for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
    if len(v) < 30:
        continue
    lines.append(f"{k},{len(v)},{sum(v):.1f},{sum(v) / tot:.3f},{' '.join(f'{x:.0f}' for x in v[:8])},{' '.join(f'{x:.0f}' for x in v[-4:])}")
(PROF / f"{tag}_launch_shares_hd1080.txt").write_text("\n".join(lines) + "\n")
print("\n".join(lines))

# ---- ncc_kernel: details, key metrics, op mix, traffic
rep = OUT / "ncc_kernel_hd1080_frame40.ncu-rep"
(PROF / f"{tag}_ncu_details_ncc_kernel_hd1080_frame40.txt").write_text(ncu("-i", str(rep), "--page", "details"))
raw = list(csv.reader(ncu("-i", str(rep), "--page", "raw", "--csv").splitlines()))
names, units, vals = raw[0], raw[1], raw[2]
d = {n: {"value": v, "unit": u} for n, u, v in zip(names, units, vals)}
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__inst_issued.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]
(PROF / f"{tag}_ncu_key_metrics_ncc_kernel_hd1080_frame40.json").write_text(json.dumps({k: d[k] for k in keys if k in d}, indent=1) + "\n")
srcpage = ncu("-i", str(rep), "--page", "source", "--csv")
op = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_opmix.py")], input=srcpage, stdout=subprocess.PIPE, text=True).stdout
(PROF / f"{tag}_ncu_opmix_ncc_kernel_hd1080_frame40.txt").write_text(op)


def to_bytes(e):
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[e["unit"]]
    return int(round(float(e["value"].replace(",", "")) * mult))


mufu = [l for l in op.splitlines() if "MUFU.RSQ64H" in l]
traffic = {"kernel": "dmf::ncc_kernel", "workload": "hd_1920x1080",
           "capture": "ncu --set full --clock-control none -k regex:ncc_kernel -s 39 -c 1 python tools/profile_run.py hd_1920x1080 42 (launch of frame 40)",
           "dram_bytes_read": to_bytes(d["dram__bytes_read.sum"]), "dram_bytes_write": to_bytes(d["dram__bytes_write.sum"]),
           "gpu_time_duration_us": float(d["gpu__time_duration.sum"]["value"].replace(",", "")),
           "algorithmic_l1_bytes_per_eval": 64 + 68}
(PROF / f"{tag}_ncc_traffic.json").write_text(json.dumps(traffic, indent=1) + "\n")
print(json.dumps(traffic, indent=1))

# ---- the other kernels
rep = OUT / "aux_kernels_hd1080_frame40.ncu-rep"
(PROF / f"{tag}_ncu_details_aux_kernels_hd1080_frame40.txt").write_text(ncu("-i", str(rep), "--page", "details"))
print("profiles written with tag", tag)
