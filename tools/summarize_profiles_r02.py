"""Turns the captures of tools/make_profiles_r02.sh (gpurun_out/r02/) into the tracked summaries under profiles/:
    r02_launch_shares_hd1080.txt, r02_ncc_pipe_mix.json (read by bench.py), r02_ncu_details_*.txt, r02_ncu_opmix_*.txt
"""
import collections, csv, json, re, subprocess, sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
OUT, PROF = ROOT / "gpurun_out" / "r02", ROOT / "profiles"


def ncu(*args):
    return subprocess.run(["ncu", *args], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def num(s):
    return float(s.replace(",", ""))


# ---- launch list -> shares
src = OUT / "launches_hd1080_60frames.csv"
if src.exists():
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    h = rows[0]
    ik, iv = h.index("Kernel Name"), h.index("Metric Value")
    per = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0].replace("dmf::", "").replace("void ", "").split("<")[0]
        if not name or "render_kernel" in r[ik]:
            continue
        per.setdefault(name, []).append(num(r[iv]) / 1e3)
    (PROF / "r02_launches_hd1080_60frames.csv").write_text(src.read_text())
    tot = sum(sum(v) for k, v in per.items() if k.endswith("_kernel") and k not in ("ref_stats_kernel", "ref_expand_kernel", "fill_state_kernel"))
    lines = ["# ncu launch list summary — r02, hd_1920x1080, updates 1..60 (ncu --metrics gpu__time_duration.sum --clock-control none)",
             "# per-launch times are cold-cache and serialised; compare SHARES with bench.py's kernel_ms_per_step",
             "kernel,launches,total_us,share,first8_us,last4_us"]
    for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
        if len(v) < 30:
            continue
        lines.append(f"{k},{len(v)},{sum(v):.1f},{sum(v) / tot:.3f},{' '.join(f'{x:.0f}' for x in v[:8])},{' '.join(f'{x:.0f}' for x in v[-4:])}")
    (PROF / "r02_launch_shares_hd1080.txt").write_text("\n".join(lines) + "\n")
    print("\n".join(lines))

# ---- pipe mix of one ncc_kernel launch per workload
mix = {}
for wl in ("hd_1920x1080", "uhd_3840x2160"):
    f, lg = OUT / f"ncc_pipes_{wl}.csv", OUT / f"ncc_pipes_{wl}.log"
    if not f.exists() or not lg.exists():
        continue
    m = re.search(r"update_counters (\{.*\})", lg.read_text())
    if not m:
        continue
    cnt = json.loads(m.group(1))
    rows = [r for r in csv.reader(open(f)) if len(r) > 5]
    h = rows[0]
    im, iu, iv = h.index("Metric Name"), h.index("Metric Unit"), h.index("Metric Value")
    d = {r[im]: (num(r[iv]), r[iu]) for r in rows[1:]}
    def byt(k):
        v, u = d[k]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    ev = cnt["ncc_evals"]
    mix[wl] = {
        "launch": f"update {cnt['update']} of {wl} (ncu --metrics ... -k regex:ncc_kernel -s 39 -c 1 python tools/profile_run.py {wl} 42 40)",
        "ncc_evals": ev, "active_px": cnt["active"],
        "inst_executed_per_eval": d["smsp__inst_executed.sum"][0] / ev,
        "pipe_fmaheavy_per_eval": d["sm__inst_executed_pipe_fmaheavy.sum"][0] / ev,
        "pipe_fp64_per_eval": d["sm__inst_executed_pipe_fp64.sum"][0] / ev,
        "pipe_xu_per_eval": d["sm__inst_executed_pipe_xu.sum"][0] / ev,
        "pipe_alu_per_eval": d["sm__inst_executed_pipe_alu.sum"][0] / ev,
        "pipe_lsu_per_eval": d["sm__inst_executed_pipe_lsu.sum"][0] / ev,
        "lsu_wavefronts_per_eval": d["l1tex__data_pipe_lsu_wavefronts.sum"][0] / ev,
        "dram_bytes_per_launch": byt("dram__bytes_read.sum") + byt("dram__bytes_write.sum"),
        "under_ncu": {"duration_us": d["gpu__time_duration.sum"][0] * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(d["gpu__time_duration.sum"][1], 1.0), "lsu_wavefronts_pct_of_peak": d["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][0],
                      "issue_active_pct": d["sm__inst_issued.avg.pct_of_peak_sustained_active"][0], "warps_active_pct": d["sm__warps_active.avg.pct_of_peak_sustained_active"][0],
                      "l1_hit_pct": d["l1tex__t_sector_hit_rate.pct"][0], "l2_hit_pct": d["lts__t_sector_hit_rate.pct"][0],
                      "registers_per_thread": d["launch__registers_per_thread"][0]},
        "note": "warp-level instruction counts per NCC evaluation (x32 for thread-ops); the per-evaluation figures include the unit "
                "prologue (record + reference patch) amortised over the unit's samples",
    }
if mix:
    (PROF / "r02_ncc_pipe_mix.json").write_text(json.dumps(mix, indent=1) + "\n")
    print(json.dumps(mix, indent=1))

# ---- full captures: details + op mix
for tag, name in (("ncc_kernel_hd1080_update40", "ncc_kernel_hd1080_update40"), ("ncc_kernel_uhd2160_update40", "ncc_kernel_uhd2160_update40"),
                  ("aux_kernels_hd1080_update40", "aux_kernels_hd1080_update40")):
    rep = OUT / f"{tag}.ncu-rep"
    if not rep.exists():
        continue
    (PROF / f"r02_ncu_details_{name}.txt").write_text(ncu("-i", str(rep), "--page", "details"))
    if "uhd" in tag:
        continue
    srcpage = ncu("-i", str(rep), "--page", "source", "--csv")
    if srcpage.strip():
        op = subprocess.run([sys.executable, str(ROOT / "tools" / "ncu_opmix.py")], input=srcpage, stdout=subprocess.PIPE, text=True).stdout
        (PROF / f"r02_ncu_opmix_{name}.txt").write_text(op)
print("profiles written")
