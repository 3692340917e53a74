#!/usr/bin/env python
"""Turn the scratch outputs of a gpurun session (gpurun_out/) into the tracked evidence under profiles/.

    python scripts/summarize_profiles.py r01

  * bench_<workload>.json            -> profiles/<round>_bench_<workload>.json (verbatim bench lines)
  * launches_<workload>.csv          -> profiles/<round>_launches_<workload>.txt (per-kernel launch count / time / share
                                        from the `ncu --metrics gpu__time_duration.sum` pass: cold-cache, serialised)
  * prof_*.ncu-rep                   -> profiles/<round>_ncu_<name>.txt (key metrics per captured launch, via
                                        `ncu -i … --page raw --csv`) and profiles/traffic.json (dram bytes per launch)
  * profiles/README.md               -> the round's table
"""
import csv
import glob
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"), ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("launch__occupancy_limit_shared_mem", "occ limit smem (blocks)"), ("launch__waves_per_multiprocessor", "waves/SM"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor instructions"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared bank conflicts"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
]
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def ncu_raw(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return [], [], []
    return rows[0], rows[1], rows[2:]


def short_name(n):
    n = n.replace("void ", "").replace("b200::", "")
    return n.split("(")[0]


def summarize_rep(rep, round_tag, traffic):
    hdr, units, rows = ncu_raw(rep)
    if not rows:
        return None
    ix = {h: i for i, h in enumerate(hdr)}
    name = os.path.basename(rep).replace(".ncu-rep", "")
    lines = ["# %s — `ncu --set full --clock-control none` capture (per launch; read with `ncu -i … --page raw --csv`)" % name, ""]
    per_kernel = defaultdict(list)
    for r in rows:
        kname = short_name(r[ix["Kernel Name"]])
        lines.append("## %s" % kname)
        for key, label in KEYS:
            if key in ix and r[ix[key]] != "":
                lines.append("  %-28s %s %s" % (label, r[ix[key]], units[ix[key]]))
        try:
            rd = float(r[ix["dram__bytes_read.sum"]]) * UNIT_SCALE.get(units[ix["dram__bytes_read.sum"]], 1.0)
            wr = float(r[ix["dram__bytes_write.sum"]]) * UNIT_SCALE.get(units[ix["dram__bytes_write.sum"]], 1.0)
            per_kernel[kname].append(rd + wr)
        except (KeyError, ValueError):
            pass
        lines.append("")
    for k, v in per_kernel.items():
        base = k.split("<")[0]
        traffic[base] = {"dram_bytes_per_launch": sum(v) / len(v), "launches_captured": len(v), "source": "%s_ncu_%s.txt" % (round_tag, name)}
    path = os.path.join(PROF, "%s_ncu_%s.txt" % (round_tag, name))
    open(path, "w").write("\n".join(lines))
    return path


def summarize_launches(path, round_tag):
    rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
    if not rows:
        return None
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit.startswith("ns") or unit == "nsecond" else (v if unit.startswith("us") else v * 1e3)
        k = short_name(r[ix["Kernel Name"]])
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += us
        total += us
    name = os.path.basename(path).replace(".csv", "")
    out = ["# %s — ncu launch list (gpu__time_duration.sum, --clock-control none): cold-cache, serialised; compare SHARES" % name,
           "%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total_us", "avg_us", "share")]
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-60s %8d %12.1f %10.2f %6.1f%%" % (k[:60], n, us, us / n, 100 * us / total))
    p = os.path.join(PROF, "%s_%s.txt" % (round_tag, name))
    open(p, "w").write("\n".join(out) + "\n")
    return p


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    traffic_path = os.path.join(PROF, "traffic.json")
    try:
        traffic = json.load(open(traffic_path))
    except (OSError, ValueError):
        traffic = {}
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        p = summarize_rep(rep, tag, traffic)
        print("ncu ->", p)
    json.dump(traffic, open(traffic_path, "w"), indent=1, sort_keys=True)
    for path in sorted(glob.glob(os.path.join(OUT, "launches_*.csv"))):
        print("launches ->", summarize_launches(path, tag))
    table = []
    for path in sorted(glob.glob(os.path.join(OUT, "bench_*.json"))):
        if os.path.basename(path) in ("bench_default.json", "bench_ref.json"):
            continue
        try:
            line = json.loads(open(path).read().strip().splitlines()[-1])
        except (ValueError, IndexError):
            continue
        name = os.path.basename(path)
        json.dump(line, open(os.path.join(PROF, "%s_%s" % (tag, name)), "w"), indent=1)
        r = line.get("roofline") or {}
        cb = line.get("cpu_baseline") or {}
        table.append((name.replace("bench_", "").replace(".json", ""), line["n_gpus"], line["value"], line["unit"], line["ms_per_step"],
                      line["e2e"]["value"], r.get("kernel"), r.get("bound"), r.get("achieved"), r.get("unit"), r.get("frac"),
                      r.get("share_of_step"), cb.get("value")))
    write_readme(tag, table)
    return table


def write_readme(tag, table):
    lines = ["# profiles/ — measured evidence (round %s)" % tag.lstrip("r"), "",
             "All numbers: one B200 box through `gpurun`, CUDA-event timing on the library's stream, inputs larger than L2,",
             "peaks from `MEASURED_PEAKS.json` (HBM 6452.8 GB/s measured copy; TF32 tensor = measured sustained cuBLAS bf16 / 2;",
             "BF16 = sustained cuBLAS bf16; 3xTF32 = TF32 / 3; fp32 = the FFMA stream measured live by csrc/ubench.cu, 72.5 TFLOP/s).",
             "`value` = inputs resident in HBM, `e2e` = host-buffer call, `cpu` = the CPU arm of that line (cores and kind are in the JSON).",
             "Files: `%s_bench_<workload>[_gN].json` full bench lines; `%s_launches_<workload>.txt` ncu launch lists;" % (tag, tag),
             "`%s_ncu_<capture>.txt` per-launch metrics of the `ncu --set full` captures; `traffic.json` DRAM bytes per launch;" % tag,
             "`%s_learning_curve_*.json` Pendulum return curves (`scripts/train_pendulum.py`)." % tag, "",
             "| workload | GPUs | value | unit | ms/step | e2e | dominant kernel | bound | achieved | frac of peak | share of step | CPU (1 core) |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for (name, n, v, unit, ms, e2e, kern, bound, ach, aunit, frac, share, cpu) in table:
        lines.append("| %s | %d | %.4g | %s | %.3f | %.4g | %s | %s | %s | %s | %s | %s |" % (
            name, n, v, unit, ms, e2e, kern, bound, ("%.4g %s" % (ach, aunit)) if ach else "-", ("%.3f" % frac) if frac else "-",
            ("%.2f" % share) if share else "-", ("%.4g" % cpu) if cpu else "-"))
    lines += ["", "Round-to-round comparison, microbenchmarks, phase timelines and the experiments that were measured and not kept: `NOTES.md`.",
              "SASS evidence (tcgen05 / TMA mnemonics per kernel): `%s_sass_mnemonics.txt`.  Multi-GPU equivalence logs: `%s_dist_equivalence_*gpu.txt`." % (tag, tag)]
    open(os.path.join(PROF, "README.md"), "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    for row in main():
        print(row)
