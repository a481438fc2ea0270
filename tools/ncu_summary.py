#!/usr/bin/env python
"""Condense one `ncu --set full --import-source on` capture into the JSON summary kept under profiles/.

    ncu -i X.ncu-rep --page raw --csv > X.raw.csv
    ncu -i X.ncu-rep --page source --csv > X.source.csv        (optional: instruction mix + SASS excerpt)
    python tools/ncu_summary.py X.raw.csv [X.source.csv] --evals N [--sass-out profiles/X_sass.txt] > profiles/X.json

--evals: Cooper-Frye evaluations the captured launch covers (cells x species x pT x phi x slots); turns instruction counts into
per-evaluation figures.  The SASS excerpt lists the instructions that account for the top 90 % of executed warp instructions
(the steady-state inner loop) in address order, with their share."""
import argparse
import csv
import json
import re
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpc__cycles_elapsed.avg.per_second", "sm__cycles_elapsed.max",
    "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_cbu.avg.pct_of_peak_sustained_active",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
]


def read_raw(path):
    rows = list(csv.reader(open(path, newline="")))
    # header row, units row, then one row per captured launch
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, vals = rows[hdr], rows[hdr + 1], rows[hdr + 2]
    out = {}
    for n, u, v in zip(names, units, vals):
        out[n] = (v, u)
    return out


def num(v):
    try:
        return float(v.replace(",", ""))
    except ValueError:
        return v


SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "msecond": 1e6, "usecond": 1e3, "second": 1e9, "nsecond": 1.0, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("source", nargs="?")
    ap.add_argument("--evals", type=float, default=0.0)
    ap.add_argument("--sass-out")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = read_raw(a.raw)
    out = {"kernel": raw.get("Kernel Name", ("?", ""))[0], "grid": raw.get("Grid Size", ("", ""))[0], "block": raw.get("Block Size", ("", ""))[0]}
    if a.note:
        out["note"] = a.note
    for k in KEEP:
        if k in raw:
            v, u = raw[k]
            x = num(v)
            if isinstance(x, float) and u in SCALE and (k.startswith("dram__bytes") or k.startswith("lts__t_bytes") or k == "gpu__time_duration.sum"):
                x *= SCALE[u]          # bytes / ns
            out[k] = x
    for k, (v, u) in raw.items():
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and "not_issued" not in k:
            out[k] = num(v)
    if a.evals:
        out["evaluations"] = a.evals
        if isinstance(out.get("smsp__inst_executed.sum"), float):
            out["warp_instr_per_warp_evaluation"] = out["smsp__inst_executed.sum"] * 32.0 / a.evals
        t = out.get("gpu__time_duration.sum")
        if isinstance(t, float) and t > 0:
            out["evaluations_per_s"] = a.evals / (t * 1e-9)
    if a.source:
        rows = list(csv.reader(open(a.source, newline="")))
        hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        names = rows[hdr]
        i_src, i_exec = names.index("Source"), names.index("Instructions Executed")
        i_stall = names.index("Warp Stall Sampling (All Samples)") if "Warp Stall Sampling (All Samples)" in names else None
        inst = []
        for r in rows[hdr + 1:]:
            if len(r) <= i_exec:
                continue
            try:
                n = float(r[i_exec])
            except ValueError:
                continue
            inst.append((r[0], r[i_src].strip(), n, float(r[i_stall]) if i_stall is not None and r[i_stall] else 0.0))
        total = sum(n for _, _, n, _ in inst) or 1.0
        mix = {}
        for _, src, n, _ in inst:
            m = re.match(r"(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", src)
            op = m.group(1) if m else "?"
            mix[op] = mix.get(op, 0.0) + n
        fp64 = sum(v for k, v in mix.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
        out["opcode_mix_warp_instr"] = {k: v for k, v in sorted(mix.items(), key=lambda kv: -kv[1])[:24]}
        out["fp64_share_of_warp_instr"] = fp64 / total
        if a.evals:
            out["fp64_instr_per_evaluation"] = fp64 * 32.0 / a.evals
            out["other_instr_per_evaluation"] = (total - fp64) * 32.0 / a.evals
        if a.sass_out:
            order = sorted(range(len(inst)), key=lambda i: -inst[i][2])
            keep, acc = set(), 0.0
            for i in order:
                keep.add(i); acc += inst[i][2]
                if acc >= 0.9 * total:
                    break
            stall_total = sum(s for _, _, _, s in inst) or 1.0
            with open(a.sass_out, "w") as f:
                f.write("# %s\n# instructions covering 90 %% of the executed warp instructions, in address order\n" % out["kernel"])
                f.write("# address  share-of-executed  share-of-stall-samples  SASS\n")
                for i in sorted(keep):
                    adr, src, n, s = inst[i]
                    f.write("%s  %6.3f%%  %6.3f%%  %s\n" % (adr, 100 * n / total, 100 * s / stall_total, src))
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main()
