"""Summarises an ncu report of integrate_lanes: headline counters, stall reasons, and the source page aggregated by
source line (executed warp instructions, stall samples, opcode mix).
usage: ncu_summary.py report.ncu-rep [n_rhs]     n_rhs = RHS evaluations of the workload per WARP-member unit, optional"""
import collections, csv, io, subprocess, sys


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    raw = page(rep, "raw")
    hdr, units, vals = raw[0], raw[1], raw[2]
    kv = dict(zip(hdr, vals))
    for k in ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active",
              "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
              "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio"]:
        if k in kv:
            print(f"{k} = {kv[k]}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            v = float(kv[k] or 0)
            if v >= 0.05:
                print(f"stall {k.split('issue_stalled_')[1].split('_per_issue')[0]} = {v:.3f}")
    rows = page(rep, "source", ("--print-source", "sass,cuda"))
    blocks, cur = [], None
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur = {"file": r[1].split("/")[-1], "rows": []}
            blocks.append(cur)
        elif len(r) > 5 and r[0] == "Line No":
            cur["hdr"] = r
        elif cur is not None and len(r) > 5:
            cur["rows"].append(r)
    total_exec = 0
    lines = collections.OrderedDict()
    ops = collections.Counter()
    stall_cols = None
    # the page lists every source line followed by the SASS instructions generated for it
    for b in blocks:
        h = b["hdr"]
        ci = {}
        for i, name in enumerate(h):
            ci.setdefault(name, i)
        i_exec, i_smp = ci["Instructions Executed"], ci["# Samples"]
        if stall_cols is None:
            stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        cur_line, cur_src = "?", ""
        for r in b["rows"]:
            if r[0] != "":                       # a source line (its counters are the sums of the SASS rows below)
                cur_line, cur_src = r[0], r[1]
                continue
            if r[2] in ("", "...") or not r[2].startswith("0x"):
                continue
            try:
                ex = float(r[i_exec] or 0); smp = float(r[i_smp] or 0)
            except ValueError:
                continue
            key = (b["file"], cur_line)
            e = lines.setdefault(key, {"src": cur_src, "static": 0, "exec": 0.0, "smp": 0.0, "ops": collections.Counter(),
                                       "st": collections.Counter()})
            e["static"] += 1; e["exec"] += ex; e["smp"] += smp
            toks = r[3].split()
            op = toks[0] if toks else "?"
            if op.startswith("@") and len(toks) > 1:
                op = toks[1]
            op = op.split(".")[0]
            e["ops"][op] += ex; ops[op] += ex
            for sc in stall_cols:
                try:
                    e["st"][sc] += float(r[ci[sc]] or 0)
                except ValueError:
                    pass
            total_exec += ex
    tot_smp = sum(e["smp"] for e in lines.values())
    n_rhs = float(sys.argv[2]) if len(sys.argv) > 2 else None
    scale = (1.0 / n_rhs) if n_rhs else 100.0 / total_exec
    unit = "dyn/RHS" if n_rhs else "% of executed"
    print(f"\ntotal executed warp instructions {total_exec:.4g}; samples {tot_smp:.0f}")
    per_file = collections.Counter(); per_file_smp = collections.Counter()
    for (f, ln), e in lines.items():
        per_file[f] += e["exec"]; per_file_smp[f] += e["smp"]
    print("per file (" + unit + " / % samples): " + ", ".join(f"{f} {v * scale:.1f} / {per_file_smp[f] / tot_smp * 100:.1f}%" for f, v in per_file.most_common()))
    print("opcode mix (" + unit + "): " + ", ".join(f"{k} {v * scale:.1f}" for k, v in ops.most_common(32)))
    print(f"\n--- top 60 lines by executed instructions ({unit}; smp = % of stall samples; top stall reasons)")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["exec"])[:60]:
        st = ", ".join(f"{k[6:]} {v / max(e['smp'], 1) * 100:.0f}%" for k, v in e["st"].most_common(3))
        print(f"{f:20s}:{ln:>5s} static {e['static']:4d} {e['exec'] * scale:7.2f} smp {e['smp'] / tot_smp * 100:5.2f}% [{st}] | {e['src'].strip()[:90]}")
    print(f"\n--- top 40 lines by stall samples")
    for (f, ln), e in sorted(lines.items(), key=lambda kv: -kv[1]["smp"])[:40]:
        st = ", ".join(f"{k[6:]} {v / max(e['smp'], 1) * 100:.0f}%" for k, v in e["st"].most_common(3))
        topops = ", ".join(f"{k} {v * scale:.1f}" for k, v in e["ops"].most_common(4))
        print(f"{f:20s}:{ln:>5s} smp {e['smp'] / tot_smp * 100:5.2f}% exec {e['exec'] * scale:7.2f} [{st}] ({topops}) | {e['src'].strip()[:80]}")


if __name__ == "__main__":
    main()
