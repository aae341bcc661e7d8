"""Turn an ncu report (gpurun_out/*.ncu-rep) into a small committed summary under profiles/.

    python scripts/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1x_name.md [units_per_launch]
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum"]


def ncu(rep, page):
    return subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    lines = [f"# ncu summary of `{rep}`", "", "Captured with `ncu --set full --clock-control none --import-source on` "
             "(times under ncu are cold-cache and serialised; compare shares, not absolutes).", ""]
    rows = list(csv.reader(io.StringIO(ncu(rep, "raw"))))
    h = rows[0]
    ki = h.index("Kernel Name")
    for r in rows[2:]:
        lines.append(f"## {r[ki]}")
        lines.append("")
        lines.append("| metric | value | unit |")
        lines.append("|---|---|---|")
        for k in KEYS:
            if k in h:
                i = h.index(k)
                lines.append(f"| {k} | {r[i]} | {rows[1][i]} |")
        st = sorted(((float(r[i].replace(",", "")), n) for i, n in enumerate(h)
                     if n.startswith("smsp__pcsamp_warps_issue_stalled_") and not n.endswith("_not_issued") and r[i]),
                    reverse=True)[:8]
        lines.append("")
        lines.append("Top warp-stall samples: " + ", ".join(f"{n.replace('smsp__pcsamp_warps_issue_stalled_', '')}={int(v)}" for v, n in st))
        lines.append("")
    # source page: opcode mix + hottest SASS lines per kernel
    rows = list(csv.reader(io.StringIO(ncu(rep, "source"))))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "rows": []}
            secs.append(cur)
        elif cur is not None:
            cur["rows"].append(r)
    seen = set()
    for sec in secs:
        if sec["name"] in seen or not sec["rows"]:
            continue
        seen.add(sec["name"])
        hh = sec["rows"][0]
        si, ni, ei = hh.index("Source"), hh.index("# Samples"), hh.index("Instructions Executed")
        data = [r for r in sec["rows"][1:] if len(r) > ei]
        tot_e = sum(int(r[ei]) for r in data) or 1
        tot_s = sum(int(r[ni]) for r in data) or 1
        op, ops = collections.Counter(), collections.Counter()
        for r in data:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[si])
            if m:
                o = m.group(2).split(".")[0]
                op[o] += int(r[ei]); ops[o] += int(r[ni])
        lines.append(f"### SASS opcode mix — {sec['name'][:70]}")
        lines.append("")
        lines.append(f"warp-instructions executed: {tot_e}; stall samples: {tot_s}")
        lines.append("")
        lines.append("| opcode | % of executed | % of samples |")
        lines.append("|---|---|---|")
        for o, c in op.most_common(16):
            lines.append(f"| {o} | {100 * c / tot_e:.1f} | {100 * ops[o] / tot_s:.1f} |")
        lines.append("")
        lines.append("Hottest SASS lines (samples, executed, instruction):")
        lines.append("")
        for r in sorted(data, key=lambda r: -int(r[ni]))[:8]:
            lines.append(f"- {r[ni]} / {r[ei]} / `{r[si].strip()[:80]}`")
        lines.append("")
    open(out, "w").write("\n".join(lines))
    print("wrote", out)


if __name__ == "__main__":
    main()
