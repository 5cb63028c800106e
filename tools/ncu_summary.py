"""Summarise an Nsight Compute report (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_msda_ncu_full.txt

Per profiled launch: duration, DRAM bytes, L1/L2 sectors and hit rates, pipe utilisation, issue-slot
utilisation, top stall reasons, and the executed-SASS opcode mix (from the source page)."""
import collections
import csv
import io
import subprocess
import sys

RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]


def run(args):
    return subprocess.run(["ncu", *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    lines = [f"# ncu summary of {rep} ({len(body)} profiled launches; `ncu --set full --clock-control none`)", ""]
    for k, r in enumerate(body):
        lines.append(f"## launch {k}: {r[col['Kernel Name']][:110]}")
        for m in RAW:
            if m in col and r[col[m]] not in ("", "n/a"):
                lines.append(f"  {m:78s} {r[col[m]]:>18s} {units[col[m]]}")
        stalls = sorted(((float(r[i]), h) for h, i in col.items()
                         if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
                         and r[i] not in ("", "n/a")), reverse=True)[:6]
        lines.append("  top stalls (warps stalled per issue-active cycle): " +
                     ", ".join(f"{h.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, h in stalls))
        lines.append("")
    src = list(csv.reader(io.StringIO(run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"]))))
    seen, cur, name = set(), None, None
    for r in src + [["Kernel Name", ""]]:
        if r and r[0] == "Kernel Name":
            if cur and name not in seen:
                seen.add(name)
                tot = sum(cur.values()) or 1.0
                lines.append(f"## executed SASS opcode mix: {name[:100]}  (total {tot:.0f} warp instructions)")
                lines.append("  " + ", ".join(f"{op} {c / tot * 100:.1f}%" for op, c in cur.most_common(18)))
                lines.append("")
            cur, name, h = collections.Counter(), (r[1] if len(r) > 1 else ""), None
            continue
        if r and r[0] == "Address":
            iS, iE = r.index("Source"), r.index("Instructions Executed")
            continue
        if cur is not None and r and len(r) > max(iS, iE):
            t = r[iS].split()
            if not t:
                continue
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            try:
                cur[op.split(".")[0]] += float(r[iE] or 0)
            except ValueError:
                pass
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
