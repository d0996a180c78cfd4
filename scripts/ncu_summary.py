"""Summarise an .ncu-rep (read here, no GPU needed) into a small markdown file for profiles/.

usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep profiles/x.md [title]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "smsp__pipe_tensor_subpipe_dmma_cycles_active.avg", "sm__cycles_active.avg",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
    "dram__bytes_write.sum.per_second", "lts__t_sector_hit_rate.pct", "lts__t_sectors.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed_op_global_red.sum", "smsp__issue_active.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def ncu(rep, page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else rep
    raw = ncu(rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = [f"# {title}", "", f"source: `{rep}` (ncu --set full --clock-control none --import-source on); "
             "numbers under the profiler are cold-cache and serialised, never bench values.", ""]
    for k, row in enumerate(raw[2:]):
        d = dict(zip(hdr, row))
        lines += [f"## launch {k}: `{d.get('Kernel Name', '?')}`", "", "| metric | value | unit |", "|---|---|---|"]
        for key in KEYS:
            for h, u, v in zip(hdr, units, row):
                if h == key or h.endswith("." + key):
                    lines.append(f"| {key} | {v} | {u} |")
                    break
        lines.append("")
    src = ncu(rep, "source")
    # first kernel only: top instructions by stall samples
    try:
        h = src[1]
        idx = {x: i for i, x in enumerate(h)}
        body = []
        for r in src[2:]:
            if len(r) < len(h) or r[0] == "Kernel Name":
                break
            body.append(r)
        tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
        top = sorted(body, key=lambda r: -int(r[idx["# Samples"]] or 0))[:15]
        lines += ["## hottest SASS instructions of launch 0 (warp-stall samples)", "",
                  f"total samples {tot}", "", "| samples | share | instruction | dominant stall |", "|---|---|---|---|"]
        stall_cols = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
        for r in top:
            n = int(r[idx["# Samples"]] or 0)
            dom = max(stall_cols, key=lambda c: int(r[idx[c]] or 0))
            lines.append(f"| {n} | {n / max(tot, 1):.3f} | `{r[idx['Source']].strip()[:80]}` | {dom} |")
        mn = {}
        for r in body:
            op = r[idx["Source"]].strip().split()[0] if r[idx["Source"]].strip() else ""
            if op.startswith("@"):
                op = r[idx["Source"]].strip().split()[1]
            op = op.split(".")[0]
            mn[op] = mn.get(op, 0) + int(r[idx["Instructions Executed"]] or 0)
        keep = {k: v for k, v in mn.items() if k in ("DMMA", "LDGSTS", "REDG", "RED", "LDS", "LDG", "STG", "UBLKCP",
                                                      "UTMALDG", "BAR", "ATOMG")}
        lines += ["", "## executed warp instructions by SASS mnemonic (launch 0)", "",
                  ", ".join(f"{k}: {v}" for k, v in sorted(keep.items(), key=lambda kv: -kv[1]))]
    except Exception as e:  # noqa: BLE001
        lines.append(f"(source page not parsed: {e})")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst)


if __name__ == "__main__":
    main()
