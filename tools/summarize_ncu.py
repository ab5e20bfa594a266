"""Summarise ncu outputs into small tracked files under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md "<command>"
  python tools/summarize_ncu.py raw gpurun_out/prof.ncu-rep profiles/r01_fa_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "sm__inst_executed_pipe_xu.sum", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__cycles_elapsed.max",
        "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
        "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_wait.ratio",
        "smsp__average_warp_latency_issue_stalled_sleeping.ratio",
        "smsp__average_warp_latency_issue_stalled_membar.ratio",
        "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_no_instruction.ratio",
        "smsp__average_warp_latency_issue_stalled_branch_resolving.ratio",
        "smsp__average_warp_latency_issue_stalled_dispatch_stall.ratio",
        "smsp__average_warp_latency_issue_stalled_tex_throttle.ratio",
        "smsp__average_warp_latency_issue_stalled_not_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_selected.ratio",
        "smsp__average_warp_latency_issue_stalled_imc_miss.ratio",
        "smsp__average_warp_latency_issue_stalled_drain.ratio"]


def launches(path, out, cmd):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(list)
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        agg[name].append(v)
    tot = sum(sum(v) for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised)\n\n")
        f.write(f"command: `{cmd}`\n\ntotal kernel time {tot / 1e3:.2f} ms over {sum(len(v) for v in agg.values())} launches\n\n")
        f.write("| share | launches | avg us | min us | max us | kernel |\n|---|---|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write(f"| {sum(v) / tot * 100:.1f}% | {len(v)} | {sum(v) / len(v):.1f} | {min(v):.1f} | {max(v):.1f} | `{k}` |\n")
    print(open(out).read()[:3000])


def raw(path, out):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary of {path}\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:100]}  (id {r[idx['ID']]})\n\n")
            for k in KEYS:
                if k in idx and r[idx[k]] not in ("", "n/a"):
                    f.write(f"- {k}: {r[idx[k]]} {units[idx[k]]}\n")
            f.write("\n")
    print(open(out).read()[:6000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
    else:
        raw(sys.argv[2], sys.argv[3])
