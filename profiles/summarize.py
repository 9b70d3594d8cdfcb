"""Turn the ncu captures brought back in gpurun_out/ into the small text summaries committed under profiles/.
usage: python profiles/summarize.py <report.ncu-rep> <out.txt> ; python profiles/summarize.py --launches <launches.csv> <out.txt>"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
        "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle", "smsp__pcsamp_warps_issue_stalled_barrier",
        "smsp__pcsamp_warps_issue_stalled_branch_resolving", "smsp__pcsamp_warps_issue_stalled_not_selected", "smsp__pcsamp_warps_issue_stalled_selected",
        "smsp__pcsamp_warps_issue_stalled_mio_throttle", "smsp__pcsamp_warps_issue_stalled_lg_throttle"]


def report(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    with open(out, "w") as f:
        for v in rows[2:]:
            d = dict(zip(h, v))
            f.write(f"kernel: {d.get('Kernel Name')}   grid {d.get('Grid Size')} block {d.get('Block Size')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"  {k:70s} {d[k]:>18s} {u[h.index(k)]}\n")
            f.write("\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
        srows = list(csv.reader(src.splitlines()))[2:]
        tot = sum(int(r[5]) for r in srows if len(r) > 8 and r[5].isdigit())
        f.write(f"SASS instructions executed (warp-level): {tot}\n")
        mix = collections.Counter()
        for r in srows:
            if len(r) > 8 and r[5].isdigit():
                op = r[1].strip().split()
                op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
                mix[op.split(".")[0]] += int(r[5])
        f.write("top opcodes by executed count: " + ", ".join(f"{k} {v / tot:.3f}" for k, v in mix.most_common(16)) + "\n")


def launches(csvp, out):
    rows = [r for r in csv.reader(open(csvp)) if len(r) > 5]
    hdr = None; agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if "Kernel Name" in r: hdr = r; continue
        if hdr is None: continue
        d = dict(zip(hdr, r))
        try: v = float(d["Metric Value"].replace(",", ""))
        except Exception: continue
        agg[d["Kernel Name"].split("(")[0]][0] += 1; agg[d["Kernel Name"].split("(")[0]][1] += v
    tot = sum(v[1] for k, v in agg.items() if "synth" not in k)
    with open(out, "w") as f:
        f.write("per-kernel device time from `ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k[:64]:64s} launches={v[0]:4d} total_ms={v[1] / 1e6:9.3f} avg_us={v[1] / v[0] / 1e3:9.1f} share_of_step={v[1] / tot:.3f}\n")


if __name__ == "__main__":
    if sys.argv[1] == "--launches": launches(sys.argv[2], sys.argv[3])
    else: report(sys.argv[1], sys.argv[2])
