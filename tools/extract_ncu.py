"""Turn ncu reports / launch lists brought back in gpurun_out/ into the small CSV summaries kept under profiles/.

  python tools/extract_ncu.py raw  <report.ncu-rep> [<report2.ncu-rep> ...] > profiles/<name>.csv
      one row per profiled launch, the columns a reader needs (times, instruction counts, pipes, caches, DRAM bytes,
      occupancy limits); every other metric stays in the .ncu-rep (not committed: tens of MB)
  python tools/extract_ncu.py stalls <report.ncu-rep>
      instruction mix and warp-stall reasons of the first launch, from the source page
  python tools/extract_ncu.py shares <launches.csv>
      per-kernel share of the summed launch durations of an `ncu --metrics gpu__time_duration.sum --csv` list
"""
import csv
import subprocess
import sys
from collections import Counter

KEEP = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "derived__smsp__sass_thread_inst_executed_op_dfma_pred_on_x2", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed", "thread_inst_executed", "thread_inst_executed_true",
        "sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "idc__request_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def page(rep, which, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", which, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(out.splitlines()))


def raw(reps):
    w = csv.writer(sys.stdout)
    first = True
    for rep in reps:
        rows = page(rep, "raw")
        hdr, units = rows[0], rows[1]
        cols = [hdr.index(k) for k in KEEP if k in hdr]
        if first:
            w.writerow(["report"] + [hdr[i] for i in cols])
            w.writerow(["unit"] + [units[i] for i in cols])
            first = False
        for r in rows[2:]:
            w.writerow([rep.split("/")[-1]] + [r[i] for i in cols])


def stalls(rep):
    rows = page(rep, "source", ["--print-source", "sass"])
    hdr = rows[1]
    data = []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name":
            break
        if len(r) == len(hdr) and r[0].startswith("0x"):
            data.append(r)
    ci, cs = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    ti, ts = sum(int(r[ci]) for r in data), sum(int(r[cs]) for r in data)
    print(f"kernel: {rows[0][1][:80]}\nwarp instructions {ti}, stall samples {ts}, SASS lines {len(data)}")
    mix, st = Counter(), Counter()
    for r in data:
        op = r[1].split()
        o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
        mix[o] += int(r[ci]); st[o] += int(r[cs])
    print("opcode, % of warp instructions, % of stall samples")
    for o, c in mix.most_common(16):
        print(f"{o},{100 * c / ti:.2f},{100 * st[o] / ts:.2f}")
    print("stall reason, % of samples")
    for j, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            s = sum(int(r[j] or 0) for r in data)
            if s > 0.005 * ts:
                print(f"{h},{100 * s / ts:.1f}")


def shares(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    tot, cnt = Counter(), Counter()
    for r in rows:
        name = r[4].split("(")[0]
        tot[name] += float(r[14]); cnt[name] += 1
    s = sum(tot.values())
    print("kernel, launches, total ms, share %")
    for k, v in tot.most_common():
        print(f"{k},{cnt[k]},{v / 1e6:.3f},{100 * v / s:.2f}")


if __name__ == "__main__":
    {"raw": lambda: raw(sys.argv[2:]), "stalls": lambda: stalls(sys.argv[2]), "shares": lambda: shares(sys.argv[2])}[sys.argv[1]]()
