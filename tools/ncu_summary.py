"""Summarise an .ncu-rep (raw + source pages) into a short text report: python tools/ncu_summary.py rep [kernel-regex]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum",
        "sm__cycles_elapsed.max", "local_load/store: smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(f"{w} [{units[i]}]: {[r[i] for r in data]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr, data = rows[h], rows[h + 1:]
ix = {k: i for i, k in enumerate(hdr)}
stalls = collections.Counter()
ops, samp = collections.Counter(), collections.Counter()
for r in data:
    if len(r) < len(hdr):
        continue
    for k in hdr:
        if k.startswith("stall_") and "Not Issued" not in k:
            try:
                stalls[k] += int(r[ix[k]])
            except ValueError:
                pass
    toks = r[ix["Source"]].split()
    op = toks[1] if toks[0].startswith("@") else toks[0]
    op = op.split(".")[0]
    try:
        ops[op] += int(r[ix["Instructions Executed"]]); samp[op] += int(r[ix["# Samples"]])
    except ValueError:
        pass
print("SASS lines:", len(data))
ts = sum(stalls.values())
print("stalls:", ", ".join(f"{k[6:]} {v / ts * 100:.1f}%" for k, v in stalls.most_common(9)))
ti, tsm = sum(ops.values()), sum(samp.values())
print("warp instructions:", ti)
for op, c in ops.most_common(22):
    print(f"  {op:10s} {c / ti * 100:6.2f}% inst  {samp[op] / max(tsm, 1) * 100:6.2f}% samples")
