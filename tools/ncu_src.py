"""Developer tool: per-phase (barrier-delimited) and per-opcode summary of an ncu --set full report's SASS page.
usage: python tools/ncu_src.py report.ncu-rep [metric names to print from the raw page ...]"""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active"] + sys.argv[2:]
for h, u, v in zip(hdr, units, vals):
    if h in want or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.3):
        print(f"{h:100s} {u:10s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
num = lambda r, k: int(r[ix[k]] or 0)
ts, ti = sum(num(r, "# Samples") for r in data), sum(num(r, "Instructions Executed") for r in data)
print(f"samples {ts}  warp instructions {ti}")
seg = acc_s = acc_i = start = 0
for k, r in enumerate(data):
    acc_s += num(r, "# Samples"); acc_i += num(r, "Instructions Executed")
    if "BAR.SYNC" in r[ix["Source"]] or k == len(data) - 1:
        print(f"  phase {seg}: sass rows {start}-{k}: samples {100 * acc_s / max(ts, 1):5.1f}%  instructions {100 * acc_i / max(ti, 1):5.1f}%")
        seg += 1; acc_s = acc_i = 0; start = k + 1
byop, byop_s = collections.Counter(), collections.Counter()
for r in data:
    s = r[ix["Source"]].split()
    op = s[1] if s[0].startswith("@") else s[0]
    byop[op] += num(r, "Instructions Executed"); byop_s[op] += num(r, "# Samples")
for op, c in byop.most_common(14):
    print(f"  {op:24s} instructions {100 * c / ti:5.1f}%  samples {100 * byop_s[op] / max(ts, 1):5.1f}%")
top = sorted(data, key=lambda r: -num(r, "# Samples"))[:12]
print("hottest instructions:")
for r in top:
    print(f"  {100 * num(r, '# Samples') / max(ts, 1):5.1f}%  {r[ix['Address']][-5:]}  {r[ix['Source']][:90]}")
