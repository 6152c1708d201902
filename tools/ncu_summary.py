"""Summarise an .ncu-rep (read on the CPU box): key metrics per launch + top stall instructions."""
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
out_json = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
        "launch__registers_per_thread", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.max", "sm__cycles_elapsed.max.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum.per_cycle_elapsed", "gpc__cycles_elapsed.max.per_second",
        "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct"]
summary = []
for r in rows[2:]:
    d = {}
    for i, h in enumerate(hdr):
        if h in keep and i < len(r):
            d[h] = f"{r[i]} {units[i]}".strip()
    summary.append(d)
print(json.dumps(summary, indent=1))
if out_json:
    json.dump(summary, open(out_json, "w"), indent=1)
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                      text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
secs = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
start = secs[0]
end = secs[1] if len(secs) > 1 else len(rows)
h = rows[start + 1]
idx = {x: i for i, x in enumerate(h)}
body = rows[start + 2:end]


def f(r, k):
    try:
        return float(r[idx[k]])
    except Exception:
        return 0.0


tot = sum(f(r, "# Samples") for r in body)
print("instructions", len(body), "samples", tot)
stall_cols = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
agg = {c: sum(f(r, c) for r in body) for c in stall_cols}
print("stall totals:", {k: round(100 * v / tot, 1) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(body, key=lambda r: -f(r, "# Samples"))[:25]:
    st = sorted(((f(r, c), c) for c in stall_cols), reverse=True)[:2]
    print(f"{100 * f(r, '# Samples') / tot:5.1f}% exec={r[idx['Instructions Executed']]:>10} {r[idx['Source']][:64]:64s} {[(int(a), b) for a, b in st]}")
