"""Prints the headline metrics of an .ncu-rep (first profiled launch) — used for profiles/*.txt."""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_static", "launch__waves_per_multiprocessor", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "smsp__sass_inst_executed_op_global_st.sum",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
]
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w:78s} {units[i]:16s} {r[i]}")
stalls = []
for i, h in enumerate(hdr):
    if "warps_issue_stalled" in h and "per_issue_active" in h:
        try:
            stalls.append((float(r[i]), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
        except ValueError:
            pass
print("stall reasons (warps per issue-active cycle): " + ", ".join(f"{h}={v:.2f}" for v, h in sorted(stalls, reverse=True)[:8]))
try:
    rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale[units[hdr.index("dram__bytes_read.sum")]]; wr *= scale[units[hdr.index("dram__bytes_write.sum")]]
    print(f"traffic = dram read + write per launch = {rd + wr:.0f} bytes")
except Exception as e:  # noqa: BLE001
    print("traffic: n/a", e)
