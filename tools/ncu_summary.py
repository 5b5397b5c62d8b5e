"""Text summary of one `ncu --set full --import-source on` capture: key raw metrics + the SASS lines with most stall samples.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep "title / command" > profiles/rN_xxx.txt
"""
import csv
import io
import subprocess
import sys

rep, title = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active",
        "sm__inst_executed_pipe_tensor", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "sm__inst_executed_pipe_xu", "sm__inst_executed_pipe_fma", "smsp__inst_executed_pipe_fma", "smsp__inst_executed_pipe_xu",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {title}\n# source: {rep} (ncu --set full --clock-control none --import-source on)\n")
for vals in rows[2:]:
    d = dict(zip(hdr, vals))
    print("kernel:", d.get("Kernel Name"))
    for h, u, v in zip(hdr, units, vals):
        if any(k in h for k in KEYS) and v not in ("", "0", "n/a"):
            print(f"  {h:90s} {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if len(rows) > 2:
    h = rows[1]
    ix = {n: i for i, n in enumerate(h)}
    data = [r for r in rows[2:] if len(r) == len(h)]
    stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    agg = {n: sum(int(r[ix[n]] or 0) for r in data) for n in stalls}
    print(f"\nwarp-state samples: {tot}; by reason: " + ", ".join(f"{k[6:]} {100 * v / max(tot, 1):.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    print("top SASS lines by samples (samples, executed, instruction, top stall reasons):")
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:14]:
        st = {n[6:]: int(r[ix[n]] or 0) for n in stalls}
        st = ", ".join(f"{k} {v}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
        print(f"  {int(r[ix['# Samples']] or 0):6d} {r[ix['Instructions Executed']]:>9s}  {r[ix['Source']][:64]:64s} {st}")
