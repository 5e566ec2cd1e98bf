"""Compact text summary of an .ncu-rep (ncu --set full): per captured launch the duration, DRAM bytes / utilisation, tensor-pipe
and LSU utilisation, occupancy, registers, and the top warp-stall reasons.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep ["header text"] > profiles/x.txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
want = [("gpu__time_duration.sum", "duration"), ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
        ("lts__t_bytes.sum", "L2 bytes"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
        ("sm__inst_executed.avg.per_cycle_active", "IPC (active cycles)"), ("launch__registers_per_thread", "registers / thread"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__cycles_active.avg", "SM sub-partition active cycles")]
if len(sys.argv) > 2:
    print("# " + sys.argv[2])
print(f"# source: ncu -i {rep.split('/')[-1]} --page raw --csv  (ncu --set full --clock-control none; ~40 replays per launch, cold caches)")
for r in rows[2:]:
    print(f"\nkernel: {r[col['Kernel Name']][:150]}")
    for key, label in want:
        if key in col:
            print(f"  {label:40s} {r[col[key]]:>16s} {units[col[key]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio") or (h.startswith("smsp__average_warp_latency_issue_stalled_") and h.endswith(".ratio")):
            try:
                stalls.append((float(r[i].replace(",", "")), h.split("stalled_")[1].split("_per_issue")[0].replace(".ratio", "")))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    if stalls:
        print("  top warp-stall reasons (cycles per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))
