"""Key metrics of the first kernel in an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread",
        "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
for r in rows[2:]:
    print(f"# {rep}")
    for h, u, v in zip(hdr, units, r):
        if h in want:
            print(f"{h:75s} {v} {u}")
    stalls = []
    for h, v in zip(hdr, r):
        if "issue_stalled" in h and h.endswith("_per_issue_active.ratio"):
            try:
                stalls.append((float(v), h))
            except ValueError:
                pass
    print("stall reasons (average warps stalled per issue-active cycle):")
    for v, h in sorted(stalls, reverse=True)[:8]:
        print(f"   {v:6.3f}  {h}")
