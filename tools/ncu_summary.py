"""Extract the metrics DESIGN.md / profiles/ quote from an .ncu-rep (run where ncu is installed)."""
import csv, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
pats = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput",
        "dram__cycles_active", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "launch__shared_mem_per_block",
        "smsp__issue_active.avg.pct", "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "issue_stalled", "sm__throughput.avg.pct", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg"]
with open(out, "w") as f:
    for r in rows[2:]:
        f.write("== %s  grid %s block %s\n" % (r[hdr.index("Kernel Name")], r[hdr.index("Grid Size")] if "Grid Size" in hdr else "", r[hdr.index("Block Size")] if "Block Size" in hdr else ""))
        for i, k in enumerate(hdr):
            if any(p in k for p in pats) and "pcsamp" not in k:
                try:
                    if "issue_stalled" in k and float(r[i].replace(",", "")) < 0.05:
                        continue
                except ValueError:
                    pass
                f.write("  %-92s %s %s\n" % (k, r[i], units[i]))
print("wrote", out)
