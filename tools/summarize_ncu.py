"""Write the text summary of an ncu report kept under profiles/: per-kernel raw metrics + per-source-line shares.
usage: python tools/summarize_ncu.py report.ncu-rep "command line that was profiled" kernel-regex [kernel-regex ...]"""
import csv, io, subprocess, sys
rep, cmd, kerns = sys.argv[1], sys.argv[2], sys.argv[3:]
KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# ncu --set full --clock-control none --import-source on; {cmd}; report {rep}")
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
for r in rows[2:]:
    d = dict(zip(h, r))
    print("##", d["Kernel Name"][:90])
    for k in KEYS:
        if k in d:
            print(f"   {k} {d[k]} {rows[1][h.index(k)]}")
for k in kerns:
    print(f"\n## per source line: {k}")
    out = subprocess.run([sys.executable, "tools/ncu_lines.py", rep, k, "22"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    print(out)
