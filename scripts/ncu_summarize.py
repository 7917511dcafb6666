"""Per-launch summary of an `ncu --set full` report: python scripts/ncu_summarize.py file.ncu-rep > profiles/...txt"""
import csv, io, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__cycles_elapsed.avg.per_second"]
idx = {h: i for i, h in enumerate(hdr)}
for n, r in enumerate(data):
    print("== launch %d" % n)
    for w in want:
        if w in idx:
            print("   %-68s %s %s" % (w, r[idx[w]], units[idx[w]]))
