"""Condense an `ncu --page raw --csv` dump with many kernels into one row per profiled launch:
usage: python tools/ncu_summary.py raw.csv out.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[0]
col = {k: i for i, k in enumerate(h)}
want = [("gpu__time_duration.sum", "time_us"), ("dram__bytes_read.sum", "dram_read"), ("dram__bytes_write.sum", "dram_write"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor_pct"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsu_wavefront_pct"),
        ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "tc_smem_read_pct"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_pct"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy_pct"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block")]
def find(key):
    for k, i in col.items():
        if k == key or k.endswith("." + key):
            return i
    return None
idx = [(find(k), n) for k, n in want]
kn = col["Kernel Name"]
units = rows[1]
body = [r for r in rows[2:] if len(r) > kn]
with open(sys.argv[2], "w") as f:
    f.write("launch,kernel," + ",".join(n + ("[%s]" % units[i] if i is not None and units[i] else "") for i, n in idx) + "\n")
    for j, r in enumerate(body):
        f.write('%d,"%s",' % (j, r[kn][:100].replace('"', "'")) + ",".join((r[i].replace(",", "") if i is not None else "") for i, n in idx) + "\n")
print("launches:", len(body))
