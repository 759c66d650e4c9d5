"""Condenses an `ncu --page raw --csv` dump into one block per kernel launch with the counters the roofline argument needs."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_barrier.ratio",
        "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio"]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    print(r[ik][:150])
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"    {w:90s} {r[i]} {units[i]}")
    if "dram__bytes_read.sum" in hdr and "gpu__time_duration.sum" in hdr:
        def val(name):
            i = hdr.index(name); v = float(r[i]); u = units[i]
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        t = float(r[hdr.index("gpu__time_duration.sum")]) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(units[hdr.index("gpu__time_duration.sum")], 1e-6)
        print(f"    -> measured DRAM traffic {(val('dram__bytes_read.sum') + val('dram__bytes_write.sum')) / t / 1e9:.0f} GB/s")
