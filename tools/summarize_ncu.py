"""Development aid: text summary of an .ncu-rep (selected raw metrics per captured kernel) for profiles/."""
import csv, subprocess, sys
WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("=" * 100)
        print(f"{rep}: {vals[hdr.index('Kernel Name')]}")
        for w in WANT:
            if w in hdr:
                print(f"{w:80s} {units[hdr.index(w)]:16s} {vals[hdr.index(w)]}")
        stalls = sorted(((float(vals[i]), h) for i, h in enumerate(hdr) if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and vals[i] not in ("", "n/a")), reverse=True)
        for v, h in stalls[:7]:
            print(f"{h:80s} {'inst':16s} {v:.3f}")
