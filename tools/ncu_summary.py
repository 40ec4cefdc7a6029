"""Dev tool: one line of key metrics per kernel launch of an .ncu-rep (ncu --set full).  usage: ncu_summary.py <report> [...]"""
import csv, io, subprocess, sys
KEYS = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "dramR"), ("dram__bytes_write.sum", "dramW"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
        ("launch__shared_mem_per_block_dynamic", "smem_dyn"), ("smsp__inst_executed.sum", "inst")]
for rep in sys.argv[1:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print("##", rep)
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        name = name.split("::")[-1][:46]
        out = []
        for k, short in KEYS:
            if k in hdr:
                v = r[hdr.index(k)]
                try:
                    v = "%.4g" % float(v)
                except ValueError:
                    pass
                out.append("%s=%s%s" % (short, v, units[hdr.index(k)] if short in ("us", "dramR", "dramW", "smem_dyn") else ""))
        print(name, " ".join(out))
