"""Dev tool: per-source-line share of stall samples / executed instructions from an .ncu-rep (needs -lineinfo + --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'hmma_cycles_active.avg.pct', 'data_bank_conflicts_pipe_lsu_mem_shared.sum', 'registers_per_thread', 'issue_stalled', 'lts__t_bytes.sum', 'sm__throughput.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ',
        'sm__inst_executed_pipe_', 'lts__t_sectors_srcunit_tex_op_read.sum', 'dram__throughput.avg.pct']
for i, k in enumerate(rows[0]):
    if any(s in k for s in keys) and 'Not Issued' not in k and 'not_issued' not in k:
        print(k, rows[1][i], rows[-1][i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
cur = None; agg = {}
for r in csv.reader(io.StringIO(src)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if len(r) < 10 or r[0] == "Line No" or r[2] != '-': continue
    try: ln = int(r[0])
    except ValueError: continue
    agg[(cur, ln)] = (int(r[6] or 0), int(r[7] or 0), r[1].strip()[:100])
ts = sum(v[0] for v in agg.values()) or 1; ti = sum(v[1] for v in agg.values()) or 1
print("total samples", ts, "inst", ti)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{k[0]}:{k[1]:4d} samp {100*v[0]/ts:5.1f}% inst {100*v[1]/ti:5.1f}%  {v[2]}")
