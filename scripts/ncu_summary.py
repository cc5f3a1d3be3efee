"""Summarise an .ncu-rep (read here, no GPU): key metrics, stall mix, opcode mix, hottest source lines."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "sm__cycles_elapsed.avg.per_second",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, h in enumerate(hdr):
    if h in want:
        print(f"{h:70s} {units[i]:12s} {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def col(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot_i = sum(col(r, "Instructions Executed") for r in data)
tot_s = sum(col(r, "# Samples") for r in data)
print(f"\nwarp insts {tot_i:.3e}  samples {tot_s:.0f}")
st = {h: sum(col(r, h) for r in data) for h in hdr if h.startswith("stall_") and "Not Issued" not in h}
print("stalls:", ", ".join(f"{k[6:]} {100*v/tot_s:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
op, sm = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]])
    if m:
        o = m.group(2).split(".")[0]
        op[o] += col(r, "Instructions Executed"); sm[o] += col(r, "# Samples")
print("opcodes:", ", ".join(f"{o} {100*v/tot_i:.1f}%" for o, v in op.most_common(18)))
if len(sys.argv) > 2:
    print("\nhottest instructions by samples:")
    for r in sorted(data, key=lambda r: -col(r, "# Samples"))[:int(sys.argv[2])]:
        print(f"{col(r,'# Samples'):8.0f} {col(r,'Instructions Executed'):12.0f}  {r[ix['Source']][:90]}")
