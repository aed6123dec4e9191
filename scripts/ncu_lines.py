"""Summarise one launch of an ncu report: headline metrics, warp-stall mix, opcode mix and the
source lines with the most samples.  usage: ncu_lines.py report.ncu-rep [launch index] [top n]"""
import collections, csv, io, subprocess, sys
def I(x):
    try: return int(x)
    except ValueError: return 0
rep = sys.argv[1]; idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0; top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
def ncu(*a):
    return subprocess.run(["ncu", "-i", rep, "--launch-skip", str(idx), "--launch-count", "1", *a], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
h, v = raw[0], raw[2]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__sass_inst_executed_op_local_ld.sum", "smsp__sass_inst_executed_op_local_st.sum"]
for w in want:
    if w in h:
        print(f"{w:75s} {v[h.index(w)][:110]} {raw[1][h.index(w)]}")
rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]; ix = {}
for i, k in enumerate(hdr): ix.setdefault(k, i)
src_col = 1; sass_col = 3
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
lines = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] not in ("-", "")]
sass = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0] in ("-", "")]
n = sum(I(r[ix["# Samples"]]) for r in sass)
tot = {k: sum(I(r[ix[k]]) for r in sass) for k in stalls}
print("samples", n, " | ", "  ".join(f"{k[6:]} {100 * x / n:.1f}" for k, x in sorted(tot.items(), key=lambda t: -t[1])[:9]))
opc = collections.Counter(); ops = collections.Counter()
for r in sass:
    t = [o for o in r[sass_col].split() if not o.startswith("@")]
    if not t: continue
    op = t[0].split(".")[0]
    opc[op] += I(r[ix["Instructions Executed"]]); ops[op] += I(r[ix["# Samples"]])
te = sum(opc.values())
print("sass instructions (static)", len(sass), " executed", te)
print("opcodes:", "  ".join(f"{o} {100 * c / te:.1f}%({100 * ops[o] / n:.0f}%s)" for o, c in opc.most_common(14)))
lines.sort(key=lambda r: -I(r[ix["# Samples"]]))
for r in lines[:top]:
    g = lambda k: r[ix[k]].rjust(5)
    print(r[0].rjust(4), g("# Samples"), "long", g("stall_long_sb"), "noinst", g("stall_no_inst"), "wait", g("stall_wait"), "short", g("stall_short_sb"),
          "bar", g("stall_barrier"), "mio", g("stall_mio"), "math", g("stall_math"), "|", r[src_col].strip()[:70])
