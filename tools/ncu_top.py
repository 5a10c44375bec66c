"""Summarise an .ncu-rep: key raw metrics + top stall lines of the SASS page.  usage: ncu_top.py rep [N]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "launch__registers_per_thread",
        "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
for i, n in enumerate(h):
    if n in want:
        print(f"{n:75s} {[r[i] for r in rows[2:]]} {rows[1][i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hidx = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
h = rows[hidx[0]]; body = rows[hidx[0] + 1:(hidx[1] - 1 if len(hidx) > 1 else None)]
ci = {n: i for i, n in enumerate(h)}
S = "# Samples"
body = [r for r in body if len(r) > ci[S]]
tot = sum(int(r[ci[S]] or 0) for r in body)
stalls = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
print("total samples", tot)
for r in sorted(body, key=lambda r: -int(r[ci[S]] or 0))[:topn]:
    s = int(r[ci[S]])
    st = sorted([(int(r[ci[n]] or 0), n) for n in stalls], reverse=True)[:2]
    print(f"{s:7d} {100 * s / tot:5.1f}%  {r[ci['Source']][:64]:64s} {st}")
