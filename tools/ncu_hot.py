"""Reads an .ncu-rep (here, no GPU needed): headline metrics + the hottest SASS lines with their top stall reasons."""
import csv, subprocess, sys, io
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    for k in ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
              "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
              "sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active",
              "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_active.avg",
              "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "lts__t_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
              "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]:
        if k in d:
            print(f"{k:80s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) >= len(hdr)]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
agg = {}
for n, r in enumerate(data):
    s = int(r[ix["# Samples"]] or 0)
    for h in stalls:
        agg[h] = agg.get(h, 0) + int(r[ix[h]] or 0)
    if s >= tot * thr:
        top = sorted(((int(r[ix[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
        print(n, r[ix["Source"]][:72].ljust(72), s, r[ix["Instructions Executed"]], top)
print(sorted(((v, k) for k, v in agg.items()), reverse=True)[:8])
# aggregate by opcode within a line range (argv[3], argv[4])
if len(sys.argv) > 4:
    a, b = int(sys.argv[3]), int(sys.argv[4])
    by = {}
    for n in range(a, b):
        r = data[n]
        op = r[ix["Source"]].split()[0]
        if op.startswith("@"):
            op = r[ix["Source"]].split()[1]
        op = op.split(".")[0]
        e = by.setdefault(op, [0, 0, {}])
        e[0] += int(r[ix["# Samples"]] or 0)
        e[1] += 1
        for h in stalls:
            e[2][h[6:]] = e[2].get(h[6:], 0) + int(r[ix[h]] or 0)
    tot2 = sum(e[0] for e in by.values())
    print("range", a, b, "samples", tot2)
    for op, e in sorted(by.items(), key=lambda kv: -kv[1][0]):
        top = sorted(((v, k) for k, v in e[2].items()), reverse=True)[:3]
        print(f"{op:12s} n={e[1]:4d} samples={e[0]:7d} ({100*e[0]/tot2:.1f}%)", top)
