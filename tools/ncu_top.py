"""Summarise an .ncu-rep: key raw metrics per kernel and the top stall instructions of one kernel."""
import csv
import subprocess
import sys

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "lts__t_bytes.sum"]
for r in rows[2:]:
    print("----")
    for w in want:
        if w in hdr:
            print("  %-70s %s %s" % (w, r[hdr.index(w)], rows[1][hdr.index(w)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
blocks, cur, h = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []
        blocks.append((r[1], cur))
        continue
    if r and r[0] == "Address":
        h = r
        continue
    if cur is not None and h and len(r) == len(h):
        cur.append(r)
name, out = blocks[which]
print("==== source hot spots:", name[:80])
si, sc, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
stall = [i for i, x in enumerate(h) if x.startswith("stall_") and "Not Issued" not in x]
tot = sum(int(r[si]) for r in out)
agg = {}
for r in out:
    for i in stall:
        agg[h[i][6:]] = agg.get(h[i][6:], 0) + int(r[i])
print("total samples", tot, "instructions", len(out))
print("stall totals:", sorted(agg.items(), key=lambda x: -x[1])[:10])
for r in sorted(out, key=lambda r: -int(r[si]))[:topn]:
    st = sorted(((h[i][6:], int(r[i])) for i in stall if int(r[i]) > 0), key=lambda x: -x[1])[:3]
    print("%6d %5.1f%% exec=%8s  %-64s %s" % (int(r[si]), 100 * int(r[si]) / tot, r[ie], r[sc].strip()[:64], st))
