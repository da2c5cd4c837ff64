"""Per-GEMM durations of the MLP evaluation from an ncu launch list: the batched tensor-core GEMM launches repeat in the
order H1, H2, gW2, dA1, gW1; prints the median of every position (us) and the other kernels' totals per evaluation."""
import collections
import csv
import statistics
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if r and r[0].isdigit()]
bg = [float(r[-1]) / 1e3 for r in rows if "k_tc_bgemm" in r[4]]
n_eval = len(bg) // 5
names = ["H1", "H2", "gW2", "dA1", "gW1"]
print("evaluations", n_eval, " ".join("%s %.1f" % (names[i], statistics.median(bg[i::5])) for i in range(5)),
      "| sum %.1f" % sum(statistics.median(bg[i::5]) for i in range(5)))
other = collections.OrderedDict()
for r in rows:
    if "bhmc::" in r[4] and "k_tc_bgemm" not in r[4]:
        other[r[4][:34]] = other.get(r[4][:34], 0.0) + float(r[-1]) / 1e3
print(" ".join("%s %.1f" % (k.replace("bhmc::", "").replace("void ", ""), v / max(n_eval, 1)) for k, v in other.items()))
