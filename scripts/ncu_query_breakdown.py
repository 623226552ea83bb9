"""Per-kernel time of the LAST engine-2 vsm_query call in an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys
from collections import OrderedDict
rows = list(csv.reader(open(sys.argv[1])))
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        h, start = r, i + 1
        break
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
names = ("query_tc_kernel", "rescore", "row_norm", "pad_prompts", "prompt_prep", "query_exact", "query_merge", "query_select")
seq = []
for r in rows[start:]:
    if len(r) > mv and any(k in r[kn] for k in names):
        try:
            seq.append((r[kn].split("(")[0][-44:], float(r[mv].replace(",", "")) / 1e3))
        except ValueError:
            pass
tcs = [i for i, (n, _) in enumerate(seq) if "query_tc_kernel" in n]
last = tcs[-1]
# a call ends with query_select after its last tc kernel; it starts after the previous call's final select
end = next(i for i in range(last, len(seq)) if "query_select" in seq[i][0])
# walk back over this call's levels (at most two earlier selects belong to it)
begin = last
n_tc = 0
while begin > 0:
    n = seq[begin - 1][0]
    if "query_tc_kernel" in n:
        n_tc += 1
    if "query_select" in n and n_tc == 0 and begin - 1 < tcs[0]:
        break
    if "rescore" in n and sum("query_tc" in x[0] for x in seq[begin - 1:last]) >= 2:
        break
    begin -= 1
agg, tot = OrderedDict(), 0.0
for n, t in seq[begin:end + 1]:
    agg.setdefault(n, [0, 0.0])
    agg[n][0] += 1
    agg[n][1] += t
    tot += t
for n, (c, t) in agg.items():
    print(f"{n:46s} n={c:3d} {t:10.1f} us")
print(f"total {tot:.1f} us")
