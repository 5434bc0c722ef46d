"""Aggregate an `ncu --page source --csv` dump: where do instructions / stall samples go?"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
tot_inst = tot_thr = tot_samp = 0
recs = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try:
        inst = int(r[ix["Instructions Executed"]]); thr = int(r[ix["Thread Instructions Executed"]]); samp = int(r[ix["# Samples"]])
    except ValueError:
        continue
    recs.append((r[ix["Address"]], r[ix["Source"]], inst, thr, samp, int(r[ix["stall_long_sb"]] or 0)))
    tot_inst += inst; tot_thr += thr; tot_samp += samp
print("total warp-inst", tot_inst, "thread-inst", tot_thr, "avg lanes", tot_thr / max(tot_inst, 1), "samples", tot_samp)
# opcode histogram
ops = collections.Counter(); opthr = collections.Counter(); opsamp = collections.Counter()
for a, s, inst, thr, samp, lsb in recs:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += inst; opthr[op] += thr; opsamp[op] += samp
print("\nopcode        warp-inst%   avg-lanes  samples%")
for op, n in ops.most_common(25):
    print(f"{op:12s} {100 * n / tot_inst:8.2f}   {opthr[op] / max(n, 1):8.1f}  {100 * opsamp[op] / tot_samp:8.2f}")
print("\ntop stall-sample instructions")
for a, s, inst, thr, samp, lsb in sorted(recs, key=lambda t: -t[4])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    print(f"{100 * samp / tot_samp:6.2f}%  inst {100 * inst / tot_inst:5.2f}%  lanes {thr / max(inst, 1):5.1f}  {s.strip()[:110]}")
