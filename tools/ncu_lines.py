"""Join an `ncu --page source --csv` dump with `nvdisasm -g` line info: instructions executed / stall samples per CUDA source line.
usage: ncu_lines.py <source.csv> <libsmem_gpu.so> <mangled kernel name> [top N]"""
import collections, csv, re, subprocess, sys, tempfile, os
src_csv, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
lines, cur, on = [], None, False          # source line of every instruction of the kernel, in order
for l in sass:
    if l.startswith(".text."):
        on = l.strip() == f".text.{kern}:"
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
recs = []
for r in rows[2:]:
    try:
        recs.append((int(r[ix["Instructions Executed"]]), int(r[ix["Thread Instructions Executed"]]), int(r[ix["# Samples"]])))
    except (ValueError, IndexError):
        continue
print(f"{len(recs)} profiled instructions, {len(lines)} disassembled", file=sys.stderr)
n = min(len(recs), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0])
for (inst, thr, samp), ln in zip(recs[:n], lines[:n]):
    a = agg[ln]; a[0] += inst; a[1] += thr; a[2] += samp
ti = sum(a[0] for a in agg.values()); ts = sum(a[2] for a in agg.values())
print(f"total warp-inst {ti}  samples {ts}")
print("warp-inst%  samples%  avg-lanes  line")
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * a[0] / ti:8.2f}  {100 * a[2] / ts:8.2f}  {a[1] / max(a[0], 1):8.1f}   {ln}")
