"""SASS opcode histogram of one kernel of the built library (static instruction counts):
usage: sass_hist.py <libsmem_gpu.so> <substring of the mangled kernel name> [top N]"""
import collections, re, subprocess, sys
so, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cur, on, ops = None, False, collections.Counter()
name = None
for l in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", l)
    if m:
        on = pat in m.group(1) and name in (None, m.group(1))
        if on:
            name = m.group(1)
        continue
    if not on:
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
    if m:
        ops[m.group(1)] += 1
tot = sum(ops.values())
base = collections.Counter()
for k, v in ops.items():
    base[k.split(".")[0]] += v
print(f"kernel {name}: {tot} SASS instructions")
print("-- by mnemonic")
for k, v in base.most_common(top):
    print(f"{v:6d}  {100 * v / tot:5.1f}%  {k}")
print("-- memory / shuffle / popcount forms")
for k, v in sorted(ops.items(), key=lambda kv: -kv[1]):
    if re.match(r"(LDG|STG|LDS|STS|LDC|LDCU|ATOM|RED|SHFL|POPC|UBLKCP|SYNCS|LDL|STL)", k):
        print(f"{v:6d}  {k}")
