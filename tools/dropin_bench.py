#!/usr/bin/env python
"""Throughput of the DROP-IN surface itself (missing #1): the reference's own `bwa mem` with only
bwt_smem1_batched swapped for the GPU adapter (oracle/_ref/bwa_gpu), driven the way the reference drives its accelerator --
`-t` worker threads calling the batch API with `-b` reads each (bwamem.c:390-393, kthread_batch.c:46-59, fastmap.c:54,62) --
next to the unmodified CPU path (`bwa_ref_timed mem -t T -b 1`, which never offloads).

Reports per run: wall time, reads/s, seeding time (wall time inside bwt_smem1_batched summed over the worker threads; for the
CPU path the same figure from oracle/timed_batched.c), GPU calls, requests per call, and whether the SAM equals the CPU path's.
The index is built by the GPU builder and written in the reference's file formats (.bwt / .sa; .pac / .ann / .amb by the
reference's own `bwa fa2pac`); it is checked bit-for-bit against `bwa index` at fixture sizes in tests/.

  python tools/dropin_bench.py --ref-bp 100000000 --reads 1000000 --out gpurun_out/r2_dropin.json
"""
import argparse
import hashlib
import importlib
import json
import os
import re
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref", "bwa_ref")
REF_TIMED = os.path.join(ROOT, "oracle", "_ref", "bwa_ref_timed")
GPU = os.path.join(ROOT, "oracle", "_ref", "bwa_gpu")


def write_fasta(path, contigs):
    lut = np.frombuffer(b"ACGT", np.uint8)
    with open(path, "wb") as f:
        for name, sym in contigs:
            f.write(f">{name}\n".encode())
            txt = lut[sym]
            full = len(txt) // 80 * 80
            body = np.empty((full // 80, 81), np.uint8)
            body[:, :80] = txt[:full].reshape(-1, 80)
            body[:, 80] = 10
            f.write(body.tobytes())
            if full < len(txt):
                f.write(txt[full:].tobytes() + b"\n")


def write_fastq(path, reads, prefix="r"):
    lut = np.frombuffer(b"ACGTN", np.uint8)
    n, l = reads.shape
    seq = lut[np.minimum(reads, 4)]
    qual = b"I" * l
    with open(path, "wb") as f:
        for s0 in range(0, n, 100_000):
            out = bytearray()
            for i in range(s0, min(n, s0 + 100_000)):
                out += b"@%s%d\n" % (prefix.encode(), i) + seq[i].tobytes() + b"\n+\n" + qual + b"\n"
            f.write(out)


def run_mem(binary, threads, batch, fa, fqs, env=None):
    """bwa mem -> (wall seconds, md5 of the SAM without @PG, stderr)"""
    t0 = time.perf_counter()
    p = subprocess.Popen([binary, "mem", "-t", str(threads), "-b", str(batch), fa] + fqs, stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                         env=dict(os.environ, **(env or {})))
    md5 = hashlib.md5()
    n_lines = 0
    import threading
    err = []
    th = threading.Thread(target=lambda: err.append(p.stderr.read()))
    th.start()
    tail = b""
    while True:
        chunk = p.stdout.read(1 << 22)
        if not chunk:
            break
        data = tail + chunk
        cut = data.rfind(b"\n") + 1
        tail = data[cut:]
        for line in data[:cut].split(b"\n")[:-1]:
            if not line.startswith(b"@PG"):
                md5.update(line + b"\n"); n_lines += 1
    p.wait(); th.join()
    wall = time.perf_counter() - t0
    if p.returncode != 0:
        raise RuntimeError(f"{binary} mem failed ({p.returncode}): {err[0][-2000:].decode(errors='replace')}")
    return wall, md5.hexdigest(), n_lines, err[0].decode(errors="replace")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bp", type=int, default=100_000_000)
    ap.add_argument("--contigs", type=int, default=4)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=101)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--paired", action="store_true")
    ap.add_argument("--threads", default="", help="comma list of -t values for the GPU runs (default: 1 and the core count)")
    ap.add_argument("--batches", default="64,1024,16384,65536")
    ap.add_argument("--handles", default="0", help="comma list of SMEM_GPU_ADAPTER_HANDLES values (0 = the adapter's own choice from the batch size)")
    ap.add_argument("--out", default="")
    ap.add_argument("--tmp", default="/dev/shm" if os.path.isdir("/dev/shm") else None)
    args = ap.parse_args()
    import torch
    fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
    sy = importlib.import_module("bwa-mem-harp2_b200.synth")
    ncores = len(os.sched_getaffinity(0))
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    with tempfile.TemporaryDirectory(dir=args.tmp) as tmp:
        t0 = time.time()
        ref = sy.make_reference(args.ref_bp, 11, dev)
        refn = ref.cpu().numpy()
        fa = os.path.join(tmp, "g.fa")
        per = (args.ref_bp + args.contigs - 1) // args.contigs
        write_fasta(fa, [(f"chr{k + 1}", refn[k * per:(k + 1) * per]) for k in range(args.contigs)])
        subprocess.run([REF, "fa2pac", "-f", fa, fa], check=True, capture_output=True)     # .pac .ann .amb (forward only, as `bwa index` leaves them)
        ix = fm.build_index(ref, sa_intv=32)
        ix.save(fa + ".bwt"); ix.save_sa(fa + ".sa")
        reads = sy.simulate_reads(ref, args.reads, args.read_len, args.err, seed=21, paired=args.paired).cpu().numpy()
        if args.paired:
            write_fastq(os.path.join(tmp, "r1.fq"), reads[0::2]); write_fastq(os.path.join(tmp, "r2.fq"), reads[1::2])
            fqs = [os.path.join(tmp, "r1.fq"), os.path.join(tmp, "r2.fq")]
        else:
            write_fastq(os.path.join(tmp, "r.fq"), reads)
            fqs = [os.path.join(tmp, "r.fq")]
        del ref, reads
        if dev == "cuda":
            torch.cuda.empty_cache()
        print(f"[dropin] files ready in {time.time() - t0:.1f}s ({args.ref_bp} bp, {args.reads} reads, {ncores} cores)", file=sys.stderr, flush=True)
        rows = []
        # the CPU path: the unmodified reference, -b 1 never offloads
        wall, md5_cpu, n_lines, err = run_mem(REF_TIMED if os.path.exists(REF_TIMED) else REF, ncores, 1, fa, fqs)
        m = re.search(r"\[ref_timed\] seed_s=([0-9.]+) calls=(\d+)", err)
        cpu = {"impl": "reference CPU path", "binary": "bwa_ref_timed mem", "t": ncores, "b": 1, "wall_s": round(wall, 3), "reads_per_s": round(args.reads / wall),
               "seed_s_sum_over_threads": float(m.group(1)) if m else None, "bwt_smem1_batched_calls": int(m.group(2)) if m else None, "sam_lines": n_lines}
        rows.append(cpu)
        print("[dropin]", cpu, file=sys.stderr, flush=True)
        tlist = [int(x) for x in args.threads.split(",")] if args.threads else sorted({1, ncores})
        for handles in [int(x) for x in args.handles.split(",")]:
            for t in tlist:
                for b in [int(x) for x in args.batches.split(",")]:
                    if t == 1 and b < 1024 and args.reads > 200_000:
                        continue                      # one thread, tiny batches: minutes of latency-bound calls that say nothing new
                    env = {"SMEM_GPU_ADAPTER_STATS": "1"}
                    if handles > 0:
                        env["SMEM_GPU_ADAPTER_HANDLES"] = str(handles)
                    wall, md5, n_l, err = run_mem(GPU, t, b, fa, fqs, env)
                    st = [l for l in err.splitlines() if "lists_from_cache" in l]
                    kv = dict(x.split("=") for x in st[-1].split()[1:]) if st else {}
                    row = {"impl": "drop-in (bwa_gpu mem: reference + GPU adapter)", "t": t, "b": b, "handles": int(kv.get("handles", handles)), "wall_s": round(wall, 3),
                           "reads_per_s": round(args.reads / wall), "sam_identical_to_cpu_path": md5 == md5_cpu and n_l == n_lines,
                           "seed_s_sum_over_threads": float(kv.get("adapter_s", "nan")), "startup_s_sum_over_threads": float(kv.get("startup_s", "nan")),
                           "gpu_call_s": float(kv.get("gpu_s", "nan")), "leader_gather_s": float(kv.get("gather_s", "nan")), "leader_launch_s": float(kv.get("launch_s", "nan")),
                           "leader_handback_s": float(kv.get("handback_s", "nan")), "grow_events": int(kv.get("grow_events", -1)),
                           "gpu_calls": int(kv.get("gpu_calls", -1)), "requests": int(kv.get("requests", -1)), "reads_sent": int(kv.get("reads_sent", -1)),
                           "lists_from_cache": int(kv.get("lists_from_cache", -1))}
                    if cpu["seed_s_sum_over_threads"] and row["seed_s_sum_over_threads"] > 0:
                        # same thread count only: the sums are over the worker threads of each run
                        row["seeding_speedup_vs_cpu_path"] = round(cpu["seed_s_sum_over_threads"] / row["seed_s_sum_over_threads"], 2) if t == ncores else None
                    row["wall_speedup_vs_cpu_path"] = round(cpu["wall_s"] / wall, 3)
                    rows.append(row)
                    print("[dropin]", row, file=sys.stderr, flush=True)
        out = {"tool": "tools/dropin_bench.py", "ref_bp": args.ref_bp, "reads": args.reads, "read_len": args.read_len, "paired": args.paired,
               "host_cores": ncores, "rows": rows,
               "notes": "seed_s_sum_over_threads = wall time inside bwt_smem1_batched added over the worker threads (CPU: the reference's own function, "
                        "oracle/timed_batched.c; GPU: the adapter, SMEM_GPU_ADAPTER_STATS); wall_s covers the whole `bwa mem` incl. file I/O, chaining, "
                        "Smith-Waterman and SAM output, which stay on the host"}
        txt = json.dumps(out)
        if args.out:
            with open(args.out, "w") as f:
                f.write(txt + "\n")
        print(txt)


if __name__ == "__main__":
    main()
