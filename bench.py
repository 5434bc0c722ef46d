#!/usr/bin/env python
"""bench.py -- SMEM-seeded 101 bp reads/s on N B200s (BASELINE.json metric), one JSON line on stdout.

A "step" is one pass of the hot path (smem_next2 to exhaustion for every read == the enumeration loop
of mem_insert_seed, bwamem.c:453-460) over one batch of simulated reads per GPU.

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, through the C ABI)
  python bench.py --impl reference [...]                      the reference's own CPU code on the host cores

Workload (BASELINE.json configs[2]): 3.1 Gbp synthetic reference (25 contigs x 124 Mbp), 1 M simulated
101 bp read pairs (= 2 M reads) per GPU per step, default seeding options (-k 19 -r 1.5, split_width 10).
The index is built on the GPU by bwa-mem-harp2_b200/fmindex.py (verified bit-for-bit against `bwa index`
in tests/), outside every timed region.  Multi-GPU: index replicated, reads sharded (weak scaling: every
rank seeds its own 2 M reads), no collective on the data path.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print("[bench]", *a, file=sys.stderr, flush=True)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-bp", type=int, default=int(os.environ.get("SMEM_BENCH_REF_BP", 3_100_000_000)))
    ap.add_argument("--contigs", type=int, default=25)
    ap.add_argument("--repeat-frac", type=float, default=0.0, help="robustness runs: overwrite this share of the reference with diverged copies of a few 300 bp repeat families (synth.add_repeat_families)")
    ap.add_argument("--reads", type=int, default=int(os.environ.get("SMEM_BENCH_READS", 2_000_000)), help="reads per GPU per step")
    ap.add_argument("--read-len", type=int, default=101)
    ap.add_argument("--err", type=float, default=0.01)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--blocks-per-sm", type=int, default=0)
    ap.add_argument("--l2-hot-min-intv", type=int, default=-1)
    ap.add_argument("--lanes", type=int, default=1, help="pipeline lanes per GPU of the end-to-end handle")
    ap.add_argument("--host-threads", type=int, default=0, help="host worker threads of the end-to-end legs, each with its own handle sharing one index (the reference's -t N pattern); 0 = 3, or 2 when the rank has fewer than 4 cores to itself")
    ap.add_argument("--no-extras", action="store_true", help="skip the extra legs (32-byte end-to-end, seeds / chains end-to-end, unique walks off, configs 4 and 5)")
    ap.add_argument("--pcie-probe", action="store_true", help="also time the step's copy volumes alone and next to a running seed kernel (on by default in multi-GPU runs)")
    ap.add_argument("--isa-shift", type=int, default=2, help="the unique-walk tables sample the inverse suffix array every 2^this positions")
    ap.add_argument("--spare-sms", type=int, default=-1)
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="smem_gpu_set_param on the device-resident handle (repeatable)")
    ap.add_argument("--no-repeat-filter", action="store_true", help="do not build / use the repeat filter of the re-seeding pass (DESIGN.md section 10)")
    ap.add_argument("--no-text-index", action="store_true", help="do not build / use the unique-walk tables")
    ap.add_argument("--text-index", action="store_true", help="(default when HBM has room) also build the unique-walk tables (text at 4 bits per base, 33-bit suffix array, sampled inverse: 6.2 bytes per text position) and seed with them (DESIGN.md section 10)")
    ap.add_argument("--record-bytes", type=int, default=0, help="interval records of the headline end-to-end leg: 11 or 12 bytes (0 = 12 unless a copy probe shows the link to be the limit)")
    ap.add_argument("--rf-kmer", type=int, default=0)
    ap.add_argument("--rf-log2-bits", type=int, default=0)
    ap.add_argument("--no-bind", action="store_true", help="multi-GPU runs: do not pin each rank to the CPUs local to its GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--sweep", default="", help="comma list of blocks_per_sm[:l2_hot_min_intv[:b_cap[::l2_mode]]] to time (stderr), e.g. 6,8:16384,9::17")
    ap.add_argument("--param-sweep", default="", help='time the seed kernel over values of library knobs: "name=v1,v2;name2=v1,v2" (stderr)')
    ap.add_argument("--probe", action="store_true", help="also run the random-access roofline sweep")
    ap.add_argument("--full-compare", action="store_true", help="compare every interval of the step with the oracle (config 2)")
    return ap.parse_args()


class ClockSampler:
    """SM clocks and throttle reasons DURING the timed region, sampled in-process through NVML every 5 ms
    (nvidia-smi needs longer to start than a timed region lasts; same counters as the B200_PROFILING.md recipe)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, cuda_index):
        self.samples, self.bits, self.max_mhz, self.power = [], 0, None, []
        self._stop = threading.Event()
        self._thread = None
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            log("NVML unavailable, no clock samples:", e)
            self.h = None

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
            except Exception:
                pass
            time.sleep(0.005)

    def start(self):
        if self.h is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self._thread:
            self._stop.set()
            self._thread.join(timeout=2)
        if self.samples:
            out.update(sm_mhz=float(np.median(self.samples)), samples=len(self.samples), power_w_max=max(self.power) if self.power else None,
                       reasons=sorted(n for b, n in self.REASONS.items() if self.bits & b))
        return out


def make_workload(args, rank, device, want_cfg4=False):
    """Reference + index on the GPU, this rank's reads on the host (pinned by the caller)."""
    import torch
    fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
    sy = importlib.import_module("bwa-mem-harp2_b200.synth")
    t0 = time.time()
    fwd = sy.make_reference(args.ref_bp, 13, device)
    if args.repeat_frac > 0:
        fwd = sy.add_repeat_families(fwd, args.repeat_frac, 17)
    ix = fm.build_index(fwd, sa_intv=0 if args.impl == 'reference' else 32)
    torch.cuda.synchronize()
    t_index = time.time() - t0
    reads = sy.simulate_reads(fwd, args.reads, args.read_len, args.err, seed=1000 + rank, paired=True)
    seq, offs = sy.to_batch(reads)
    cfg4 = None
    if want_cfg4:          # BASELINE configs[3]: 250 bp reads at 2 % errors on the same reference
        cfg4 = sy.to_batch(sy.simulate_reads(fwd, 500_000, 250, 0.02, seed=4000 + rank))
    sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
    pac = sg.pack_pac(fwd) if (args.text_index or not args.no_repeat_filter) else None          # the reference's .pac layout of the forward text
    del fwd, reads
    torch.cuda.empty_cache()
    return ix, seq, offs, t_index, pac, cfg4


def cpu_engine(ix_host):
    """(engine, kind): the reference's own objects if oracle/_ref travelled, else the C port."""
    from oracle.binding import Oracle, Reference
    try:
        return Reference(ix_host), "reference"
    except Exception as e:  # noqa: BLE001
        log("reference objects unavailable, using the oracle port:", e)
        return Oracle(ix_host), "port"


def cpu_time(engine, seq, offs, n, threads, opt):
    o = offs[: n + 1]
    return engine.time_collect(seq[: int(o[-1])], o, opt, threads)


_REAL_STDOUT = None


def emit(line: str):
    """The ONE JSON line goes to the process' original stdout; everything else (NCCL's version banner, library
    chatter) was re-routed to stderr at start-up."""
    out = _REAL_STDOUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def bind_near_gpu(local):
    """Multi-GPU runs: pin this rank (and the pinned host memory it is about to allocate) to the CPUs NVML reports as local
    to its GPU.  Without it the ranks of the far socket push their 0.9 GB per step across the socket interconnect and
    the end-to-end leg of all ranks together stalls at ~100 GB/s.  Returns the CPU list or None."""
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId("%08x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1} & set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception as e:       # binding is an optimisation, never a requirement
        log("cpu binding skipped:", repr(e))
    return None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference" and rank != 0:
        return 0                                   # rank 0 alone runs the CPU arm
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the seeding path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    use_dist = world > 1 and args.impl == "ours"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        host_group = dist.new_group(backend="gloo")       # host-side barrier for the legs in which one rank drives every GPU
    from oracle.binding import SeedOpt as OSeedOpt
    sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    cpu_binding = bind_near_gpu(local) if (world > 1 and args.impl == "ours" and not args.no_bind) else None
    workload = (f"{args.ref_bp / 1e9:.2f} Gbp synthetic reference ({args.contigs} contigs{', %.0f%% interspersed 300 bp repeat families' % (100 * args.repeat_frac) if args.repeat_frac > 0 else ''}), "
                f"{args.reads // 2} simulated {args.read_len}bp read pairs ({args.reads} reads) per GPU per step, "
                f"{args.err:.0%} substitutions, -k 19 -r 1.5 re-seeding on")
    config = {"workload": workload, "baseline_config": "configs[2]", "ref_bp": args.ref_bp, "reads_per_gpu_per_step": args.reads,
              "read_len": args.read_len, "l2": "inputs larger than L2 (index %.1f GB, reads %.0f MB)" %
              (args.ref_bp / 1e9, args.reads * args.read_len / 1e6), "parallelism": f"replicated index, reads sharded x{args.gpus}"}

    # unique-walk tables: 16.5 bytes per text position next to the index; on unless told otherwise
    # or the device is too small for them (results never depend on them)
    uw_need = int((0.5 + 32 / 7 * (1 + 0.5 ** args.isa_shift)) * 2 * args.ref_bp)
    uw_room = torch.cuda.get_device_properties(device).total_memory - 8 * 2 * args.ref_bp - (24 << 30)   # index + builder scratch + results
    if args.impl == "reference" or args.no_text_index:
        args.text_index = False
    elif not args.text_index:
        args.text_index = uw_need <= uw_room
        if not args.text_index:
            log(f"unique-walk tables skipped: {uw_need / 1e9:.0f} GB needed, {uw_room / 1e9:.0f} GB to spare on this device")
    log(f"rank {rank}/{world}: building workload ({args.ref_bp} bp)")
    want_cfg4 = (args.impl == "ours" and not args.no_extras and world == 1 and args.read_len == 101 and args.reads >= 1_000_000 and args.ref_bp >= 1_000_000_000)
    ix, seq, offs, t_index, pac, cfg4 = make_workload(args, rank, device, want_cfg4)
    log(f"index built on GPU in {t_index:.1f}s: seq_len={ix.seq_len} bwt_size={ix.bwt_size} primary={ix.primary}")
    n = len(offs) - 1

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
        ixh = fm.BwtIndex(ix.primary, ix.L2, ix.seq_len, ix.bwt_size, ix.words_numpy())
        eng, kind = cpu_engine(ixh)
        opt = OSeedOpt()
        probe = cpu_time(eng, seq, offs, min(n, 20_000), ncores, opt)
        per_step = min(n, max(10_000, int(probe["reads_per_s"] * min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))))
        for _ in range(args.warmup):
            cpu_time(eng, seq, offs, per_step, ncores, opt)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_time(eng, seq, offs, per_step, ncores, opt)
        dt = time.perf_counter() - t0
        v = per_step * args.steps / dt
        sample = f"first {per_step} reads of the step's batch per step, {ncores} host threads (kt_for)"
        emit(json.dumps({"impl": "reference", "metric": "smem_seeded_101bp_reads_per_sec", "value": v, "unit": "reads/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                          "config": config, "cpu_baseline": {"value": v, "unit": "reads/s", "cores": ncores, "kind": kind, "sample": sample},
                          "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return 0

    # ------------------------------------------------------------------ our arm (CUDA through the C ABI)
    import ctypes as C
    g = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=[local])
    if args.blocks_per_sm:
        g.set_param("blocks_per_sm", args.blocks_per_sm)
    for kv in args.set:
        k, v = kv.split("=")
        g.set_param(k, int(v))
    if args.l2_hot_min_intv >= 0:
        g.set_param("l2_hot_min_intv", args.l2_hot_min_intv)
    g.upload_index(ix)                      # device -> device copy of the packed bwt_t into the library's HBM buffer
    g.upload_sa(ix)                         # suffix-array samples: unique-walk tables, seed-level API
    lib = g.lib
    rf_info = None
    if pac is not None and not args.no_repeat_filter:
        torch.cuda.synchronize()
        tb0 = time.time()
        g.build_repeat_filter((pac, args.ref_bp), args.rf_kmer, args.rf_log2_bits)
        rf_info = {"kmer": g.get_param("rf_kmer"), "log2_bits": g.get_param("rf_log2_bits"), "bytes": 1 << (g.get_param("rf_log2_bits") - 3),
                   "build_s": round(time.time() - tb0, 3)}
        log("repeat filter:", rf_info)
    uw_info = None
    if pac is not None and args.text_index:
        torch.cuda.synchronize()
        tb0 = time.time()
        try:
            g.set_param("unique_walk_isa_shift", args.isa_shift)
            g.build_text_index((pac, args.ref_bp))
            uw_info = {"bytes": g.get_param("uw_table_bytes"), "isa_shift": args.isa_shift, "build_s": round(time.time() - tb0, 3)}
        except RuntimeError as e:           # optional tables: seeding runs without them (same results, FM extends instead)
            uw_info = {"skipped": str(e)}
        log("unique-walk tables:", uw_info)
    uw_on = bool(uw_info and "bytes" in uw_info)
    pac = None
    torch.cuda.empty_cache()
    # the step's batch in pinned host memory, in both forms the C ABI takes: bwa's one byte per base + CSR offsets
    # (smem_gpu_collect) and the compact wire format (smem_gpu_collect_packed: two bits per base, fixed stride)
    pseq = sg.PinnedArray(lib, (len(seq),), np.uint8); pseq.array[:] = seq
    poffs = sg.PinnedArray(lib, (n + 1,), np.int64); poffs.array[:] = offs
    tp0 = time.perf_counter()
    packed = sg.PackedReads(lib, seq, offs, pinned=True, threads=min(ncores, 16))
    t_pack = time.perf_counter() - tp0
    opt = sg.SeedOpt()

    # CPU baseline + full-size parity sample (rank 0 only)
    cpu_baseline, parity, bytes_per_read, executed = None, None, None, None
    if rank == 0:
        fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
        from oracle.binding import Oracle
        ixh = fm.BwtIndex(ix.primary, ix.L2, ix.seq_len, ix.bwt_size, ix.words_numpy())
        orc = Oracle(ixh)
        ns = min(n, 20_000)
        st = orc.collect(seq[: int(offs[ns])], offs[: ns + 1], OSeedOpt(), nthreads=ncores, stats=True)
        s = st["stats"]
        bytes_per_read = (64.0 * s["blocks"] + 128.0 * ns + 32.0 * s["intervals"]) / ns
        log(f"algorithmic work per read (oracle, {ns} reads): extends={s['extends'] / ns:.1f} blocks={s['blocks'] / ns:.1f} "
            f"intervals={s['intervals'] / ns:.2f} bytes={bytes_per_read:.0f}")
        if rf_info:
            # what the kernel's exact shortcuts leave of that work (oracle MODEL of them, exact window counts instead of
            # the device's hashed bit table): reported next to the algorithmic figure, which stays the reference algorithm's
            orc.lib.orc_set_skip_kmer(int(rf_info["kmer"])); orc.lib.orc_set_spec_walk(int(g.get_param("spec_walk"))); orc.lib.orc_set_unique_walk(int(uw_on and g.get_param("unique_walk")))
            try:
                s2 = orc.collect(seq[: int(offs[ns])], offs[: ns + 1], OSeedOpt(), nthreads=ncores, stats=True)["stats"]
            finally:
                orc.lib.orc_set_skip_kmer(0); orc.lib.orc_set_spec_walk(0); orc.lib.orc_set_unique_walk(0)
            executed = {"extends_per_read": s2["extends"] / ns, "blocks_per_read": s2["blocks"] / ns,
                        "bytes_per_read": (64.0 * s2["blocks"] + 128.0 * ns + 32.0 * s2["intervals"]) / ns,
                        "note": "same counting rule applied to the extends left after the repeat filter, the speculative walk and" + (" the unique walks (three gathers each; the sampled inverse suffix array adds 1.5 extends per walk on average, not counted)" if uw_on else " (not built here) the unique walks") + " (oracle model)"}
            log("executed work per read with the shortcuts (oracle model):", executed)
        got = g.collect(seq[: int(offs[ns])], offs[: ns + 1], opt)
        ok = (np.array_equal(got["read_off"], st["read_off"]) and np.array_equal(got["intv"], st["intv"])
              and np.array_equal(got["step"], st["step"]))
        # ... and the same reads through the compact wire format
        gp = g.collect_packed(sg.PackedReads(lib, seq[: int(offs[ns])], offs[: ns + 1]), opt)
        ok = ok and np.array_equal(gp["read_off"], st["read_off"]) and np.array_equal(gp["intv"], st["intv"])
        gp = g.collect_packed12(sg.PackedReads(lib, seq[: int(offs[ns])], offs[: ns + 1]), opt)
        ok = ok and np.array_equal(gp["read_off"], st["read_off"]) and np.array_equal(gp["intv"], st["intv"])
        if args.read_len <= 512:
            gp = g.collect_packed11(sg.PackedReads(lib, seq[: int(offs[ns])], offs[: ns + 1]), opt)
            ok = ok and np.array_equal(gp["read_off"], st["read_off"]) and np.array_equal(gp["intv"], st["intv"])
        parity = {"reads_checked": ns, "bit_exact": bool(ok), "checker": "oracle/liboracle.so",
                  "wire_formats_checked": ["bytes + bwtintv_t", "2-bit reads + 16-byte records", "2-bit reads + 12-byte records"] + (["2-bit reads + 11-byte records"] if args.read_len <= 512 else [])}
        if not ok:
            raise SystemExit("PARITY FAILURE: GPU intervals differ from the oracle on the bench workload")
        if not args.skip_cpu and world == 1:       # the CPU baseline is a single-GPU-run figure (it would idle the other ranks)
            eng, kind = cpu_engine(ixh)
            probe = cpu_time(eng, seq, offs, min(n, 20_000), ncores, OSeedOpt())
            m = min(n, max(20_000, int(probe["reads_per_s"] * args.cpu_seconds)))
            r = cpu_time(eng, seq, offs, m, ncores, OSeedOpt())
            one = cpu_time(eng, seq, offs, min(m, 20_000), 1, OSeedOpt())
            gk = (g.collect_packed11 if args.read_len <= 512 else g.collect_packed12)(sg.PackedReads(lib, seq[: int(offs[m])], offs[: m + 1]), opt)
            ck = orc.checksum(gk["intv"], gk["read_off"])
            parity.update(reads_checked=m, bit_exact=bool(ok and ck == r["checksum"]), checker=f"oracle + {kind} checksum (through the compact wire format)")
            if ck != r["checksum"]:
                raise SystemExit("PARITY FAILURE: checksum of GPU intervals differs from the CPU arm")
            cpu_baseline = {"value": r["reads_per_s"], "unit": "reads/s", "cores": ncores, "kind": kind,
                            "sample": f"first {m} reads of rank 0's batch, seeding only, {ncores} threads",
                            "one_thread_reads_per_s": one["reads_per_s"]}
            log("cpu baseline:", cpu_baseline)
            del eng
        del orc, ixh

    def sync():
        torch.cuda.synchronize()
        if use_dist:
            dist.barrier()

    # ---- device-resident leg: inputs staged in HBM before the timed region
    g.stage_packed(packed)
    if args.sweep:
        keep = (g.get_param("blocks_per_sm"), g.get_param("l2_hot_min_intv"), g.get_param("b_cap"))
        for item in args.sweep.split(","):
            f = item.split(":")
            b, hot = f[0], (f[1] if len(f) > 1 and f[1] else "0")
            g.set_param("l2_mode", int(f[4]) if len(f) > 4 else 0)
            g.set_param("blocks_per_sm", int(b)); g.set_param("l2_hot_min_intv", int(hot))
            g.set_param("b_cap", int(f[2]) if len(f) > 2 and f[2] else keep[2])
            ms = []
            for _ in range(4):
                g.run_collect(opt); ms.append(g.timing()["seed_kernel_ms"])
            log(f"sweep blocks_per_sm={b} l2_hot_min_intv={hot} b_cap={g.get_param('b_cap')} l2_mode={g.get_param('l2_mode')}: seed kernel {min(ms[1:]):.2f} ms -> {n / min(ms[1:]) / 1e3:.2f} M reads/s")
        g.set_param("blocks_per_sm", keep[0]); g.set_param("l2_hot_min_intv", keep[1]); g.set_param("b_cap", keep[2]); g.set_param("l2_mode", 0)
    if args.param_sweep:            # generic knob sweep on the resident batch: --param-sweep "name=v1,v2;other=v1,v2" (one knob at a time, restored after)
        for spec in args.param_sweep.split(";"):
            name, vals = spec.split("=")
            keep_v = g.get_param(name)
            for v in vals.split(","):
                g.set_param(name, int(v))
                ms = []
                for _ in range(4):
                    g.run_collect(opt); ms.append(g.timing()["seed_kernel_ms"])
                log(f"param sweep {name}={v}: seed kernel {min(ms[1:]):.3f} ms")
            g.set_param(name, keep_v)
    for _ in range(max(args.warmup, 3)):
        total = g.run_collect(opt)
    sampler = ClockSampler(local)
    sampler.start()
    seed_ms, dev_ms, launches = [], [], 0
    sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        total = g.run_collect(opt)
        t = g.timing()
        seed_ms.append(t["seed_kernel_ms"]); dev_ms.append(t["total_device_ms"]); launches += t["kernel_launches"]
    sync()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()
    overflow = g.timing()["overflow_reads"]
    if rf_info:                                          # untimed: how many re-seeding passes the filter proved void
        g.set_param("count_skips", 1); g.run_collect(opt)
        rf_info["pass2_skipped_per_step"] = g.get_param("pass2_skipped")
        if uw_on:
            uw_info["walks_per_step"] = g.get_param("unique_walks")
        g.set_param("count_skips", 0)
    if uw_on and not args.no_extras:                     # untimed: what the tables buy, in kernel time per GB of HBM
        g.set_param("unique_walk", 0)
        ms0 = []
        for _ in range(4):
            g.run_collect(opt); ms0.append(g.timing()["seed_kernel_ms"])
        g.set_param("unique_walk", 1)
        uw_info["seed_kernel_ms_without"] = float(np.mean(ms0[1:]))
        uw_info["seed_kernel_ms_with"] = float(np.mean(seed_ms))
        uw_info["ms_gained_per_GB"] = (uw_info["seed_kernel_ms_without"] - uw_info["seed_kernel_ms_with"]) / (uw_info["bytes"] / 1e9)

    # ---- end-to-end legs: host buffers in, host buffers out, copies inside the timed region, through the public one-call API.
    # Each of T worker threads owns a handle (sharing the first one's index copy, smem_gpu_share_index -- the reference's
    # -t N pattern) and its own pinned result buffers, and makes the public call for the steps it is given: one call's
    # copies overlap another call's kernels (the GPU's kernel turn).
    g_owner = g                             # owns the index and the tables; the end-to-end handles alias them
    cores_per_rank = max(1, ncores // max(1, world))
    T = args.host_threads if args.host_threads > 0 else (3 if cores_per_rank >= 4 else 2)
    workers = []
    for _ in range(T):
        w = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=[local] * max(1, args.lanes))
        if args.blocks_per_sm:
            w.set_param("blocks_per_sm", args.blocks_per_sm)
        if args.spare_sms >= 0:
            w.set_param("spare_sms", args.spare_sms)
        w.share_index_from(g_owner)
        workers.append(w)
    cap16 = total + total // 16 + 1024
    prec = [sg.PinnedArray(lib, (cap16, 2), np.uint64) for _ in range(T)]
    proff32 = [sg.PinnedArray(lib, (n + 1,), np.uint32) for _ in range(T)]
    prec12 = [sg.PinnedArray(lib, (cap16, 3), np.uint32) for _ in range(T)]          # 12-byte records (smem_intv12_t) + their exception lists
    prec11 = [sg.PinnedArray(lib, (cap16, 11), np.uint8) for _ in range(T)]          # 11-byte records (smem_intv11_t)
    exc_cap = max(4096, n // 4)
    pexc = [sg.PinnedArray(lib, (exc_cap,), sg.EXC_DTYPE) for _ in range(T)]

    def run_pool(step_fn, k_steps):
        """k_steps calls spread over the worker threads; a free worker takes the next step (kt_for_batch hands out batches
        the same way, kthread_batch.c:20-44).  Returns the last result of any worker."""
        res = [None] * T
        nxt = [0]
        lock = threading.Lock()
        err = []
        def work(t):
            try:
                while True:
                    with lock:
                        k = nxt[0]; nxt[0] += 1
                    if k >= k_steps:
                        break
                    res[t] = step_fn(t)
            except BaseException as e:  # noqa: BLE001
                err.append(e)
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        if err:
            raise SystemExit(f"end-to-end worker failed: {err[0]!r}")
        return next((r for r in res if r is not None), None)

    def check_rc(rc, t, what):
        if rc:
            raise RuntimeError(f"{what} failed: {rc} {lib.smem_gpu_last_error(workers[t].h).decode()}")

    def acc_reset():
        for w in workers:
            w.set_param("acc_reset", 1)

    def acc_read():
        """per-call wall time of the stages of the one-call form, averaged over this rank's calls"""
        c = max(1, sum(w.get_param("acc_calls") for w in workers))
        f = lambda k: sum(w.get_param(k) for w in workers) / c / 1e3
        st, fe = f("acc_stage_us"), f("acc_fetch_us")
        h2d, d2h = sum(w.get_param("acc_h2d_bytes") for w in workers) / c, sum(w.get_param("acc_d2h_bytes") for w in workers) / c
        return {"h2d_ms": st, "wait_for_kernel_turn_ms": f("acc_turn_us"), "kernels_ms": f("acc_run_us"), "d2h_ms": fe,
                "h2d_GBps": h2d / max(st, 1e-6) / 1e6, "d2h_GBps": d2h / max(fe, 1e-6) / 1e6}

    def timed_leg(step_fn, k_steps):
        run_pool(step_fn, 2 * T)
        acc_reset()
        sync()
        t1 = time.perf_counter()
        r = run_pool(step_fn, k_steps)
        sync()
        return time.perf_counter() - t1, r

    sh = importlib.import_module("bwa-mem-harp2_b200.sharding")      # the host-side multi-GPU plumbing: max / gather over ranks
    max_over_ranks = lambda x: sh.max_over_ranks(x, device)
    gather_ranks = lambda vals: sh.gather_rows(vals, device)

    # (1) the headline: compact wire format both ways: 2-bit reads in, 11-byte interval records out
    n_exc_last = [0] * T
    pos_bits = [0]
    def step_packed_small(t, fn, bufs, what):
        tot, ne, pb = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        rc = fn(workers[t].h, C.byref(packed.desc), C.byref(opt), C.c_void_p(bufs[t].array.ctypes.data), C.c_int64(cap16),
                C.c_void_p(proff32[t].array.ctypes.data), C.c_void_p(pexc[t].array.ctypes.data), C.c_int64(exc_cap), C.byref(ne), C.byref(pb), C.byref(tot))
        check_rc(rc, t, what)
        n_exc_last[t] = int(ne.value); pos_bits[0] = int(pb.value)
        return int(tot.value)
    # Which record size: the 12-byte form costs the GPU 0.15 ms less per step (no squeeze pass), the 11-byte form costs the link 8 % less.
    # So: 12 bytes unless the link is what limits -- all ranks copy one step's 12-byte volume at once from / to pinned memory, and if
    # the slowest rank needs longer for that than the GPU needs for the step, the 11-byte form is used (--record-bytes overrides).
    record_choice = {"record_bytes": 12, "why": "--record-bytes" if args.record_bytes else "one GPU: the link is not shared"}
    if args.record_bytes in (11, 12):
        record_choice["record_bytes"] = args.record_bytes
    elif world > 1 and args.read_len <= 512:
        h_in = torch.empty(packed.h2d_bytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(12 * total + 4 * n, dtype=torch.uint8).pin_memory()
        d_in = torch.empty_like(h_in, device=device); d_out = torch.empty_like(h_out, device=device)
        side = torch.cuda.Stream(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sync()
        with torch.cuda.stream(side):
            d_in.copy_(h_in, non_blocking=True); h_out.copy_(d_out, non_blocking=True)      # (warm-up)
            e0.record()
            for _ in range(4):
                d_in.copy_(h_in, non_blocking=True); h_out.copy_(d_out, non_blocking=True)
            e1.record()
        e1.synchronize()
        copy_ms = max_over_ranks(e0.elapsed_time(e1) / 4)
        step_ms = max_over_ranks(float(np.mean(dev_ms)))
        record_choice = {"record_bytes": 11 if copy_ms > 0.97 * step_ms else 12, "copy_ms_of_a_12_byte_step_all_ranks_at_once": round(copy_ms, 3),
                         "device_ms_per_step": round(step_ms, 3), "why": "the slowest rank's link time for a step of 12-byte records against the GPU's time for the step"}
        del h_in, h_out, d_in, d_out
        sync()
    use11 = record_choice["record_bytes"] == 11 and args.read_len <= 512      # (the 11-byte form needs query positions of <= 9 bits)
    if rank == 0:
        log("record size of the end-to-end leg:", record_choice)
    def step_packed11(t):
        return step_packed_small(t, lib.smem_gpu_collect_packed11, prec11, "smem_gpu_collect_packed11")
    def step_packed12(t):
        return step_packed_small(t, lib.smem_gpu_collect_packed12, prec12, "smem_gpu_collect_packed12")
    def step_packed(t):
        tot = C.c_int64(0)
        rc = lib.smem_gpu_collect_packed(workers[t].h, C.byref(packed.desc), C.byref(opt), C.c_void_p(prec[t].array.ctypes.data), C.c_int64(cap16),
                                         C.c_void_p(proff32[t].array.ctypes.data), C.byref(tot))
        check_rc(rc, t, "smem_gpu_collect_packed")
        return int(tot.value)
    dt_e2e, tot_e2e = timed_leg(step_packed11 if use11 else step_packed12, args.steps)
    te = workers[0].timing()
    stages = acc_read()
    stage_keys = list(stages)
    stages_per_rank = [dict(zip(stage_keys, (round(v, 3) for v in row))) for row in gather_ranks([stages[k] for k in stage_keys])]
    dt_e2e = max_over_ranks(dt_e2e)
    e2e_overflow = workers[0].timing()["overflow_reads"]
    # every worker's last result is the step's result: checksum of rank 0's against the device-resident run
    if rank == 0:
        a = g.fetch_packed(total)
        got12 = (sg.unpack_intv11(prec11[0].array[:tot_e2e], pos_bits[0], pexc[0].array[:n_exc_last[0]]) if use11 else
                 sg.unpack_intv12(prec12[0].array[:tot_e2e], pos_bits[0], pexc[0].array[:n_exc_last[0]]))
        if not (tot_e2e == total and np.array_equal(a["intv"], got12) and np.array_equal(a["read_off"], proff32[0].array.astype(np.int64))):
            raise SystemExit("PARITY FAILURE: end-to-end result differs from the device-resident run")
        del got12

    # ---- optional: every interval of the step against the oracle (BASELINE config 2)
    if args.full_compare and rank == 0:
        from oracle.binding import Oracle
        fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
        ixh = fm.BwtIndex(ix.primary, ix.L2, ix.seq_len, ix.bwt_size, ix.words_numpy())
        t0c = time.time()
        want = Oracle(ixh).collect(seq, offs, OSeedOpt(), nthreads=ncores)
        same = (np.array_equal(want["read_off"], proff32[0].array.astype(np.int64)) and np.array_equal(want["intv"], (sg.unpack_intv11(prec11[0].array[:tot_e2e], pos_bits[0], pexc[0].array[:n_exc_last[0]]) if use11 else
                                               sg.unpack_intv12(prec12[0].array[:tot_e2e], pos_bits[0], pexc[0].array[:n_exc_last[0]]))))
        log(f"full compare: {n} reads, {len(want['intv'])} intervals, bit_exact={same} (oracle {time.time() - t0c:.1f}s)")
        parity = dict(parity or {}, full_compare_reads=n, full_compare_intervals=int(len(want["intv"])), full_compare_bit_exact=bool(same))
        if not same:
            raise SystemExit("PARITY FAILURE in --full-compare")
        del want, ixh

    exc_e2e = int(n_exc_last[0])
    extras = {}
    if not args.no_extras:
        if not use11 and args.read_len <= 512:       # (1a) the same call with 11-byte records (8 % fewer bytes over the link, one more pass on the GPU)
            d11, tot11 = timed_leg(step_packed11, args.steps)
            t11 = workers[0].timing(); s11 = acc_read()
            d11 = max_over_ranks(d11)
            extras["e2e_11byte_records"] = {"value": world * n * args.steps / d11, "unit": "reads/s", "ms_per_step": d11 / args.steps * 1e3, "api": "smem_gpu_collect_packed11",
                                            "h2d_bytes_per_step": int(t11["h2d_bytes"]), "d2h_bytes_per_step": int(t11["d2h_bytes"]),
                                            "stages_rank0": {k: round(v, 3) for k, v in s11.items()}}
            log("e2e, 2-bit reads in / 11-byte records out:", extras["e2e_11byte_records"])
            if rank == 0 and not (tot11 == tot_e2e and np.array_equal(a["intv"], sg.unpack_intv11(prec11[0].array[:tot11], pos_bits[0], pexc[0].array[:n_exc_last[0]]))):
                raise SystemExit("PARITY FAILURE: 11-byte end-to-end result differs from the device-resident run")
        if use11:       # (1a) the same call with 12-byte records
            d12, tot12 = timed_leg(step_packed12, args.steps)
            t12 = workers[0].timing(); s12 = acc_read()
            d12 = max_over_ranks(d12)
            extras["e2e_12byte_records"] = {"value": world * n * args.steps / d12, "unit": "reads/s", "ms_per_step": d12 / args.steps * 1e3, "api": "smem_gpu_collect_packed12",
                                            "h2d_bytes_per_step": int(t12["h2d_bytes"]), "d2h_bytes_per_step": int(t12["d2h_bytes"]),
                                            "stages_rank0": {k: round(v, 3) for k, v in s12.items()}}
            log("e2e, 2-bit reads in / 12-byte records out:", extras["e2e_12byte_records"])
            if rank == 0 and not (tot12 == tot_e2e and np.array_equal(a["intv"], sg.unpack_intv12(prec12[0].array[:tot12], pos_bits[0], pexc[0].array[:n_exc_last[0]]))):
                raise SystemExit("PARITY FAILURE: 12-byte end-to-end result differs from the device-resident run")
        # (1b) the same call with 16-byte records (no exception list; 4 more bytes per interval over the link)
        d16, tot16 = timed_leg(step_packed, args.steps)
        t16 = workers[0].timing(); s16 = acc_read()
        d16 = max_over_ranks(d16)
        extras["e2e_16byte_records"] = {"value": world * n * args.steps / d16, "unit": "reads/s", "ms_per_step": d16 / args.steps * 1e3, "api": "smem_gpu_collect_packed",
                                        "h2d_bytes_per_step": int(t16["h2d_bytes"]), "d2h_bytes_per_step": int(t16["d2h_bytes"]),
                                        "stages_rank0": {k: round(v, 3) for k, v in s16.items()}}
        log("e2e, 2-bit reads in / 16-byte records out:", extras["e2e_16byte_records"])
        if rank == 0 and not (tot16 == tot_e2e and np.array_equal(a["rec"], prec[0].array[:tot16])):
            raise SystemExit("PARITY FAILURE: 16-byte end-to-end result differs from the device-resident run")
        # (2) bwa's own formats: one byte per base in, bwtintv_t out (the call the link-compatible adapter makes)
        pintvs = [sg.PinnedArray(lib, (total + 1024, 4), np.uint64) for _ in range(T)]
        proffs = [sg.PinnedArray(lib, (n + 1,), np.int64) for _ in range(T)]
        def step_bytes(t):
            tot = C.c_int64(0)
            rc = lib.smem_gpu_collect(workers[t].h, C.c_int64(n), pseq.array.ctypes.data_as(C.POINTER(C.c_uint8)),
                                      poffs.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(opt),
                                      pintvs[t].array.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(pintvs[t].array.shape[0]),
                                      proffs[t].array.ctypes.data_as(C.POINTER(C.c_int64)), None, C.byref(tot))
            check_rc(rc, t, "smem_gpu_collect")
            return int(tot.value)
        d2, _ = timed_leg(step_bytes, args.steps)
        tb = workers[0].timing(); sb = acc_read()
        d2 = max_over_ranks(d2)
        extras["e2e_bwtintv"] = {"value": world * n * args.steps / d2, "unit": "reads/s", "ms_per_step": d2 / args.steps * 1e3, "api": "smem_gpu_collect",
                                 "h2d_bytes_per_step": int(tb["h2d_bytes"]), "d2h_bytes_per_step": int(tb["d2h_bytes"]), "stages_rank0": {k: round(v, 3) for k, v in sb.items()}}
        log("e2e, byte reads in / bwtintv_t out:", extras["e2e_bwtintv"])
        for a_ in pintvs + proffs:
            a_.free()
        # (3) compact reads in -> mem_seed_t out (section 8f-1) and (4) -> chains after mem_chain_flt out (section 8f-3)
        seed_off = [sg.PinnedArray(lib, (n + 1,), np.int64) for _ in range(T)]
        pseeds = [sg.PinnedArray(lib, (4 * n, 2), np.int64) for _ in range(T)]
        chain_off = [sg.PinnedArray(lib, (n + 1,), np.int64) for _ in range(T)]
        pchains = [sg.PinnedArray(lib, (3 * n, 3), np.int64) for _ in range(T)]
        copt = sg.ChainOpt(100, 10000, 19, 0.5, 0.5, 1)
        def step_seeds(t, fetch=True):
            h = workers[t].h
            check_rc(lib.smem_gpu_stage_reads_packed(h, C.byref(packed.desc)), t, "stage_reads_packed")
            tot = C.c_int64(0)
            check_rc(lib.smem_gpu_run_collect(h, C.byref(opt), C.byref(tot)), t, "run_collect")
            rc = lib.smem_gpu_seeds(h, C.c_int(19), C.c_int64(10000), C.c_void_p(pseeds[t].array.ctypes.data) if fetch else None,
                                    C.c_int64(pseeds[t].array.shape[0] if fetch else 0), seed_off[t].array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(tot))
            if rc and not (rc == -5 and not fetch):
                check_rc(rc, t, "smem_gpu_seeds")
            return int(tot.value)
        d3, n_seeds = timed_leg(step_seeds, args.steps)
        d3 = max_over_ranks(d3)
        extras["e2e_seeds"] = {"value": world * n * args.steps / d3, "unit": "reads/s", "ms_per_step": d3 / args.steps * 1e3, "api": "smem_gpu_stage_reads_packed + run_collect + seeds",
                               "seeds_per_step_per_gpu": n_seeds, "h2d_bytes_per_step": packed.h2d_bytes, "d2h_bytes_per_step": n_seeds * 16 + (n + 1) * 8,
                               "note": "compact reads in -> mem_seed_t {rbeg,qbeg,len} out (min_seed_len 19, max_occ 10000)"}
        log("e2e, seeds out:", extras["e2e_seeds"])
        def step_chains(t):
            step_seeds(t, fetch=False)                   # seeds stay in HBM: only their count comes back
            nc, ns_ = C.c_int64(0), C.c_int64(0)
            rc = lib.smem_gpu_chains(workers[t].h, C.byref(copt), C.c_int64(int(ix.seq_len) // 2), C.c_void_p(pchains[t].array.ctypes.data), C.c_int64(pchains[t].array.shape[0]),
                                     C.c_void_p(pseeds[t].array.ctypes.data), C.c_int64(pseeds[t].array.shape[0]),
                                     chain_off[t].array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nc), C.byref(ns_))
            check_rc(rc, t, "smem_gpu_chains")
            return int(nc.value), int(ns_.value)
        d4, (n_chains, n_cseeds) = timed_leg(step_chains, args.steps)
        d4 = max_over_ranks(d4)
        extras["e2e_chains"] = {"value": world * n * args.steps / d4, "unit": "reads/s", "ms_per_step": d4 / args.steps * 1e3, "api": "... + smem_gpu_chains",
                                "chains_per_step_per_gpu": n_chains, "chain_seeds_per_step_per_gpu": n_cseeds, "h2d_bytes_per_step": packed.h2d_bytes,
                                "d2h_bytes_per_step": n_chains * 24 + n_cseeds * 16 + (n + 1) * 16,
                                "note": "compact reads in -> mem_chain_t after mem_chain_flt out (w 100, max_chain_gap 10000, mask_level 0.5, chain_drop_ratio 0.5)"}
        log("e2e, chains out:", extras["e2e_chains"])
        for a_ in seed_off + pseeds + chain_off + pchains:
            a_.free()

    # ---- the other multi-GPU form of the C ABI: ONE handle spanning all N devices of the box (a C host program's form; the
    # library shards the batch and stitches the CSR on the host).  Rank 0 drives it while the other ranks wait on the host.
    if use_dist and not args.no_extras:
        torch.cuda.synchronize()
        dist.barrier(group=host_group)
        if rank == 0:
            try:
                gs = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=list(range(world)))
                gs.upload_index(ix)                      # peer copies of the packed bwt_t to every GPU of the handle
                want_rec, want_off = prec[0].array[:tot_e2e].copy(), proff32[0].array.copy()
                def call_single():
                    tot = C.c_int64(0)
                    rc = lib.smem_gpu_collect_packed(gs.h, C.byref(packed.desc), C.byref(opt), C.c_void_p(prec[0].array.ctypes.data), C.c_int64(cap16),
                                                     C.c_void_p(proff32[0].array.ctypes.data), C.byref(tot))
                    if rc:
                        raise RuntimeError(f"single handle: {rc} {lib.smem_gpu_last_error(gs.h).decode()}")
                    return int(tot.value)
                call_single()
                ts_ = time.perf_counter()
                for _ in range(3):
                    tot_s = call_single()
                dts = (time.perf_counter() - ts_) / 3
                same = tot_s == tot_e2e and np.array_equal(prec[0].array[:tot_s], want_rec) and np.array_equal(proff32[0].array, want_off)
                extras["single_handle_all_devices"] = {"devices": world, "reads_per_call": n, "bit_exact_vs_one_gpu_run": bool(same), "ms_per_call": dts * 1e3,
                                                       "reads_per_s": n / dts, "note": "one smem_gpu_t over all GPUs of the box, one host call (strong scaling of ONE batch; "
                                                       "index only on the other GPUs, no accelerator tables)"}
                log("single handle over all devices:", extras["single_handle_all_devices"])
                if not same:
                    raise SystemExit("PARITY FAILURE: the N-device handle returns different intervals")
                gs.close()
            except RuntimeError as e:
                extras["single_handle_all_devices"] = {"skipped": str(e)}
        dist.barrier(group=host_group)

    # ---- where the end-to-end time goes when several GPUs share the host: the copy volumes of a step alone, and next to a running seed kernel
    pcie = None
    if args.pcie_probe or world > 1:
        h_in = torch.empty(packed.h2d_bytes, dtype=torch.uint8).pin_memory(); h_out = torch.empty(16 * tot_e2e + 4 * n, dtype=torch.uint8).pin_memory()
        d_in = torch.empty_like(h_in, device=device); d_out = torch.empty_like(h_out, device=device)
        side = torch.cuda.Stream(device)
        def copies(k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(side):
                e0.record()
                for _ in range(k):
                    d_in.copy_(h_in, non_blocking=True); h_out.copy_(d_out, non_blocking=True)
                e1.record()
            return e0, e1
        sync()
        e0, e1 = copies(10); e1.synchronize()
        alone = e0.elapsed_time(e1) / 10
        sync()
        stop = [False]
        def spin():
            while not stop[0]:
                g.run_collect(opt)
        thk = threading.Thread(target=spin); thk.start()
        time.sleep(0.05)
        e0, e1 = copies(10); e1.synchronize()
        stop[0] = True; thk.join()
        busy = e0.elapsed_time(e1) / 10
        nb = (h_in.numel() + h_out.numel()) / 1e6
        rows = gather_ranks([alone, busy])
        pcie = {"bytes_per_step_MB": nb, "copy_ms_alone_per_rank": [round(r[0], 3) for r in rows], "copy_ms_next_to_seed_kernel_per_rank": [round(r[1], 3) for r in rows],
                "box_GBps_alone": world * nb / max(r[0] for r in rows), "box_GBps_next_to_seed_kernel": world * nb / max(r[1] for r in rows),
                "note": "all ranks at once: H2D of a step's compact reads + D2H of its 16-byte records from/to pinned memory on a side stream, 10 rounds; second figure with the device-resident seed kernel looping"}
        log("pcie probe:", pcie)
        del h_in, h_out, d_in, d_out

    # ---- BASELINE configs 4 and 5 from the same invocation (single-GPU runs only)
    if not args.no_extras and world == 1 and rank == 0 and args.read_len == 101 and args.reads >= 1_000_000:
        # config 5: batch-size sweep, one smem_gpu_collect_packed call of B reads from host buffers, L2 evict_last hint on shallow levels off / on
        sweep = []
        for hot in (0, 16384):
            workers[0].set_param("l2_hot_min_intv", hot)
            for B in (64, 256, 1024, 4096, 16384, 65536, 262144, 1048576):
                sub = sg.Reads2(B, packed.seq2.ctypes.data, packed.stride, args.read_len, None, None, 0)
                def call():
                    tot = C.c_int64(0)
                    rc = lib.smem_gpu_collect_packed(workers[0].h, C.byref(sub), C.byref(opt), C.c_void_p(prec[0].array.ctypes.data), C.c_int64(cap16),
                                                     C.c_void_p(proff32[0].array.ctypes.data), C.byref(tot))
                    assert rc == 0, rc
                for _ in range(3):
                    call()
                reps = 20 if B <= 16384 else 5
                tb_ = time.perf_counter()
                for _ in range(reps):
                    call()
                ms = (time.perf_counter() - tb_) / reps * 1e3
                row = {"reads_per_call": B, "l2_hot_min_intv": hot, "latency_ms": round(ms, 3), "reads_per_s": round(B / ms * 1e3)}
                if hot == 0 and B <= 2048:
                    # the call the link-compatible adapter makes (bytes in, bwtintv_t out): batches this small take the library's latency path
                    li = np.empty((B * 32 + 1024, 4), np.uint64); lo_ = np.zeros(B + 1, np.int64); ls = np.empty(B * 32 + 1024, np.uint16)
                    def call_b():
                        tot = C.c_int64(0)
                        rc = lib.smem_gpu_collect(workers[0].h, C.c_int64(B), pseq.array.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                  poffs.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(opt), li.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                  C.c_int64(li.shape[0]), lo_.ctypes.data_as(C.POINTER(C.c_int64)), ls.ctypes.data_as(C.POINTER(C.c_uint16)), C.byref(tot))
                        assert rc == 0, rc
                    for _ in range(3):
                        call_b()
                    tb_ = time.perf_counter()
                    for _ in range(20):
                        call_b()
                    row["latency_ms_byte_api_latency_path"] = round((time.perf_counter() - tb_) / 20 * 1e3, 3)
                sweep.append(row)
        workers[0].set_param("l2_hot_min_intv", 0)
        extras["config5_batch_sweep"] = {"api": "smem_gpu_collect_packed, one handle, host buffers", "rows": sweep,
                                         "note": "reads without ambiguous bases only (the sub-batches reuse the step's records without its exception list)" if packed.n_amb else None}
        log("config 5:", sweep)
        # config 4: 250 bp reads at 2 % errors on the same index, re-seeding on (-k 19 -r 1.5, split_width 10)
        if cfg4 is not None:
            from oracle.binding import Oracle
            seq4, offs4 = cfg4
            n4 = len(offs4) - 1
            g4 = sg.SmemGpu(max_batch_reads=n4, max_read_len=250, devices=[local])
            g4.share_index_from(g_owner)
            pk4 = sg.PackedReads(lib, seq4, offs4, pinned=True, threads=min(ncores, 16))
            fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
            ixh = fm.BwtIndex(ix.primary, ix.L2, ix.seq_len, ix.bwt_size, ix.words_numpy())
            orc = Oracle(ixh)
            m4 = min(n4, 20_000)
            tc0 = time.perf_counter()
            want4 = orc.collect(seq4[: int(offs4[m4])], offs4[: m4 + 1], OSeedOpt(), nthreads=ncores, stats=True)
            cpu4 = m4 / (time.perf_counter() - tc0)
            got4 = g4.collect_packed(sg.PackedReads(lib, seq4[: int(offs4[m4])], offs4[: m4 + 1]), opt)
            ok4 = np.array_equal(got4["intv"], want4["intv"]) and np.array_equal(got4["read_off"], want4["read_off"])
            if not ok4:
                raise SystemExit("PARITY FAILURE on the 250 bp configuration")
            s4 = want4["stats"]
            g4.stage_packed(pk4)
            ms4 = []
            for _ in range(6):
                tot4 = g4.run_collect(opt); ms4.append(g4.timing()["total_device_ms"])
            rec4 = sg.PinnedArray(lib, (tot4 + 1024, 2), np.uint64); off4 = sg.PinnedArray(lib, (n4 + 1,), np.uint32)
            def call4():
                tot = C.c_int64(0)
                rc = lib.smem_gpu_collect_packed(g4.h, C.byref(pk4.desc), C.byref(opt), C.c_void_p(rec4.array.ctypes.data), C.c_int64(rec4.array.shape[0]),
                                                 C.c_void_p(off4.array.ctypes.data), C.byref(tot))
                assert rc == 0, (rc, lib.smem_gpu_last_error(g4.h))
            call4()
            tb_ = time.perf_counter()
            for _ in range(5):
                call4()
            e4 = (time.perf_counter() - tb_) / 5
            extras["config4_250bp"] = {"workload": f"{n4} simulated 250 bp reads at 2 % substitutions on the same index, -k 19 -r 1.5 (split_len 28, split_width 10)",
                                       "value": n4 / (float(np.mean(ms4[1:])) * 1e-3), "unit": "reads/s", "device_ms_per_step": float(np.mean(ms4[1:])),
                                       "e2e_one_handle_reads_per_s": n4 / e4, "intervals_per_read": tot4 / n4, "overflow_reads": int(g4.timing()["overflow_reads"]),
                                       "algorithmic_bytes_per_read": (64.0 * s4["blocks"] + 128.0 * m4 + 32.0 * s4["intervals"]) / m4,
                                       "parity": {"reads_checked": m4, "bit_exact": bool(ok4), "checker": "oracle/liboracle.so, entry by entry"},
                                       "cpu_port_reads_per_s": cpu4, "cpu_cores": ncores}
            log("config 4:", extras["config4_250bp"])
            rec4.free(); off4.free(); g4.close(); del orc, ixh

    probe = None
    if args.probe and rank == 0:
        probe = {}
        for bb in (64, 32):
            for chains in (512, 1024, 2048):
                probe[f"{bb}B_x{chains}"] = round(g.gather_roofline(bb, 0, chains, 1000), 1)
        log("random-access probe GB/s:", probe)
    rand64 = rand64_split = rand64_bulk = None
    if rank == 0:
        # the roofline of the access pattern: dependent 64 B gathers over the whole index, one coalesced request
        # per block (lane pair, 2 x 32 B) -- and, for reference, the same gathers issued as two per-lane requests
        g.set_param("probe_variant", 10); rand64 = g.gather_roofline(64, 0, 2048, 1000)
        g.set_param("probe_variant", 0); rand64_split = g.gather_roofline(64, 0, 2048, 1000)
        try:      # the same gathers through the bulk-copy engine: one cp.async.bulk of 64 B per lane into shared memory behind an mbarrier
            g.set_param("probe_variant", 20); rand64_bulk = g.gather_roofline(64, 0, 2048, 1000)
        except Exception as e:  # noqa: BLE001
            rand64_bulk = None; log("bulk-copy probe failed:", e)
        g.set_param("probe_variant", 0)

    # ---- max over ranks
    dt, seed_avg_ms = max_over_ranks(dt), max_over_ranks(float(np.mean(seed_ms)))

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic, traffic_src = None, None
        try:   # dram__bytes_read + dram__bytes_write of seed_kernel from the committed `ncu --set full` capture of this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "seed_traffic.json")))
            if not uw_on:
                tr = tr.get("without_text_index", {})          # the capture of the kernel without the unique-walk tables
            if tr.get("ref_bp") == args.ref_bp and tr.get("reads") == args.reads and tr.get("read_len") == args.read_len:
                traffic = tr["dram_bytes_per_launch"]
                traffic_src = "static: profiles/seed_traffic.json (" + tr.get("capture", "ncu --set full capture of this workload") + "), not measured in this run"
        except Exception:
            pass
        value = world * n * args.steps / dt
        achieved = n * bytes_per_read / (seed_avg_ms * 1e-3) / 1e9
        exec_gbs = n * executed["bytes_per_read"] / (seed_avg_ms * 1e-3) / 1e9 if executed else None
        # the bound that actually holds for this access pattern: DRAM-missing requests per second.
        # requests of the kernel = ncu DRAM bytes / 64 (the sector pairs a lane pair gathers); ceiling = the coalesced 64 B probe
        request_rate, frac_req = None, None
        if traffic and rand64:
            rq = traffic / 64.0 / (seed_avg_ms * 1e-3) / 1e9
            frac_req = rq / (rand64 / 64.0)
            request_rate = {"achieved_G_per_s": rq, "ceiling_G_per_s": rand64 / 64.0, "frac": frac_req,
                            "note": "DRAM requests of 64 B per second: ncu dram bytes of the committed capture / 64 / live kernel time, against the coalesced gather probe measured in this run"}
        out = {
            "metric": "smem_seeded_101bp_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": world * n * args.steps / dt_e2e, "unit": "reads/s", "h2d_bytes_per_step": int(te["h2d_bytes"]),
                    "d2h_bytes_per_step": int(te["d2h_bytes"]), "ms_per_step": dt_e2e / args.steps * 1e3,
                    "api": ("smem_gpu_collect_packed11 (2-bit reads in, 11-byte interval records + exception list out; include/smem_gpu.h)" if use11 else
                            "smem_gpu_collect_packed12 (2-bit reads in, 12-byte interval records + exception list out; include/smem_gpu.h)"),
                    "exceptions_per_step": exc_e2e, "record_pos_bits": pos_bits[0], "record_choice": record_choice,
                    "pipeline_lanes_per_gpu": args.lanes, "host_threads": T, "intervals": int(tot_e2e),
                    "host_pack_s_per_step_untimed": round(t_pack, 4),
                    "stages_per_rank": stages_per_rank},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "frac_executed": exec_gbs / rand64 if (exec_gbs and rand64) else None,
                         "frac_dram_requests": frac_req,
                         "frac_notes": "frac = the REFERENCE algorithm's bytes (SURVEY 8d) / kernel time / streaming peak; frac_executed = bytes of the extends the kernel really "
                                       "executes (oracle model of its exact shortcuts) / kernel time / the coalesced random-gather probe; frac_dram_requests = DRAM requests / s "
                                       "against the same probe's request rate",
                         "kernel": "seed_kernel<COLLECT>",
                         "kernel_ms": seed_avg_ms, "algorithmic_bytes_per_read": bytes_per_read, "executed": executed,
                         "request_rate": request_rate,
                         "random_access_peak": rand64, "frac_of_random_access": achieved / rand64 if rand64 else None,
                         "random_access_peak_two_requests": rand64_split, "random_access_bulk_copy_64B_per_lane": rand64_bulk,
                         "random_access_note": "dependent 64 B gathers over the whole index (smem_gpu_gather_roofline): one coalesced "
                                               "request per block by a lane pair; L2 hits on hot blocks let the kernel exceed it"},
            "cpu_baseline": cpu_baseline, "parity": parity, "clocks": clocks,
            "intervals_per_step_per_gpu": int(total), "overflow_reads": int(overflow), "index_build_s": t_index,
            "blocks_per_sm": g.get_param("blocks_per_sm"), "l2_hot_min_intv": g.get_param("l2_hot_min_intv"),
            "repeat_filter": rf_info, "unique_walk_tables": uw_info, "hbm_table_bytes": g.get_param("table_bytes"), "hbm_index_bytes": g.get_param("index_bytes"),
            "cpu_binding": ("%d cpus local to the GPU (NVML): %d..%d" % (len(cpu_binding), cpu_binding[0], cpu_binding[-1])) if cpu_binding else None,
            "device_ms_per_step": float(np.mean(dev_ms)),
        }
        out.update(extras)
        # throughput through the link-compatible bwt_smem1_batched (reference `bwa mem -t T -b B` with only that symbol swapped):
        # measured by tools/dropin_bench.py on a GPU box (it runs the reference binaries, which this benchmark does not), committed
        dropin_rows = []
        for f in ("r2_dropin_100Mbp.json",):
            try:
                dj = json.load(open(os.path.join(ROOT, "profiles", f)))
                for r in dj["rows"]:
                    if r.get("impl", "").startswith("drop-in") and not any(x["b"] == r["b"] and x["t"] == r["t"] for x in dropin_rows):
                        dropin_rows.append({k: r[k] for k in ("t", "b", "handles", "reads_per_s", "gpu_calls", "seeding_speedup_vs_cpu_path", "wall_speedup_vs_cpu_path",
                                                            "sam_identical_to_cpu_path") if k in r})
            except (OSError, ValueError, KeyError):
                pass
        if dropin_rows:
            out["dropin"] = {"source": "static: profiles/r2_dropin_100Mbp.json (tools/dropin_bench.py, 100 Mbp reference, 2 M x 101 bp reads, 16 host cores); "
                                       "not measured in this run", "rows": sorted(dropin_rows, key=lambda r: r["b"])}
        if pcie:
            out["pcie_probe"] = pcie
        if probe:
            out["random_access_probe_gbs"] = probe
        emit(json.dumps(out))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
