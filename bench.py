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
    ap.add_argument("--host-threads", type=int, default=2, help="host worker threads of the end-to-end leg, each with its own handle sharing one index (the reference's -t N pattern)")
    ap.add_argument("--spare-sms", type=int, default=-1)
    ap.add_argument("--fast", action="store_true", help="EXPERIMENTAL: build the k-mer count pyramid and seed through the table-driven kernel (DESIGN.md section 9)")
    ap.add_argument("--direct-levels", type=int, default=-1, help="depth of the direct k-mer tables (pyramid reaches +5); -1 = from the text length")
    ap.add_argument("--fast-blocks-per-sm", type=int, default=0)
    ap.add_argument("--fast-slots", type=int, default=0)
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="smem_gpu_set_param on the device-resident handle (repeatable)")
    ap.add_argument("--no-repeat-filter", action="store_true", help="do not build / use the repeat filter of the re-seeding pass (DESIGN.md section 10)")
    ap.add_argument("--no-text-index", action="store_true", help="do not build / use the unique-walk tables")
    ap.add_argument("--text-index", action="store_true", help="(default when HBM has room) also build the unique-walk tables (text at 4 bits per base, full SA, inverse SA: 16.5 bytes per text position) and seed with them (DESIGN.md section 10)")
    ap.add_argument("--rf-kmer", type=int, default=0)
    ap.add_argument("--rf-log2-bits", type=int, default=0)
    ap.add_argument("--no-bind", action="store_true", help="multi-GPU runs: do not pin each rank to the CPUs local to its GPU")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--sweep", default="", help="comma list of blocks_per_sm[:l2_hot_min_intv[:b_cap[:reuse[:l2_mode]]]] to time (stderr), e.g. 6,8:16384,9::17")
    ap.add_argument("--probe", action="store_true", help="also run the random-access roofline sweep")
    ap.add_argument("--batch-sweep", action="store_true", help="BASELINE config 5: latency/throughput of smem_gpu_collect for 64..1M reads per call, L2 hint off/on")
    ap.add_argument("--seeds", action="store_true", help="also time collect + smem_gpu_seeds (section 8f-1: intervals -> mem_seed_t on device)")
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


def make_workload(args, rank, device):
    """Reference + index on the GPU, this rank's reads on the host (pinned by the caller)."""
    import torch
    fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
    sy = importlib.import_module("bwa-mem-harp2_b200.synth")
    t0 = time.time()
    fwd = sy.make_reference(args.ref_bp, 13, device)
    if args.repeat_frac > 0:
        fwd = sy.add_repeat_families(fwd, args.repeat_frac, 17)
    ix = fm.build_index(fwd, sa_intv=32 if (getattr(args, 'seeds', False) or getattr(args, 'text_index', False)) else 0)
    torch.cuda.synchronize()
    t_index = time.time() - t0
    reads = sy.simulate_reads(fwd, args.reads, args.read_len, args.err, seed=1000 + rank, paired=True)
    seq, offs = sy.to_batch(reads)
    sg = importlib.import_module("bwa-mem-harp2_b200.smem_gpu")
    pac = sg.pack_pac(fwd) if (args.fast or args.text_index or not args.no_repeat_filter) else None          # the reference's .pac layout of the forward text
    del fwd, reads
    torch.cuda.empty_cache()
    return ix, seq, offs, t_index, pac


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
    uw_need = int(16.5 * 2 * args.ref_bp)
    uw_room = torch.cuda.get_device_properties(device).total_memory - 8 * 2 * args.ref_bp - (24 << 30)   # index + builder scratch + results
    if args.impl == "reference" or args.no_text_index or args.fast:
        args.text_index = False
    elif not args.text_index:
        args.text_index = uw_need <= uw_room
        if not args.text_index:
            log(f"unique-walk tables skipped: {uw_need / 1e9:.0f} GB needed, {uw_room / 1e9:.0f} GB to spare on this device")
    log(f"rank {rank}/{world}: building workload ({args.ref_bp} bp)")
    ix, seq, offs, t_index, pac = make_workload(args, rank, device)
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
    g = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=[local])
    if args.blocks_per_sm:
        g.set_param("blocks_per_sm", args.blocks_per_sm)
    for kv in args.set:
        k, v = kv.split("=")
        g.set_param(k, int(v))
    if args.l2_hot_min_intv >= 0:
        g.set_param("l2_hot_min_intv", args.l2_hot_min_intv)
    g.upload_index(ix)                      # device -> device copy of the packed bwt_t into the library's HBM buffer
    lib = g.lib
    fast_info = None
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
        g.upload_sa(ix)
        try:
            g.build_text_index((pac, args.ref_bp))
            uw_info = {"bytes": 16 * (int(ix.seq_len) + 4) + int(ix.seq_len) // 2, "build_s": round(time.time() - tb0, 3)}
        except RuntimeError as e:           # optional tables: seeding runs without them (same results, FM extends instead)
            uw_info = {"skipped": str(e)}
        log("unique-walk tables:", uw_info)
    if pac is not None and args.fast:
        import ctypes as C
        DL = args.direct_levels
        if DL < 0:
            DL = 2
            while 4 ** (DL + 5) < 1.3 * ix.seq_len and DL < 13:
                DL += 1
        torch.cuda.synchronize()
        tb0 = time.time()
        g._check(lib.smem_gpu_build_kmer_tables(g.h, C.c_void_p(pac.data_ptr()), C.c_int64(args.ref_bp), C.c_int(local), C.c_int(DL)))
        fast_info = {"direct_levels": DL, "deepest_level": DL + 5, "build_s": time.time() - tb0,
                     "table_bytes": 4 ** (DL + 5) + 4 ** (DL + 4) + (4 ** (DL + 1) - 4) // 3 * 4 + (4 ** (DL + 2) - 4) // 3 * 8}
        g.set_param("fast", 1)
        if args.fast_blocks_per_sm:
            g.set_param("fast_blocks_per_sm", args.fast_blocks_per_sm)
        if args.fast_slots:
            g.set_param("fast_slots", args.fast_slots)
        log("k-mer count pyramid:", fast_info)
    pac = None
    torch.cuda.empty_cache()
    # pinned host batch + pinned result buffers for the end-to-end leg
    pseq = sg.PinnedArray(lib, (len(seq),), np.uint8); pseq.array[:] = seq
    poffs = sg.PinnedArray(lib, (n + 1,), np.int64); poffs.array[:] = offs
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
            # what the kernel's two exact shortcuts leave of that work (oracle MODEL of them, exact window counts instead of
            # the device's hashed bit table): reported next to the algorithmic figure, which stays the reference algorithm's
            import ctypes as C
            uw_on = bool(uw_info and "bytes" in uw_info and g.get_param("unique_walk"))
            orc.lib.orc_set_skip_kmer(int(rf_info["kmer"])); orc.lib.orc_set_spec_walk(int(g.get_param("spec_walk"))); orc.lib.orc_set_unique_walk(int(uw_on))
            try:
                s2 = orc.collect(seq[: int(offs[ns])], offs[: ns + 1], OSeedOpt(), nthreads=ncores, stats=True)["stats"]
            finally:
                orc.lib.orc_set_skip_kmer(0); orc.lib.orc_set_spec_walk(0); orc.lib.orc_set_unique_walk(0)
            executed = {"extends_per_read": s2["extends"] / ns, "blocks_per_read": s2["blocks"] / ns,
                        "bytes_per_read": (64.0 * s2["blocks"] + 128.0 * ns + 32.0 * s2["intervals"]) / ns,
                        "note": "same counting rule applied to the extends left after the repeat filter, the speculative walk and" + (" the unique walks, three gathers each" if uw_on else " (not built here) the unique walks") + " (oracle model)"}
            log("executed work per read with the shortcuts (oracle model):", executed)
        got = g.collect(seq[: int(offs[ns])], offs[: ns + 1], opt)
        ok = (np.array_equal(got["read_off"], st["read_off"]) and np.array_equal(got["intv"], st["intv"])
              and np.array_equal(got["step"], st["step"]))
        parity = {"reads_checked": ns, "bit_exact": bool(ok), "checker": "oracle/liboracle.so"}
        if not ok:
            raise SystemExit("PARITY FAILURE: GPU intervals differ from the oracle on the bench workload")
        if not args.skip_cpu and world == 1:       # the CPU baseline is a single-GPU-run figure (it would idle the other ranks)
            eng, kind = cpu_engine(ixh)
            probe = cpu_time(eng, seq, offs, min(n, 20_000), ncores, OSeedOpt())
            m = min(n, max(20_000, int(probe["reads_per_s"] * args.cpu_seconds)))
            r = cpu_time(eng, seq, offs, m, ncores, OSeedOpt())
            one = cpu_time(eng, seq, offs, min(m, 20_000), 1, OSeedOpt())
            gk = g.collect(seq[: int(offs[m])], offs[: m + 1], opt, want_step=False)
            ck = orc.checksum(gk["intv"], gk["read_off"])
            parity.update(reads_checked=m, bit_exact=bool(ok and ck == r["checksum"]), checker=f"oracle + {kind} checksum")
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
    g.stage(pseq.array, poffs.array)
    if args.sweep:
        keep = (g.get_param("blocks_per_sm"), g.get_param("l2_hot_min_intv"), g.get_param("b_cap"))
        for item in args.sweep.split(","):
            f = item.split(":")
            b, hot = f[0], (f[1] if len(f) > 1 and f[1] else "0")
            g.set_param("reuse", int(f[3]) if len(f) > 3 and f[3] else 0)
            g.set_param("l2_mode", int(f[4]) if len(f) > 4 else 0)
            g.set_param("blocks_per_sm", int(b)); g.set_param("l2_hot_min_intv", int(hot))
            g.set_param("b_cap", int(f[2]) if len(f) > 2 and f[2] else keep[2])
            ms = []
            for _ in range(4):
                g.run_collect(opt); ms.append(g.timing()["seed_kernel_ms"])
            log(f"sweep blocks_per_sm={b} l2_hot_min_intv={hot} b_cap={g.get_param('b_cap')} reuse={g.get_param('reuse')} l2_mode={g.get_param('l2_mode')}: seed kernel {min(ms[1:]):.2f} ms -> {n / min(ms[1:]) / 1e3:.2f} M reads/s")
        g.set_param("blocks_per_sm", keep[0]); g.set_param("l2_hot_min_intv", keep[1]); g.set_param("b_cap", keep[2]); g.set_param("reuse", 0); g.set_param("l2_mode", 0)
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
        if uw_info is not None and "bytes" in uw_info:
            uw_info["walks_per_step"] = g.get_param("unique_walks")
        g.set_param("count_skips", 0)

    # ---- end-to-end leg: host buffers in, host buffers out, copies inside the timed region.  The public call is
    # smem_gpu_collect on a handle with `--lanes` pipeline lanes on this GPU (shards overlap H2D / kernels / D2H).
    escaped = g.get_param("escaped_reads")
    g_owner = g                             # owns the index and the tables; the end-to-end handles alias them
    if args.lanes > 1:
        g = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=[local] * args.lanes)
        if args.blocks_per_sm:
            g.set_param("blocks_per_sm", args.blocks_per_sm)
        if args.spare_sms >= 0:
            g.set_param("spare_sms", args.spare_sms)
        if args.fast_blocks_per_sm:
            g.set_param("fast_blocks_per_sm", args.fast_blocks_per_sm)
        g.set_param("fast", 1 if args.fast else 0)
        g.share_index_from(g_owner)
    # Each of `--host-threads` worker threads owns a handle (sharing the first one's index copy, smem_gpu_share_index)
    # and its own pinned result buffers, and makes the public call for the steps it is given: step k's copies overlap
    # step k+1's kernels.  Every step still copies its inputs H2D and its results D2H inside the timed region.
    import ctypes as C
    T = max(1, args.host_threads)
    workers = [g]
    for _ in range(1, T):
        w = sg.SmemGpu(max_batch_reads=n, max_read_len=args.read_len, devices=[local] * max(1, args.lanes))
        if args.blocks_per_sm:
            w.set_param("blocks_per_sm", args.blocks_per_sm)
        if args.fast_blocks_per_sm:
            w.set_param("fast_blocks_per_sm", args.fast_blocks_per_sm)
        w.set_param("fast", 1 if args.fast else 0)
        w.share_index_from(g_owner)
        workers.append(w)
    pintvs = [sg.PinnedArray(lib, (total + 1024, 4), np.uint64) for _ in range(T)]
    proffs = [sg.PinnedArray(lib, (n + 1,), np.int64) for _ in range(T)]
    pintv, proff = pintvs[0], proffs[0]

    def e2e_step(t=0):
        tot = C.c_int64(0)
        rc = lib.smem_gpu_collect(workers[t].h, C.c_int64(n), pseq.array.ctypes.data_as(C.POINTER(C.c_uint8)),
                                  poffs.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(opt),
                                  pintvs[t].array.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(pintvs[t].array.shape[0]),
                                  proffs[t].array.ctypes.data_as(C.POINTER(C.c_int64)), None, C.byref(tot))
        if rc:
            raise SystemExit(f"smem_gpu_collect failed: {rc} {lib.smem_gpu_last_error(workers[t].h).decode()}")
        return int(tot.value)

    def e2e_run(k_steps):
        """k_steps calls of the public API spread over the worker threads; returns the interval count of a step."""
        if T == 1:
            r = 0
            for _ in range(k_steps):
                r = e2e_step(0)
            return r
        res = [0] * T
        nxt = [0]
        lock = threading.Lock()
        def work(t):
            # a free worker takes the next step (kt_for_batch hands out batches the same way, kthread_batch.c:20-44): with a
            # fixed striping the worker that happens to come second can be left with the last two steps, back to back
            while True:
                with lock:
                    k = nxt[0]; nxt[0] += 1
                if k >= k_steps:
                    break
                res[t] = e2e_step(t)
        th = [threading.Thread(target=work, args=(t,)) for t in range(T)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        return max(res)

    e2e_run(2 * T)
    sync()
    t1 = time.perf_counter()
    tot_e2e = e2e_run(args.steps)
    sync()
    dt_e2e = time.perf_counter() - t1
    te = g.timing()

    # ---- optional: the seed-level API (section 8f-1): stage -> collect -> seeds -> D2H of mem_seed_t only
    seeds_leg = None
    if args.seeds:
        g.upload_sa(ix)
        seed_off = sg.PinnedArray(lib, (n + 1,), np.int64)
        pseeds = sg.PinnedArray(lib, (8 * n, 2), np.int64)
        def seeds_step():
            g.stage(pseq.array, poffs.array)
            g.run_collect(opt)
            tot = C.c_int64(0)
            rc = lib.smem_gpu_seeds(g.h, C.c_int(19), C.c_int64(10000), C.c_void_p(pseeds.array.ctypes.data), C.c_int64(pseeds.array.shape[0]),
                                    seed_off.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(tot))
            assert rc == 0, (rc, lib.smem_gpu_last_error(g.h))
            return int(tot.value)
        for _ in range(2):
            seeds_step()
        sync()
        ts = time.perf_counter()
        for _ in range(args.steps):
            n_seeds = seeds_step()
        sync()
        dts = time.perf_counter() - ts
        seeds_leg = {"value": world * n * args.steps / dts, "unit": "reads/s", "ms_per_step": dts / args.steps * 1e3, "seeds_per_step_per_gpu": n_seeds,
                     "d2h_bytes_per_step": n_seeds * 16 + (n + 1) * 8, "note": "host reads in -> mem_seed_t {rbeg,qbeg,len} out (min_seed_len 19, max_occ 10000)"}
        log("seeds leg:", seeds_leg)
        # ---- section 8f-3: ... -> seeds -> chains (+ mem_chain_flt) on the device, D2H of chains and their seeds
        chain_off = sg.PinnedArray(lib, (n + 1,), np.int64)
        pchains = sg.PinnedArray(lib, (4 * n, 3), np.int64)
        pcseeds = sg.PinnedArray(lib, (8 * n, 2), np.int64)
        copt = sg.ChainOpt(100, 10000, 19, 0.5, 0.5, 1)
        def chains_step():
            g.stage(pseq.array, poffs.array)
            g.run_collect(opt)
            tot, nc, ns = C.c_int64(0), C.c_int64(0), C.c_int64(0)
            rc = lib.smem_gpu_seeds(g.h, C.c_int(19), C.c_int64(10000), None, C.c_int64(0), seed_off.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(tot))
            assert rc in (0, -5), (rc, lib.smem_gpu_last_error(g.h))          # seeds stay in HBM: only their count comes back
            rc = lib.smem_gpu_chains(g.h, C.byref(copt), C.c_int64(int(ix.seq_len) // 2), C.c_void_p(pchains.array.ctypes.data), C.c_int64(pchains.array.shape[0]),
                                     C.c_void_p(pcseeds.array.ctypes.data), C.c_int64(pcseeds.array.shape[0]),
                                     chain_off.array.ctypes.data_as(C.POINTER(C.c_int64)), C.byref(nc), C.byref(ns))
            assert rc == 0, (rc, lib.smem_gpu_last_error(g.h))
            return int(nc.value), int(ns.value)
        for _ in range(2):
            chains_step()
        sync()
        ts = time.perf_counter()
        for _ in range(args.steps):
            n_chains, n_cseeds = chains_step()
        sync()
        dts = time.perf_counter() - ts
        chains_leg = {"value": world * n * args.steps / dts, "unit": "reads/s", "ms_per_step": dts / args.steps * 1e3, "chains_per_step_per_gpu": n_chains,
                      "chain_seeds_per_step_per_gpu": n_cseeds, "chain_kernels_ms": g.get_param("chain_kernels_us") / 1e3,
                      "d2h_bytes_per_step": n_chains * 24 + n_cseeds * 16 + (n + 1) * 16,
                      "note": "host reads in -> mem_chain_t after mem_chain_flt out (w 100, max_chain_gap 10000, mask_level 0.5, chain_drop_ratio 0.5)"}
        log("chains leg:", chains_leg)
        seeds_leg["chains"] = chains_leg

    # ---- optional: every interval of the step against the oracle (BASELINE config 2)
    if args.full_compare and rank == 0:
        from oracle.binding import Oracle
        fm = importlib.import_module("bwa-mem-harp2_b200.fmindex")
        ixh = fm.BwtIndex(ix.primary, ix.L2, ix.seq_len, ix.bwt_size, ix.words_numpy())
        t0c = time.time()
        want = Oracle(ixh).collect(seq, offs, OSeedOpt(), nthreads=ncores)
        same = (np.array_equal(want["read_off"], proff.array) and np.array_equal(want["intv"], pintv.array[:tot_e2e]))
        log(f"full compare: {n} reads, {len(want['intv'])} intervals, bit_exact={same} (oracle {time.time() - t0c:.1f}s)")
        parity = dict(parity or {}, full_compare_reads=n, full_compare_intervals=int(len(want["intv"])), full_compare_bit_exact=bool(same))
        if not same:
            raise SystemExit("PARITY FAILURE in --full-compare")

    # ---- optional: batch-size sweep (BASELINE config 5): one smem_gpu_collect call of B reads, host buffers
    batch_sweep = None
    if args.batch_sweep and rank == 0:
        batch_sweep = []
        for hot in (0, 16384):
            g.set_param("l2_hot_min_intv", hot)
            for B in (64, 256, 1024, 4096, 16384, 65536, 262144, 1048576):
                if B > n:
                    continue
                o = poffs.array[: B + 1]
                s_ = pseq.array[: int(o[-1])]
                def call():
                    tot = C.c_int64(0)
                    rc = lib.smem_gpu_collect(g.h, C.c_int64(B), s_.ctypes.data_as(C.POINTER(C.c_uint8)), o.ctypes.data_as(C.POINTER(C.c_int64)),
                                              C.byref(opt), pintv.array.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_int64(pintv.array.shape[0]),
                                              proff.array.ctypes.data_as(C.POINTER(C.c_int64)), None, C.byref(tot))
                    assert rc == 0, rc
                for _ in range(3):
                    call()
                reps = 20 if B <= 16384 else 5
                tb = time.perf_counter()
                for _ in range(reps):
                    call()
                ms = (time.perf_counter() - tb) / reps * 1e3
                batch_sweep.append({"reads_per_call": B, "l2_hot_min_intv": hot, "latency_ms": round(ms, 3), "reads_per_s": round(B / ms * 1e3)})
                log(f"batch sweep B={B} l2_hot_min_intv={hot}: {ms:.3f} ms/call -> {B / ms / 1e3:.2f} M reads/s")
        g.set_param("l2_hot_min_intv", 0)

    # ---- max over ranks
    times = torch.tensor([dt, dt_e2e, float(np.mean(seed_ms))], dtype=torch.float64, device=device)
    if use_dist:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dt, dt_e2e, seed_avg_ms = [float(x) for x in times.tolist()]

    probe = None
    if args.probe and rank == 0:
        probe = {}
        for bb in (64, 32):
            for chains in (512, 1024, 2048):
                probe[f"{bb}B_x{chains}"] = round(g.gather_roofline(bb, 0, chains, 1000), 1)
        log("random-access probe GB/s:", probe)
    rand64 = rand64_split = None
    e2e_overflow = g.timing()["overflow_reads"]
    if rank == 0:
        # the roofline of the access pattern: dependent 64 B gathers over the whole index, one coalesced request
        # per block (lane pair, 2 x 32 B) -- and, for reference, the same gathers issued as two per-lane requests
        g.set_param("probe_variant", 10); rand64 = g.gather_roofline(64, 0, 2048, 1000)
        g.set_param("probe_variant", 0); rand64_split = g.gather_roofline(64, 0, 2048, 1000)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic = None
        try:   # dram__bytes_read + dram__bytes_write of seed_kernel from the committed `ncu --set full` capture of this workload
            tr = json.load(open(os.path.join(ROOT, "profiles", "seed_traffic.json")))
            if not (uw_info and "bytes" in uw_info):
                tr = tr.get("without_text_index", {})          # the capture of the kernel without the unique-walk tables
            if tr.get("ref_bp") == args.ref_bp and tr.get("reads") == args.reads and tr.get("read_len") == args.read_len:
                traffic = tr["dram_bytes_per_launch"]
        except Exception:
            pass
        value = world * n * args.steps / dt
        achieved = n * bytes_per_read / (seed_avg_ms * 1e-3) / 1e9
        # the bound that actually holds for this access pattern: DRAM-missing requests per second.
        # requests of the kernel = ncu DRAM bytes / 64 (the sector pairs a lane pair gathers); ceiling = the coalesced 64 B probe
        request_rate = None
        if traffic and rand64 and not fast_info:
            rq = traffic / 64.0 / (seed_avg_ms * 1e-3) / 1e9
            request_rate = {"achieved_G_per_s": rq, "ceiling_G_per_s": rand64 / 64.0, "frac": rq / (rand64 / 64.0),
                            "note": "DRAM requests of 64 B per second: ncu dram bytes of the committed capture / 64 / live kernel time, against the coalesced gather probe measured in this run"}
        out = {
            "metric": "smem_seeded_101bp_reads_per_sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config,
            "e2e": {"value": world * n * args.steps / dt_e2e, "unit": "reads/s", "h2d_bytes_per_step": int(te["h2d_bytes"]),
                    "d2h_bytes_per_step": int(te["d2h_bytes"]), "ms_per_step": dt_e2e / args.steps * 1e3,
                    "pipeline_lanes_per_gpu": args.lanes, "host_threads": T, "intervals": int(tot_e2e)},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic if not fast_info else None, "peak_source": peak_src,
                         "kernel": "fast_kernel + resolve_kernel (k-mer count pyramid; FM re-run of escaped reads not included)" if fast_info else "seed_kernel<COLLECT>",
                         "kernel_ms": seed_avg_ms, "algorithmic_bytes_per_read": bytes_per_read, "executed": executed,
                         "achieved_note": "algorithmic bytes of the REFERENCE algorithm (SURVEY 8d) / kernel time; `executed` and `traffic` show what the kernel really touches",
                         "request_rate": request_rate,
                         "random_access_peak": rand64, "frac_of_random_access": achieved / rand64 if rand64 else None,
                         "random_access_peak_two_requests": rand64_split,
                         "random_access_note": "dependent 64 B gathers over the whole index (smem_gpu_gather_roofline): one coalesced "
                                               "request per block by a lane pair; L2 hits on hot blocks let the kernel exceed it"},
            "cpu_baseline": cpu_baseline, "parity": parity, "clocks": clocks,
            "intervals_per_step_per_gpu": int(total), "overflow_reads": int(overflow), "index_build_s": t_index,
            "blocks_per_sm": g.get_param("blocks_per_sm"), "l2_hot_min_intv": g.get_param("l2_hot_min_intv"),
            "repeat_filter": rf_info, "unique_walk_tables": uw_info, "cpu_binding": ("%d cpus local to the GPU (NVML): %d..%d" % (len(cpu_binding), cpu_binding[0], cpu_binding[-1])) if cpu_binding else None,
            "fast_path": dict(fast_info, escaped_reads_per_step=int(escaped), blocks_per_sm=g.get_param("fast_blocks_per_sm")) if fast_info else None,
            "device_ms_per_step": float(np.mean(dev_ms)),
        }
        if probe:
            out["random_access_probe_gbs"] = probe
        if batch_sweep:
            out["batch_sweep"] = batch_sweep
        if seeds_leg:
            out["seeds_api"] = seeds_leg
        emit(json.dumps(out))
    if use_dist:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
