"""Drop-in boundary, end to end: the reference's own `bwa mem` with ONLY bwt_smem1_batched replaced by the
GPU adapter (oracle/_ref/bwa_gpu, see oracle/Makefile `gpu`) must print the same SAM as the reference's CPU
path (`-b 1` never offloads).  The binaries are prebuilt in the build container and
travel with the snapshot; nothing here reads /root/reference."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pkg

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref", "bwa_ref")
GPU = os.path.join(ROOT, "oracle", "_ref", "bwa_gpu")


def sam_body(txt):
    return [l for l in txt.splitlines() if not l.startswith("@PG")]


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(GPU)), reason="oracle/_ref binaries were not built")
@pytest.mark.parametrize("paired,batch,cache", [(False, 64, 1), (True, 64, 1), (False, 1000, 1), (False, 64, 0), (True, 2048, 1)])
def test_bwa_mem_sam_identical(tmp_path, paired, batch, cache):
    sy = pkg("synth")
    ref = sy.make_reference(300_000, 42)
    refn = ref.numpy()
    refn[40_000:41_000] = refn[5_000:6_000]
    fa = str(tmp_path / "g.fa")
    sy.write_fasta(fa, [("chrA", refn[:180_000]), ("chrB", refn[180_000:])])
    subprocess.run([REF, "index", fa], check=True, capture_output=True, cwd=tmp_path)
    n = 3000
    reads = sy.simulate_reads(ref, n, 101, 0.02, seed=7, n_frac=0.05, paired=paired).numpy()
    fqs = []
    if paired:
        sy.write_fastq(str(tmp_path / "r1.fq"), reads[0::2]); sy.write_fastq(str(tmp_path / "r2.fq"), reads[1::2])
        fqs = [str(tmp_path / "r1.fq"), str(tmp_path / "r2.fq")]
    else:
        sy.write_fastq(str(tmp_path / "r.fq"), reads)
        fqs = [str(tmp_path / "r.fq")]
    cpu = subprocess.run([REF, "mem", "-t", "2", "-b", "1", fa] + fqs, check=True, capture_output=True, cwd=tmp_path, text=True)
    env = dict(os.environ, SMEM_GPU_MAX_READ_LEN="256", SMEM_GPU_MAX_BATCH="4096", SMEM_GPU_ADAPTER_CACHE=str(cache), SMEM_GPU_ADAPTER_STATS="1")
    gpu = subprocess.run([GPU, "mem", "-t", "2", "-b", str(batch), fa] + fqs, check=True, capture_output=True, cwd=tmp_path, text=True,
                         env=env, timeout=600)
    a, b = sam_body(cpu.stdout), sam_body(gpu.stdout)
    assert len(a) == len(b) and len(a) > n
    assert a == b
    stats = [l for l in gpu.stderr.splitlines() if "lists_from_cache" in l]
    assert stats, gpu.stderr[-2000:]
    kv = dict(t.split("=") for t in stats[-1].split()[1:])
    if cache:      # (nearly) every per-round call of the reference was served from the one-launch trace of its batch
        assert int(kv["lists_from_cache"]) > 3 * n and int(kv["gpu_calls"]) <= 2 * (n // batch + 2) + 8, kv
    else:
        assert int(kv["lists_from_cache"]) == 0 and int(kv["gpu_calls"]) > 4 * (n // batch), kv


HARP = os.path.join(ROOT, "oracle", "_ref", "bwa_harp")


@pytest.mark.skipif(not (os.path.exists(REF) and os.path.exists(HARP)), reason="oracle/_ref binaries were not built")
@pytest.mark.parametrize("threads", [1, 2])
def test_harp_wire_protocol_shim(tmp_path, threads):
    """Section 8f-4: every reference object unmodified and built WITHOUT -DUSE_SW (real handshakes, its own record
    packing and manager thread); only HelloALINLB.cpp is replaced by host/harp_shim.c, which plays the AFU on the GPU.
    101 bp reads only -- the protocol's own limit (bwt.c:575)."""
    sy = pkg("synth")
    ref = sy.make_reference(250_000, 43)
    fa = str(tmp_path / "g.fa")
    sy.write_fasta(fa, [("chrA", ref.numpy())])
    subprocess.run([REF, "index", fa], check=True, capture_output=True, cwd=tmp_path)
    n = 1500
    reads = sy.simulate_reads(ref, n, 101, 0.02, seed=9, n_frac=0.05).numpy()
    sy.write_fastq(str(tmp_path / "r.fq"), reads)
    cpu = subprocess.run([REF, "mem", "-t", str(threads), "-b", "1", fa, str(tmp_path / "r.fq")], check=True, capture_output=True,
                         cwd=tmp_path, text=True)
    env = dict(os.environ, HARP_SHIM_REF_MB="16")
    hw = subprocess.run([HARP, "mem", "-t", str(threads), "-b", "64", fa, str(tmp_path / "r.fq")], check=True, capture_output=True,
                        cwd=tmp_path, text=True, env=env, timeout=600)
    a, b = sam_body(cpu.stdout), sam_body(hw.stdout)
    assert len(a) == len(b) and a == b
    stat = [l for l in hw.stderr.splitlines() if l.startswith("[harp_shim] requests=")]
    assert stat, hw.stderr[-3000:]
    kv = dict(t.split("=") for t in stat[-1].split()[1:])
    assert int(kv["index_staged"]) == 1 and int(kv["requests"]) > 0
    if threads == 1:      # nobody competes for the single in-flight request: (nearly) every call goes through the protocol
        assert int(kv["reads"]) > 3 * n, kv
