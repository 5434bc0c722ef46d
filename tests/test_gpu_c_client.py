"""A plain-C host program written against include/smem_gpu.h alone -- the way a maintainer of the reference would call the bulk API
(INTEGRATION.md section 3): create, upload a bwt_t, pack bwa's byte reads into the 2-bit form, smem_gpu_collect_packed12 /
_packed11 / _packed, decode with the header's inline decoders, and smem_gpu_collect with bwa's own formats.  Compiled with gcc,
linked against libsmem_gpu.so, run as its own process; its output must be the oracle's interval lists, bit for bit."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, pkg
from oracle.binding import Oracle, SeedOpt as OSeedOpt

pytestmark = pytest.mark.gpu

C_SRC = r'''
#include "smem_gpu.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#define CHECK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s -> %d %s (%s)\n", #x, rc_, smem_gpu_strerror(rc_), h ? smem_gpu_last_error(h) : ""); return 2; } } while (0)
static void *slurp(const char *path, size_t *n) { FILE *f = fopen(path, "rb"); if (!f) return 0; fseek(f, 0, SEEK_END); *n = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
	void *p = malloc(*n + 64); if (fread(p, 1, *n, f) != *n) return 0; fclose(f); return p; }
int main(int argc, char **argv)
{
	smem_gpu_t *h = 0;
	size_t nb;
	uint64_t *hdr = (uint64_t *)slurp(argv[1], &nb);              /* primary, L2[5], seq_len, bwt_size, then the bwt words */
	uint8_t *seq = (uint8_t *)slurp(argv[2], &nb);
	int64_t *offs = (int64_t *)slurp(argv[3], &nb);
	const int64_t n = (int64_t)(nb / 8) - 1;
	const int max_len = atoi(argv[5]);
	smem_index_desc_t ix; ix.primary = hdr[0]; memcpy(ix.L2, hdr + 1, 40); ix.seq_len = hdr[6]; ix.bwt_size = hdr[7]; ix.bwt = (const uint32_t *)(hdr + 8);
	smem_seed_opt_t opt = { 19, 1.5, 10, 1 };
	CHECK(smem_gpu_create(&h, 1, 0, n, max_len));
	CHECK(smem_gpu_upload_index(h, &ix));
	/* bwa's byte reads -> the compact form, in pinned memory */
	smem_reads2_t r; memset(&r, 0, sizeof r);
	r.n_reads = n; r.stride = (max_len + 3) / 4;
	uint8_t *seq2; uint16_t *lens; smem_amb_t *amb; int64_t amb_cap = 1 << 20;
	CHECK(smem_gpu_host_alloc((void **)&seq2, (size_t)n * r.stride + 64)); CHECK(smem_gpu_host_alloc((void **)&lens, (size_t)n * 2 + 64));
	CHECK(smem_gpu_host_alloc((void **)&amb, (size_t)amb_cap * sizeof *amb));
	CHECK(smem_gpu_pack_reads(n, seq, offs, r.stride, seq2, lens, amb, amb_cap, &r.n_amb, 4));
	r.seq2 = seq2; r.lens = lens; r.amb = amb;
	int64_t cap = 64 * n + 1024, exc_cap = 16, total = 0, n_exc = 0; int32_t pos_bits = 0;
	uint32_t *roff = (uint32_t *)malloc((size_t)(n + 1) * 4);
	smem_intv12_t *rec12 = (smem_intv12_t *)malloc((size_t)cap * 12); smem_intv11_t *rec11 = (smem_intv11_t *)malloc((size_t)cap * 11);
	smem_intv16_t *rec16 = (smem_intv16_t *)malloc((size_t)cap * 16);
	smem_x2exc_t *exc = (smem_x2exc_t *)malloc((size_t)exc_cap * sizeof *exc);
	smem_intv_t *a = (smem_intv_t *)malloc((size_t)cap * sizeof *a), *b = (smem_intv_t *)malloc((size_t)cap * sizeof *b);
	int64_t *roff64 = (int64_t *)malloc((size_t)(n + 1) * 8);
	/* 12-byte records: the exception list is sized on demand (capacity protocol) */
	int rc = smem_gpu_collect_packed12(h, &r, &opt, rec12, cap, roff, exc, exc_cap, &n_exc, &pos_bits, &total);
	if (rc == SMEM_GPU_E_CAPACITY) {
		exc_cap = n_exc + 16; exc = (smem_x2exc_t *)realloc(exc, (size_t)exc_cap * sizeof *exc);
		rc = smem_gpu_fetch_packed12(h, rec12, cap, roff, exc, exc_cap, &n_exc, &pos_bits, &total);
	}
	CHECK(rc);
	for (int64_t k = 0; k < total; ++k) smem_intv12_unpack(&rec12[k], pos_bits, &a[k]);
	for (int64_t k = 0; k < n_exc; ++k) a[exc[k].index].x[2] = (uint64_t)exc[k].x2_lo | (uint64_t)exc[k].x2_hi << 32;
	/* 11-byte and 16-byte records, and bwa's own formats: the same lists */
	int64_t t2 = 0, ne2 = 0; int32_t pb2 = 0;
	if (max_len <= 512) {
		exc_cap = 4 * total + 16; exc = (smem_x2exc_t *)realloc(exc, (size_t)exc_cap * sizeof *exc);
		CHECK(smem_gpu_collect_packed11(h, &r, &opt, rec11, cap, roff, exc, exc_cap, &ne2, &pb2, &t2));
		if (t2 != total) { fprintf(stderr, "totals differ\n"); return 3; }
		for (int64_t k = 0; k < t2; ++k) smem_intv11_unpack(&rec11[k], pb2, &b[k]);
		for (int64_t k = 0; k < ne2; ++k) b[exc[k].index].x[2] = (uint64_t)exc[k].x2_lo | (uint64_t)exc[k].x2_hi << 32;
		if (memcmp(a, b, (size_t)total * sizeof *a)) { fprintf(stderr, "11-byte records differ\n"); return 3; }
	}
	CHECK(smem_gpu_collect_packed(h, &r, &opt, rec16, cap, roff, &t2));
	for (int64_t k = 0; k < t2; ++k) smem_intv16_unpack(&rec16[k], &b[k]);
	if (t2 != total || memcmp(a, b, (size_t)total * sizeof *a)) { fprintf(stderr, "16-byte records differ\n"); return 3; }
	CHECK(smem_gpu_collect(h, n, seq, offs, &opt, b, cap, roff64, 0, &t2));
	if (t2 != total || memcmp(a, b, (size_t)total * sizeof *a)) { fprintf(stderr, "bwtintv_t rows differ\n"); return 3; }
	for (int64_t i = 0; i <= n; ++i) if ((int64_t)roff[i] != roff64[i]) { fprintf(stderr, "offsets differ\n"); return 3; }
	FILE *f = fopen(argv[4], "wb");
	fwrite(&total, 8, 1, f); fwrite(roff64, 8, (size_t)(n + 1), f); fwrite(a, sizeof *a, (size_t)total, f); fclose(f);
	fprintf(stderr, "c client: %lld reads, %lld intervals, %lld exceptions (12-byte), %lld (11-byte), pos_bits %d\n", (long long)n, (long long)total, (long long)n_exc, (long long)ne2, pos_bits);
	smem_gpu_host_free(seq2); smem_gpu_host_free(lens); smem_gpu_host_free(amb);
	CHECK(smem_gpu_destroy(h));
	return 0;
}
'''


def test_c_program_against_the_header(tmp_path, fm, synth):
    sg = pkg("smem_gpu")
    sg.load_library()
    ref = synth.make_reference(400_000, 61)
    ix = fm.build_index(ref)
    refn = ref.numpy()
    reads = list(synth.simulate_reads(ref, 3000, 101, 0.02, seed=7, n_frac=0.05)) + list(synth.simulate_reads(ref, 500, 130, 0.03, seed=8))
    reads += [refn[k * 700:k * 700 + 1 + k % 5].copy() for k in range(40)]            # very short reads: sizes beyond the records' fields
    seq, offs = synth.to_batch(reads)
    want = Oracle(ix).collect(seq, offs, OSeedOpt(), nthreads=4)
    hdr = np.array([ix.primary, *[int(v) for v in ix.L2], ix.seq_len, ix.bwt_size], np.uint64)
    with open(tmp_path / "index.bin", "wb") as f:
        f.write(hdr.tobytes()); f.write(np.ascontiguousarray(ix.words_numpy(), np.uint32).tobytes())
    np.ascontiguousarray(seq, np.uint8).tofile(tmp_path / "seq.bin")
    np.ascontiguousarray(offs, np.int64).tofile(tmp_path / "offs.bin")
    (tmp_path / "client.c").write_text(C_SRC)
    libdir = os.path.join(ROOT, "bwa-mem-harp2_b200")
    subprocess.run(["gcc", "-O1", "-Wall", "-std=c99", "-I", os.path.join(ROOT, "include"), str(tmp_path / "client.c"), "-o", str(tmp_path / "client"),
                    "-L", libdir, "-lsmem_gpu", "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(tmp_path / "client"), str(tmp_path / "index.bin"), str(tmp_path / "seq.bin"), str(tmp_path / "offs.bin"),
                        str(tmp_path / "out.bin"), "160"], capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    raw = np.fromfile(tmp_path / "out.bin", np.uint64)
    total = int(raw[0]); n = len(offs) - 1
    read_off = raw[1:n + 2].astype(np.int64)
    intv = raw[n + 2:].reshape(total, 4)
    assert np.array_equal(read_off, want["read_off"]) and np.array_equal(intv, want["intv"])
    assert "exceptions" in p.stderr
