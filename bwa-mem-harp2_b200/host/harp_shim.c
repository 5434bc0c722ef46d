/* harp_shim.c -- the HARPv2 workspace protocol served by the B200.
 *
 * Plain-C replacement of software/HelloALINLB.cpp (Intel AAL + MPF/VTP bring-up, 699 lines of C++) for the
 * LITERALLY UNCHANGED reference objects built WITHOUT -DUSE_SW: their own bwt_smem1_batched (bwt.c:444-774) packs
 * 256-byte records, their own harp_management thread (fastmap.c:320-429) copies them into SPL_BWT_input and flips
 * *handshake, their own bwa_idx_load_bwt (bwa.c:289-301) copies the index into SPL_BWT_ref and writes handshake 2.
 * This file owns the workspace those globals point into (HelloALINLB.cpp:59-72,397-412) and runs a service thread
 * that plays the AFU (hardware/afu_core.v):
 *
 *   handshake == 2      "BWT + cnt table are in place" (bwa.c:293)            -> note it, answer 8 (afu_core.v:1809-1815)
 *   handshake == 1 | 4  a batch of *read_size records is in SPL_BWT_input      -> one bwt_smem1 per record on the GPU
 *                       (fastmap.c:340-356; record layout bwt.c:574-596)          (smem_gpu_smem1), AFU-format records
 *                                                                                  into SPL_BWT_output (afu_core.v:1931-
 *                                                                                  1960, parsed at bwt.c:719-749), then 16
 *
 * The protocol's limits are kept because the unchanged host code bakes them in: 101-base queries (bwt.c:575),
 * <= 1023 reads per request, output record < 1024 words (fastmap.c:413).  The memory image carries neither bwt_size
 * nor seq_len (the AFU never needed them); they are recovered from the image itself: the array ends with the final
 * checkpoint counts (bwtindex.c:145-147), whose sum is seq_len, and bwt_size follows from seq_len.  primary and
 * L2[0..3] ride in every record (words 18, 28-31) exactly as the AFU received them.
 */
#include "../../include/smem_gpu.h"
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* the globals the unchanged bwa.c / fastmap.c extern (HelloALINLB.cpp:65-72) */
unsigned int *SPL_BWT_ref, *SPL_CNT_table, *read_size, *handshake;
unsigned long int *SPL_BWT_input, *SPL_BWT_output;

int top_main(int argc, char *argv[]);   /* top.c:63 */

#define MB(x) ((size_t)(x) << 20)
#define READ_LEN 101                     /* bwt.c:575, afu_core.v:4759 */
#define REC_WORDS 32                     /* 256-byte input record */

static size_t g_ref_bytes;
static smem_gpu_t *g_h;
static volatile int g_stop, g_index_staged, g_index_on_gpu;
static unsigned long long g_requests, g_reads, g_intervals;

static void die(const char *what, int rc)
{
	fprintf(stderr, "[harp_shim] %s: %s (%d) %s\n", what, smem_gpu_strerror(rc), rc, g_h ? smem_gpu_last_error(g_h) : "");
	abort();
}

/* bwt_size / seq_len from the image: the last 8 words are 4 x uint64 final counts whose sum S satisfies
 * bwt_size == ceil(S/16) + 8 * (ceil(S/128) + 1)   (bwtindex.c:133-134) */
static int locate_index_end(const uint32_t *w, size_t max_words, uint64_t *seq_len, uint64_t *bwt_size)
{
	size_t last = max_words, k;
	while (last > 0 && w[last - 1] == 0) --last;
	for (k = last; k <= last + 8 && k <= max_words; ++k) {
		uint64_t c[4], s;
		if (k < 16) continue;
		memcpy(c, w + k - 8, 32);
		s = c[0] + c[1] + c[2] + c[3];
		if (s > 0 && ((s + 15) >> 4) + 8 * ((s + 127) / 128 + 1) == k) { *seq_len = s; *bwt_size = k; return 0; }
	}
	return -1;
}

static void upload_index(const unsigned long int *rec)
{
	smem_index_desc_t ix;
	int rc;
	if (locate_index_end(SPL_BWT_ref, g_ref_bytes / 4, &ix.seq_len, &ix.bwt_size) != 0) {
		fprintf(stderr, "[harp_shim] cannot locate the end of the index image in the workspace\n");
		abort();
	}
	ix.primary = rec[18];
	ix.L2[0] = rec[28]; ix.L2[1] = rec[29]; ix.L2[2] = rec[30]; ix.L2[3] = rec[31]; ix.L2[4] = ix.seq_len;
	ix.bwt = SPL_BWT_ref;
	if ((rc = smem_gpu_create(&g_h, 1, 0, 1024, READ_LEN)) != 0) die("smem_gpu_create", rc);
	if ((rc = smem_gpu_upload_index(g_h, &ix)) != 0) die("smem_gpu_upload_index", rc);
	fprintf(stderr, "[harp_shim] index image: seq_len=%llu bwt_size=%llu primary=%llu -> HBM\n", (unsigned long long)ix.seq_len,
	        (unsigned long long)ix.bwt_size, (unsigned long long)ix.primary);
	g_index_on_gpu = 1;
}

static void serve_batch(void)
{
	static uint8_t seq[1024 * READ_LEN];
	static int64_t offs[1025], roff[1025];
	static int32_t x[1024], mi[1024], ret[1024];
	static smem_intv_t *out;
	static size_t out_cap;
	const unsigned n = *read_size & 0x3ff;                       /* 10 bits, afu_core.v:1212 */
	unsigned long int *o = SPL_BWT_output;
	int64_t total = 0;
	unsigned r;
	int rc;
	if (n == 0) return;
	if (!g_index_on_gpu) upload_index(SPL_BWT_input);
	for (r = 0; r < n; ++r) {
		const unsigned long int *rec = SPL_BWT_input + (size_t)r * REC_WORDS;
		memcpy(seq + (size_t)r * READ_LEN, rec, READ_LEN);       /* words 0-15: the query, one base per byte */
		offs[r] = (int64_t)r * READ_LEN;
		x[r] = (int32_t)(rec[16] & 0x7f);                        /* the AFU uses 7 bits of x and min_intv */
		mi[r] = (int32_t)(rec[17] & 0x7f);
	}
	offs[n] = (int64_t)n * READ_LEN;
	if (out_cap < (size_t)n * 128) { out_cap = (size_t)n * 128; out = (smem_intv_t *)realloc(out, out_cap * sizeof *out); }
	rc = smem_gpu_smem1(g_h, n, seq, offs, x, mi, out, (int64_t)out_cap, roff, ret, &total);
	if (rc != 0) die("smem_gpu_smem1", rc);
	for (r = 0; r < n; ++r) {                                    /* afu_core.v:1931-1960 / bwt.c:719-749 */
		const size_t cnt = (size_t)(roff[r + 1] - roff[r]);
		size_t k;
		if (8 + cnt * 4 + 4 >= 1024) { fprintf(stderr, "[harp_shim] %zu intervals do not fit the protocol's output record\n", cnt); abort(); }
		/* the output carve-out is the LAST megabyte of the workspace (HelloALINLB.cpp:61,410): a request whose records do not
		 * fit it cannot be answered through this protocol -- stop with a message rather than write past the allocation */
		if ((size_t)(o - SPL_BWT_output) + 8 + cnt * 4 + 4 > MB(1) / sizeof *o) {
			fprintf(stderr, "[harp_shim] the %u records of this request do not fit the protocol's 1 MB output region (read %u, %zu intervals)\n", n, r, cnt);
			abort();
		}
		memset(o, 0, 64);
		o[0] = r; o[1] = cnt; o[2] = (unsigned long int)ret[r];
		o += 8;
		for (k = 0; k < cnt; ++k, o += 4) memcpy(o, &out[roff[r] + (int64_t)k], 32);
		if (cnt & 1) { memset(o, 0, 32); o += 4; }
	}
	++g_requests; g_reads += n; g_intervals += (unsigned long long)total;
}

static void *service(void *arg)
{
	(void)arg;
	while (!g_stop) {
		const unsigned v = __atomic_load_n(handshake, __ATOMIC_ACQUIRE);
		if (v == 2) {                                            /* bwa.c:293: index + cnt table staged */
			g_index_staged = 1;
			__atomic_store_n(handshake, 8u, __ATOMIC_RELEASE);   /* afu_core.v:1809-1815 */
		} else if (v == 1 || v == 4) {                           /* fastmap.c:348-356 */
			serve_batch();
			__atomic_store_n(handshake, 16u, __ATOMIC_RELEASE);  /* afu_core.v:1845-1858 */
		} else usleep(20);
	}
	return 0;
}

int main(int argc, char *argv[])
{
	const char *e = getenv("HARP_SHIM_REF_MB");
	pthread_t th;
	char *ws;
	int rc;
	g_ref_bytes = MB(e ? atoi(e) : 3072);                        /* BWT_ref, HelloALINLB.cpp:60 */
	ws = (char *)calloc(g_ref_bytes + MB(3), 1);                 /* + CNT_table, BWT_input, BWT_output of 1 MB each */
	if (!ws) { fprintf(stderr, "[harp_shim] cannot allocate the %zu MB workspace\n", (g_ref_bytes >> 20) + 3); return 1; }
	SPL_BWT_ref = (unsigned int *)ws;                            /* carve-up of HelloALINLB.cpp:397-412 */
	SPL_CNT_table = (unsigned int *)(ws + g_ref_bytes);
	handshake = (unsigned int *)(ws + g_ref_bytes + MB(1)) - 1;
	read_size = (unsigned int *)(ws + g_ref_bytes + MB(1)) - 2;
	SPL_BWT_input = (unsigned long int *)(ws + g_ref_bytes + MB(1));
	SPL_BWT_output = (unsigned long int *)(ws + g_ref_bytes + MB(2));
	pthread_create(&th, 0, service, 0);
	rc = top_main(argc, argv);                                   /* HelloALINLB.cpp:460 */
	g_stop = 1;                                                  /* CSR_CTL = 7, HelloALINLB.cpp:463 */
	pthread_join(th, 0);
	fprintf(stderr, "[harp_shim] requests=%llu reads=%llu intervals=%llu index_staged=%d\n", g_requests, g_reads, g_intervals, g_index_staged);
	if (g_h) smem_gpu_destroy(g_h);
	free(ws);
	return rc;
}
