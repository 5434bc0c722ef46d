/* bwt_smem1_batched_gpu.c -- link-compatible replacement of the reference's offload wrapper.
 *
 * Replaces bwt_smem1_batched (software/bwt.c:444-774): where the reference packs 256-byte records into
 * worker_mem[tid], hands them to the harp_management thread (fastmap.c:320-429) and unpacks the AFU's
 * variable-length output -- or silently falls back to the CPU (bwt.c:651-717) -- this adapter sends the
 * same per-call work (one bwt_smem1 per not-done read) to the CUDA service through the C ABI of
 * include/smem_gpu.h.  bwamem.c, fastmap.c, bwa.c and kthread_batch.c stay unchanged above it.
 *
 * Per-call semantics kept (bwt.c:510-555, 719-749):
 *   pass 1 (is_middle == 0): x = ori_start[i], min_intv = max(start_width, 1); results -> itr[i]->matches,
 *                            itr[i]->start = ret;
 *   pass 2 (is_middle == 1): x = middle of itr[i]->matches->a[max_i[i]], min_intv = that.x[2] + 1;
 *                            results -> itr[i]->sub, start untouched;
 *   reads with done[i] != 0 are not touched; INIT / FREE only manage the reference's thread-local scratch,
 *   of which this adapter has none.
 * Differences by design: no 101 bp limit (bwt.c:575), no 128-read partitions (bwt.c:45), no CPU fallback --
 * a GPU failure aborts with a message, as the reference's own asserts do (fastmap.c:413).
 */
#include "../../include/bwa_abi.h"
#include "../../include/smem_gpu.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;   /* one handle, one caller at a time */
static smem_gpu_t *g_h = 0;
static const harp_bwt_t *g_bwt = 0;
static int64_t g_max_batch = 1 << 16;
static int g_max_len = 1024;
static uint64_t g_stats[4];   /* GPU calls, reads sent, intervals received, lists served from the batch cache */

/* staging buffers, grown on demand, protected by g_lock */
static uint8_t *g_seq; static size_t g_seq_cap;
static int64_t *g_offs, *g_roff; static int32_t *g_x, *g_mi, *g_ret, *g_idx; static size_t g_n_cap;
static smem_intv_t *g_out; static size_t g_out_cap;

static void print_stats(void);

static void die(const char *what, int rc)
{
	fprintf(stderr, "[bwt_smem1_batched/gpu] %s: %s (%d) %s\n", what, smem_gpu_strerror(rc), rc, g_h ? smem_gpu_last_error(g_h) : "");
	abort();
}

int harp_gpu_service_start(const harp_bwt_t *bwt)
{
	int devs[64], n_dev = 0, rc;
	const char *e;
	smem_index_desc_t ix;
	if (g_h && g_bwt == bwt) return 0;
	if (g_h) { smem_gpu_destroy(g_h); g_h = 0; }
	if ((e = getenv("SMEM_GPU_MAX_BATCH")) != 0) g_max_batch = atoll(e);
	if ((e = getenv("SMEM_GPU_MAX_READ_LEN")) != 0) g_max_len = atoi(e);
	e = getenv("SMEM_GPU_DEVICES");
	if (e && *e) {
		char *dup = strdup(e), *tok, *save = 0;
		for (tok = strtok_r(dup, ",", &save); tok && n_dev < 64; tok = strtok_r(0, ",", &save)) devs[n_dev++] = atoi(tok);
		free(dup);
	}
	if (n_dev == 0) devs[n_dev++] = 0;
	if ((rc = smem_gpu_create(&g_h, n_dev, devs, g_max_batch, g_max_len)) != 0) return rc;
	/* the upload step of bwa_idx_load_bwt (bwa.c:289-291), to HBM instead of the MPF workspace */
	ix.primary = bwt->primary;
	memcpy(ix.L2, bwt->L2, sizeof ix.L2);
	ix.seq_len = bwt->seq_len;
	ix.bwt_size = bwt->bwt_size;
	ix.bwt = bwt->bwt;
	if ((rc = smem_gpu_upload_index(g_h, &ix)) != 0) return rc;
	if (g_bwt == 0 && getenv("SMEM_GPU_ADAPTER_STATS")) atexit(print_stats);
	g_bwt = bwt;
	return 0;
}

void harp_gpu_service_stop(void)
{
	pthread_mutex_lock(&g_lock);
	if (g_h) smem_gpu_destroy(g_h);
	g_h = 0; g_bwt = 0;
	pthread_mutex_unlock(&g_lock);
}

void harp_gpu_service_stats(uint64_t out[4]) { memcpy(out, g_stats, sizeof g_stats); }

static void print_stats(void)
{
	fprintf(stderr, "[bwt_smem1_batched/gpu] gpu_calls=%llu reads_sent=%llu intervals=%llu lists_from_cache=%llu\n",
	        (unsigned long long)g_stats[0], (unsigned long long)g_stats[1], (unsigned long long)g_stats[2], (unsigned long long)g_stats[3]);
}

/* copy a list into a caller-owned kvec, growing it the way kv_push does (kvec.h:75-81) */
static void put_list(harp_bwtintv_v *v, const smem_intv_t *src, size_t cnt)
{
	if (cnt > v->m) {
		size_t m = v->m ? v->m : 2;
		while (m < cnt) m <<= 1;
		v->a = (harp_bwtintv_t *)realloc(v->a, m * sizeof(harp_bwtintv_t));
		v->m = m;
	}
	if (cnt) memcpy(v->a, src, cnt * sizeof(harp_bwtintv_t));
	v->n = cnt;
}

static void grow_n(size_t n)
{
	if (n <= g_n_cap) return;
	g_n_cap = n + n / 2 + 64;
	g_offs = (int64_t *)realloc(g_offs, (g_n_cap + 1) * sizeof(int64_t));
	g_roff = (int64_t *)realloc(g_roff, (g_n_cap + 1) * sizeof(int64_t));
	g_x = (int32_t *)realloc(g_x, g_n_cap * sizeof(int32_t));
	g_mi = (int32_t *)realloc(g_mi, g_n_cap * sizeof(int32_t));
	g_ret = (int32_t *)realloc(g_ret, g_n_cap * sizeof(int32_t));
	g_idx = (int32_t *)realloc(g_idx, g_n_cap * sizeof(int32_t));
}

/* ---- direct path: one bwt_smem1 per listed read, in one GPU call ------------------------------------------ */
/* idx[k] = position in the batch; x/mi per listed read.  Results go to matches (pass 1, + start) or sub (pass 2). */
static void direct_smem1(harp_smem_i **itr, const int32_t *idx, size_t n, const int32_t *x, const int32_t *mi, int is_middle)
{
	size_t k, nbytes = 0;
	int64_t total = 0;
	int rc;
	if (n == 0) return;
	grow_n(n);
	for (k = 0; k < n; ++k) { g_offs[k] = (int64_t)nbytes; nbytes += (size_t)itr[idx[k]]->len; }
	g_offs[n] = (int64_t)nbytes;
	if (nbytes > g_seq_cap) { g_seq_cap = nbytes + nbytes / 2 + 4096; g_seq = (uint8_t *)realloc(g_seq, g_seq_cap); }
	for (k = 0; k < n; ++k) memcpy(g_seq + g_offs[k], itr[idx[k]]->query, (size_t)itr[idx[k]]->len);
	if (g_out_cap < 32 * n) { g_out_cap = 32 * n + 1024; g_out = (smem_intv_t *)realloc(g_out, g_out_cap * sizeof(smem_intv_t)); }
	rc = smem_gpu_smem1(g_h, (int64_t)n, g_seq, g_offs, x, mi, g_out, (int64_t)g_out_cap, g_roff, g_ret, &total);
	if (rc == SMEM_GPU_E_CAPACITY && (size_t)total > g_out_cap) {   /* more intervals than guessed: size exactly and repeat */
		g_out_cap = (size_t)total + 1024;
		g_out = (smem_intv_t *)realloc(g_out, g_out_cap * sizeof(smem_intv_t));
		rc = smem_gpu_smem1(g_h, (int64_t)n, g_seq, g_offs, x, mi, g_out, (int64_t)g_out_cap, g_roff, g_ret, &total);
	}
	if (rc != 0) die("smem_gpu_smem1", rc);
	for (k = 0; k < n; ++k) {   /* scatter into the caller-owned kvecs (bwt.c:719-749) */
		harp_smem_i *it = itr[idx[k]];
		put_list(is_middle ? it->sub : it->matches, g_out + g_roff[k], (size_t)(g_roff[k + 1] - g_roff[k]));
		if (!is_middle) it->start = g_ret[k];
	}
	g_stats[0] += 1; g_stats[1] += n; g_stats[2] += (uint64_t)total;
}

/* ---- whole-batch cache --------------------------------------------------------------
 * The reference calls bwt_smem1_batched twice per iterator round (bwamem.c:172,204) for several rounds with a
 * shrinking active set (bwamem.c:390-393).  On the first pass-1 call of a batch (every active read still at its
 * first position) the adapter computes EVERY round of every read in one smem_gpu_trace launch and keeps the per-call
 * lists; the later calls are served from that cache after checking that the caller is where the trace expects it
 * (same smem_i, same query, same position).  Anything that does not line up (different -r/-s options than assumed,
 * a caller that is mid-way) falls back to the direct path for that read -- never to the CPU. */
typedef struct {
	int n;
	const harp_smem_i **itr; const uint8_t **query; int *len;
	int *next_pos;        /* where the next pass-1 call of the read must start (before skipping N) */
	int *step;            /* last step served by pass 1, -1 = none */
	int64_t *cur;         /* cursor into the read's trace entries */
	int64_t *roff;        /* [n+1] trace CSR, batch order (empty for reads that were done) */
	smem_intv_t *intv; uint16_t *tag, *ret; size_t cap;
	size_t n_cap;
	int32_t *list, *lx, *lmi; /* scratch for fallbacks */
} trace_cache_t;
static __thread trace_cache_t tc;
static int g_use_cache = -1, g_split_len = 28, g_split_width = 10;   /* -k 19 -r 1.5 -> (int)(19*1.5+.499), bwamem.c:456; split_width bwamem.c:60 */

static int skip_n(const uint8_t *q, int len, int pos) { while (pos < len && q[pos] > 3) ++pos; return pos; }

static void cache_grow(int n)
{
	if ((size_t)n <= tc.n_cap) return;
	tc.n_cap = (size_t)n + 64;
	tc.itr = (const harp_smem_i **)realloc((void *)tc.itr, tc.n_cap * sizeof(*tc.itr));
	tc.query = (const uint8_t **)realloc((void *)tc.query, tc.n_cap * sizeof(*tc.query));
	tc.len = (int *)realloc(tc.len, tc.n_cap * sizeof(int));
	tc.next_pos = (int *)realloc(tc.next_pos, tc.n_cap * sizeof(int));
	tc.step = (int *)realloc(tc.step, tc.n_cap * sizeof(int));
	tc.cur = (int64_t *)realloc(tc.cur, tc.n_cap * sizeof(int64_t));
	tc.roff = (int64_t *)realloc(tc.roff, (tc.n_cap + 1) * sizeof(int64_t));
	tc.list = (int32_t *)realloc(tc.list, tc.n_cap * sizeof(int32_t));
	tc.lx = (int32_t *)realloc(tc.lx, tc.n_cap * sizeof(int32_t));
	tc.lmi = (int32_t *)realloc(tc.lmi, tc.n_cap * sizeof(int32_t));
}

static void cache_free(void)
{
	free((void *)tc.itr); free((void *)tc.query); free(tc.len); free(tc.next_pos); free(tc.step); free(tc.cur); free(tc.roff);
	free(tc.intv); free(tc.tag); free(tc.ret); free(tc.list); free(tc.lx); free(tc.lmi);
	memset(&tc, 0, sizeof tc);
}

/* all rounds of the active reads of this batch in one launch; must be called with g_lock held */
static void cache_fill(harp_smem_i **itr, int batch_size, const int *done, int start_width)
{
	size_t n = 0, nbytes = 0, k;
	int i, rc;
	int64_t total = 0;
	smem_seed_opt_t opt;
	cache_grow(batch_size);
	grow_n((size_t)batch_size);
	for (i = 0; i < batch_size; ++i) {
		tc.itr[i] = itr[i]; tc.query[i] = done[i] ? 0 : itr[i]->query; tc.len[i] = done[i] ? 0 : itr[i]->len;
		tc.next_pos[i] = 0; tc.step[i] = -1;
		if (done[i]) continue;
		g_idx[n] = i; g_offs[n] = (int64_t)nbytes; nbytes += (size_t)itr[i]->len; ++n;
	}
	g_offs[n] = (int64_t)nbytes;
	if (nbytes > g_seq_cap) { g_seq_cap = nbytes + nbytes / 2 + 4096; g_seq = (uint8_t *)realloc(g_seq, g_seq_cap); }
	for (k = 0; k < n; ++k) memcpy(g_seq + g_offs[k], itr[g_idx[k]]->query, (size_t)itr[g_idx[k]]->len);
	opt.min_seed_len = g_split_len; opt.split_factor = 1.0; opt.split_width = g_split_width; opt.start_width = start_width;
	if (tc.cap < 48 * n + 1024) {
		tc.cap = 48 * n + 1024;
		tc.intv = (smem_intv_t *)realloc(tc.intv, tc.cap * sizeof(smem_intv_t));
		tc.tag = (uint16_t *)realloc(tc.tag, tc.cap * 2); tc.ret = (uint16_t *)realloc(tc.ret, tc.cap * 2);
	}
	rc = smem_gpu_trace(g_h, (int64_t)n, g_seq, g_offs, &opt, tc.intv, (int64_t)tc.cap, g_roff, tc.tag, tc.ret, &total);
	if (rc == SMEM_GPU_E_CAPACITY && (size_t)total > tc.cap) {
		tc.cap = (size_t)total + 1024;
		tc.intv = (smem_intv_t *)realloc(tc.intv, tc.cap * sizeof(smem_intv_t));
		tc.tag = (uint16_t *)realloc(tc.tag, tc.cap * 2); tc.ret = (uint16_t *)realloc(tc.ret, tc.cap * 2);
		rc = smem_gpu_trace(g_h, (int64_t)n, g_seq, g_offs, &opt, tc.intv, (int64_t)tc.cap, g_roff, tc.tag, tc.ret, &total);
	}
	if (rc != 0) die("smem_gpu_trace", rc);
	/* trace CSR is in gathered order; spread it to batch order (reads that were done own an empty range) */
	for (i = 0, k = 0; i < batch_size; ++i) {
		tc.roff[i] = g_roff[k]; tc.cur[i] = g_roff[k];
		if (!done[i]) ++k;
	}
	tc.roff[batch_size] = total;
	tc.n = batch_size;
	g_stats[0] += 1; g_stats[1] += n; g_stats[2] += (uint64_t)total;
}

void bwt_smem1_batched(harp_smem_i **itr, int *ori_start, int *max_i, int start_width, int is_middle,
                       int batch_size, const int *done, int bwt_batched_status)
{
	int i, rc;
	size_t n = 0, nf = 0;
	if (bwt_batched_status == HARP_BWT_BATCHED_FREE) { cache_free(); return; }
	if (bwt_batched_status != HARP_BWT_BATCHED_DO) return;   /* INIT: nothing thread-local to set up */
	if (batch_size <= 0) return;

	pthread_mutex_lock(&g_lock);
	if (g_use_cache < 0) {
		const char *e = getenv("SMEM_GPU_ADAPTER_CACHE");
		g_use_cache = e ? atoi(e) != 0 : 1;
		if ((e = getenv("SMEM_GPU_SPLIT_LEN")) != 0) g_split_len = atoi(e);
		if ((e = getenv("SMEM_GPU_SPLIT_WIDTH")) != 0) g_split_width = atoi(e);
	}
	for (i = 0; i < batch_size; ++i)
		if (!done[i] && (g_h == 0 || g_bwt != itr[i]->bwt)) { if ((rc = harp_gpu_service_start(itr[i]->bwt)) != 0) die("service start", rc); break; }
	cache_grow(batch_size);

	if (!is_middle) {
		/* pass 1 (bwt.c:514-546): x = ori_start[i], min_intv = max(start_width, 1) */
		int fresh = 0, active = 0;
		for (i = 0; i < batch_size; ++i) {
			if (done[i]) continue;
			++active;
			fresh += ori_start[i] == skip_n(itr[i]->query, itr[i]->len, 0);
		}
		if (g_use_cache && active > 0 && fresh == active) cache_fill(itr, batch_size, done, start_width);   /* first round of a batch */
		for (i = 0; i < batch_size; ++i) {
			harp_smem_i *it;
			int ok = 0;
			if (done[i]) continue;
			it = itr[i];
			if (g_use_cache && tc.n == batch_size && tc.itr[i] == it && tc.query[i] == it->query && tc.len[i] == it->len &&
			    skip_n(it->query, it->len, tc.next_pos[i]) == ori_start[i]) {
				const int s = tc.step[i] + 1;
				const int64_t end = tc.roff[i + 1];
				int64_t c = tc.cur[i], c0;
				while (c < end && tc.tag[c] < 2 * s) ++c;        /* skip an unused pass-2 list of the previous step */
				c0 = c;
				while (c < end && tc.tag[c] == 2 * s) ++c;
				if (c > c0) {
					put_list(it->matches, tc.intv + c0, (size_t)(c - c0));
					it->start = tc.ret[c0];                           /* bwt.c:731 */
					tc.cur[i] = c; tc.step[i] = s; tc.next_pos[i] = tc.ret[c0];
					ok = 1; ++nf;
				}
			}
			if (!ok) {                                               /* not what the trace expects: direct call for this read */
				tc.itr[i] = 0;
				tc.list[n] = i; tc.lx[n] = ori_start[i]; tc.lmi[n] = start_width; ++n;
			}
		}
		direct_smem1(itr, tc.list, n, tc.lx, tc.lmi, 0);
	} else {
		/* pass 2 (bwt.c:518-525,541-546): x = middle of matches[max_i], min_intv = its x[2] + 1, results -> sub */
		for (i = 0; i < batch_size; ++i) {
			harp_smem_i *it;
			const harp_bwtintv_t *p;
			int ok = 0;
			if (done[i]) continue;
			it = itr[i];
			p = &it->matches->a[max_i[i]];
			if (g_use_cache && tc.n == batch_size && tc.itr[i] == it && tc.query[i] == it->query && tc.step[i] >= 0) {
				const int64_t end = tc.roff[i + 1];
				int64_t c = tc.cur[i], c0 = c;
				while (c < end && tc.tag[c] == 2 * tc.step[i] + 1) ++c;
				if (c > c0) { put_list(it->sub, tc.intv + c0, (size_t)(c - c0)); tc.cur[i] = c; ok = 1; ++nf; }
			}
			if (!ok) {
				tc.list[n] = i;
				tc.lx[n] = (int)(((uint32_t)p->info + (uint32_t)(p->info >> 32)) >> 1);
				tc.lmi[n] = (int)(p->x[2] + 1);
				++n;
			}
		}
		direct_smem1(itr, tc.list, n, tc.lx, tc.lmi, 1);
	}
	g_stats[3] += nf;
	pthread_mutex_unlock(&g_lock);
}
