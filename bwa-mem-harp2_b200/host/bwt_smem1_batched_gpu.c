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
static uint64_t g_stats[3];

/* staging buffers, grown on demand, protected by g_lock */
static uint8_t *g_seq; static size_t g_seq_cap;
static int64_t *g_offs, *g_roff; static int32_t *g_x, *g_mi, *g_ret, *g_idx; static size_t g_n_cap;
static smem_intv_t *g_out; static size_t g_out_cap;

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

void harp_gpu_service_stats(uint64_t out[3]) { memcpy(out, g_stats, sizeof g_stats); }

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

void bwt_smem1_batched(harp_smem_i **itr, int *ori_start, int *max_i, int start_width, int is_middle,
                       int batch_size, const int *done, int bwt_batched_status)
{
	int i, rc;
	size_t n = 0, nbytes = 0, k;
	int64_t total = 0;
	if (bwt_batched_status != HARP_BWT_BATCHED_DO) return;   /* INIT/FREE: nothing thread-local to manage */
	if (batch_size <= 0) return;

	pthread_mutex_lock(&g_lock);
	/* gather the not-done reads of this call (bwt.c:510-555) */
	grow_n((size_t)batch_size);
	for (i = 0; i < batch_size; ++i) {
		const harp_smem_i *it;
		if (done[i]) continue;
		it = itr[i];
		if (g_h == 0 || g_bwt != it->bwt) { if ((rc = harp_gpu_service_start(it->bwt)) != 0) die("service start", rc); }
		g_idx[n] = i;
		if (!is_middle) {
			g_x[n] = ori_start[i];
			g_mi[n] = start_width;
		} else {
			const harp_bwtintv_t *p = &it->matches->a[max_i[i]];
			g_x[n] = (int)(((uint32_t)p->info + (uint32_t)(p->info >> 32)) >> 1);
			g_mi[n] = (int)(p->x[2] + 1);
		}
		g_offs[n] = (int64_t)nbytes;
		nbytes += (size_t)it->len;
		++n;
	}
	if (n == 0) { pthread_mutex_unlock(&g_lock); return; }
	g_offs[n] = (int64_t)nbytes;
	if (nbytes > g_seq_cap) { g_seq_cap = nbytes + nbytes / 2 + 4096; g_seq = (uint8_t *)realloc(g_seq, g_seq_cap); }
	for (k = 0; k < n; ++k) memcpy(g_seq + g_offs[k], itr[g_idx[k]]->query, (size_t)itr[g_idx[k]]->len);
	if (g_out_cap < 32 * n) { g_out_cap = 32 * n + 1024; g_out = (smem_intv_t *)realloc(g_out, g_out_cap * sizeof(smem_intv_t)); }

	rc = smem_gpu_smem1(g_h, (int64_t)n, g_seq, g_offs, g_x, g_mi, g_out, (int64_t)g_out_cap, g_roff, g_ret, &total);
	if (rc == SMEM_GPU_E_CAPACITY && (size_t)total > g_out_cap) {   /* more intervals than guessed: size exactly and repeat */
		g_out_cap = (size_t)total + 1024;
		g_out = (smem_intv_t *)realloc(g_out, g_out_cap * sizeof(smem_intv_t));
		rc = smem_gpu_smem1(g_h, (int64_t)n, g_seq, g_offs, g_x, g_mi, g_out, (int64_t)g_out_cap, g_roff, g_ret, &total);
	}
	if (rc != 0) die("smem_gpu_smem1", rc);

	/* scatter into the caller-owned kvecs (bwt.c:719-749) */
	for (k = 0; k < n; ++k) {
		harp_smem_i *it = itr[g_idx[k]];
		harp_bwtintv_v *v = is_middle ? it->sub : it->matches;
		const size_t cnt = (size_t)(g_roff[k + 1] - g_roff[k]);
		if (cnt > v->m) {
			size_t m = v->m ? v->m : 2;
			while (m < cnt) m <<= 1;
			v->a = (harp_bwtintv_t *)realloc(v->a, m * sizeof(harp_bwtintv_t));
			v->m = m;
		}
		if (cnt) memcpy(v->a, g_out + g_roff[k], cnt * sizeof(harp_bwtintv_t));
		v->n = cnt;
		if (!is_middle) it->start = g_ret[k];
	}
	g_stats[0] += 1; g_stats[1] += n; g_stats[2] += (uint64_t)total;
	pthread_mutex_unlock(&g_lock);
}
