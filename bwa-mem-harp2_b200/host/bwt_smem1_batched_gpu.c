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
 *   reads with done[i] != 0 are not touched; INIT / FREE only manage thread-local scratch.
 *
 * How the reference's `-t N -b B` calling pattern reaches the GPU:
 *   * whole-batch cache: the first pass-1 call of a batch computes EVERY round of every read in one smem_gpu_trace launch;
 *     the reference's later per-round calls (two per round, shrinking active set, bwamem.c:390-393) are served from the
 *     calling thread's own cache without touching any lock;
 *   * caller aggregation (what harp_management arbitrates in the reference, fastmap.c:335-420): GPU work goes through a
 *     combining queue.  A caller posts its request; whoever finds a free service handle becomes the leader, takes every
 *     compatible pending request (up to the handle's batch limit), runs them as ONE launch, hands the slices back and wakes
 *     their owners.  With -b 64 and 16 workers a launch carries the reads of all the workers that arrived while the
 *     previous launches were in flight; SMEM_GPU_ADAPTER_HANDLES (default 4) launches run concurrently on their own streams;
 *   * staging in pinned memory (smem_gpu_host_alloc), one set per service handle, grown geometrically;
 *   * handles grow in place (smem_gpu_resize) when a longer read or a larger batch arrives -- no fixed 101 bp / 128-read
 *     limits (bwt.c:575, bwt.c:45);
 *   * no CPU fallback and no abort(): a GPU failure is reported once on stderr, recorded (harp_gpu_service_last_error) and
 *     the affected lists are returned empty; SMEM_GPU_ADAPTER_ABORT=1 restores fail-fast behaviour for batch jobs.
 */
#include "../../include/bwa_abi.h"
#include "../../include/smem_gpu.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define MAX_HANDLES 16

typedef struct {
	smem_gpu_t *h;
	int dev, busy;
	int64_t max_batch; int max_len;
	/* pinned staging, used by the leader that holds the handle */
	uint8_t *seq; size_t seq_cap;
	int64_t *offs, *roff; int32_t *x, *mi, *ret; size_t n_cap;
	smem_intv_t *out; uint16_t *tag, *ret16; size_t out_cap;
} svc_handle_t;

/* one caller's GPU work: n reads of its batch, either one raw bwt_smem1 each (kind 0) or the whole iterator walk (kind 1) */
typedef struct req {
	int kind, start_width;
	size_t n;
	harp_smem_i **itr; const int32_t *idx;      /* read k of the request = itr[idx[k]] */
	const int32_t *x, *mi;                      /* kind 0 */
	/* results, written by the leader into buffers the requester owns (the requester sleeps meanwhile) */
	smem_intv_t **intv; uint16_t **tag, **ret16; size_t *cap;   /* grown by the leader */
	int64_t *roff;                              /* [n + 1] */
	int32_t *ret;                               /* kind 0: [n] */
	int done, rc;
	struct req *next;
} req_t;

static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_cond_t g_cv = PTHREAD_COND_INITIALIZER;
static svc_handle_t g_hd[MAX_HANDLES];
static int g_n_handles = 0;      /* handles in service (index uploaded) */
static int g_n_created = 0;      /* handles created (by the preload thread or the first call) */
static req_t *g_head = 0, *g_tail = 0;
static const harp_bwt_t *g_bwt = 0;
static int64_t g_max_batch = 4096;
static int g_init_len = 160;
static int g_last_rc = 0, g_abort_on_error = 0, g_reported = 0, g_start_failed = 0;
static char g_last_err[512];
/* GPU calls, reads sent, intervals received, lists served from the batch cache, requests combined into those calls,
 * nanoseconds inside the GPU calls, nanoseconds inside bwt_smem1_batched (summed over threads), of which bringing the
 * service up (index upload, handles) or waiting for the thread that does; [8..11] leader breakdown: gather, launch, hand-back
 * nanoseconds and handle / buffer growth events */
static uint64_t g_stats[12];
static int g_use_cache = -1, g_split_len = 28, g_split_width = 10;   /* -k 19 -r 1.5 -> (int)(19*1.5+.499), bwamem.c:456; split_width bwamem.c:60 */

static uint64_t now_ns(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec; }
static void stat_add(int k, uint64_t v) { __atomic_fetch_add(&g_stats[k], v, __ATOMIC_RELAXED); }

static void print_stats(void)
{
	fprintf(stderr, "[bwt_smem1_batched/gpu] gpu_calls=%llu reads_sent=%llu intervals=%llu lists_from_cache=%llu requests=%llu gpu_s=%.3f adapter_s=%.3f startup_s=%.3f "
	        "gather_s=%.3f launch_s=%.3f handback_s=%.3f grow_events=%llu handles=%d\n",
	        (unsigned long long)g_stats[0], (unsigned long long)g_stats[1], (unsigned long long)g_stats[2], (unsigned long long)g_stats[3],
	        (unsigned long long)g_stats[4], (double)g_stats[5] * 1e-9, (double)(g_stats[6] - g_stats[7]) * 1e-9, (double)g_stats[7] * 1e-9,
	        (double)g_stats[8] * 1e-9, (double)g_stats[9] * 1e-9, (double)g_stats[10] * 1e-9, (unsigned long long)g_stats[11], g_n_handles);
}

/* record a failure: message once, code kept for harp_gpu_service_last_error */
static void fail(const char *what, int rc, smem_gpu_t *h)
{
	pthread_mutex_lock(&g_mu);
	g_last_rc = rc;
	snprintf(g_last_err, sizeof g_last_err, "%s: %s (%d) %s", what, smem_gpu_strerror(rc), rc, h ? smem_gpu_last_error(h) : "");
	if (!g_reported) { fprintf(stderr, "[bwt_smem1_batched/gpu] %s -- the affected interval lists are returned EMPTY (no CPU fallback)\n", g_last_err); g_reported = 1; }
	pthread_mutex_unlock(&g_mu);
	if (g_abort_on_error) abort();
}

int harp_gpu_service_last_error(const char **msg) { if (msg) *msg = g_last_err; return g_last_rc; }

static int stage_grow(svc_handle_t *s, size_t n, size_t nbytes);
static int out_grow(svc_handle_t *s, size_t need);

/* ---- service bring-up ------------------------------------------------------------------------------------------------
 * Two parts.  (1) Everything that does not need the index -- CUDA context, kernels (CUDA_MODULE_LOADING=EAGER), the service
 * handles with their device buffers, the pinned staging sized for a full batch -- is created once, by a background thread
 * started when the host program is loaded with `mem` on its command line, i.e. WHILE bwa reads its index files
 * (SMEM_GPU_ADAPTER_PRELOAD=0 turns that off; the first call then does it).  First-use allocation inside the worker threads
 * was measured at 50-750 ms per handle (pinned allocations and cudaFree serialise on the driver while 16 threads wait).
 * (2) The index upload of bwa_idx_load_bwt (bwa.c:289-291), to HBM instead of the MPF workspace, on the first DO call -- the
 * reference gives the adapter no earlier hook -- followed by one tiny launch per handle so that no worker pays a first launch. */
static int g_pre_state = 0;       /* 0 = not started, 1 = running, 2 = done */
static int g_pre_rc = 0;

#define FIRST_BATCH 4096          /* reads per launch the handles are created for; re-sized from the first batch the reference sends */

static int create_handles(void)
{
	int devs[64], n_dev = 0, rc, k, want = MAX_HANDLES;
	const char *e;
	g_max_batch = FIRST_BATCH;
	if ((e = getenv("SMEM_GPU_MAX_READ_LEN")) != 0) g_init_len = atoi(e);       /* first size only: handles grow on demand */
	if ((e = getenv("SMEM_GPU_ADAPTER_HANDLES")) != 0) want = atoi(e);
	if ((e = getenv("SMEM_GPU_ADAPTER_ABORT")) != 0) g_abort_on_error = atoi(e) != 0;
	if (g_max_batch < 64) g_max_batch = 64;
	if (g_init_len < 32) g_init_len = 32;
	e = getenv("SMEM_GPU_DEVICES");
	if (e && *e) {
		char *dup = strdup(e), *tok, *save = 0;
		for (tok = strtok_r(dup, ",", &save); tok && n_dev < 64; tok = strtok_r(0, ",", &save)) devs[n_dev++] = atoi(tok);
		free(dup);
	}
	if (n_dev == 0) devs[n_dev++] = 0;
	if (want < n_dev) want = n_dev;
	if (want > MAX_HANDLES) want = MAX_HANDLES;
	for (k = 0; k < want; ++k) {            /* handle k lives on device k mod n_dev */
		svc_handle_t *s = &g_hd[k];
		s->dev = devs[k % n_dev]; s->max_batch = g_max_batch; s->max_len = g_init_len;
		if ((rc = smem_gpu_create(&s->h, 1, &s->dev, s->max_batch, s->max_len)) != 0) return rc;
		++g_n_created;
		smem_gpu_set_param(s->h, "turn_min_reads", 1ll << 40);   /* concurrent small launches: never serialise on the kernel turn */
		if ((rc = stage_grow(s, (size_t)s->max_batch, (size_t)s->max_batch * 110)) != 0) return rc;
		if ((rc = out_grow(s, 24 * (size_t)s->max_batch + 4096)) != 0) return rc;
	}
	return 0;
}

static void *preload_main(void *arg)
{
	int rc;
	(void)arg;
	rc = create_handles();
	pthread_mutex_lock(&g_mu);
	g_pre_rc = rc; g_pre_state = 2;
	pthread_cond_broadcast(&g_cv);
	pthread_mutex_unlock(&g_mu);
	return 0;
}

__attribute__((constructor)) static void adapter_preload(int argc, char **argv)
{
	const char *e = getenv("SMEM_GPU_ADAPTER_PRELOAD");
	pthread_t th;
	int i, is_mem = 0;
	setenv("CUDA_MODULE_LOADING", "EAGER", 0);          /* load every kernel with the context, not at a worker's first launch */
	for (i = 1; argv && i < argc && i < 3; ++i) if (argv[i] && strcmp(argv[i], "mem") == 0) is_mem = 1;
	if (e ? atoi(e) == 0 : !is_mem) return;
	g_pre_state = 1;
	if (pthread_create(&th, 0, preload_main, 0) != 0) { g_pre_state = 0; return; }
	pthread_detach(th);
}

static int service_start_inner(const harp_bwt_t *bwt, int first_batch);
static int service_start_locked(const harp_bwt_t *bwt, int first_batch)
{
	int k;
	const int rc = service_start_inner(bwt, first_batch);
	if (rc) {            /* leave nothing half-built behind, and do not try again on every call */
		for (k = 0; k < MAX_HANDLES; ++k) if (g_hd[k].h) smem_gpu_destroy(g_hd[k].h);
		memset(g_hd, 0, sizeof g_hd);
		g_n_handles = 0; g_n_created = 0; g_start_failed = 1;
	}
	return rc;
}

static int service_start_inner(const harp_bwt_t *bwt, int first_batch)
{
	int rc, k;
	smem_index_desc_t ix;
	while (g_pre_state == 1) pthread_cond_wait(&g_cv, &g_mu);       /* the background part, if it is still at it (this releases g_mu: */
	if (g_n_handles && g_bwt == bwt) return 0;                      /*  another worker may have brought the service up meanwhile) */
	if (g_pre_state == 2 && g_pre_rc) return g_pre_rc;
	if (g_n_created == 0 && (rc = create_handles()) != 0) return rc;
	ix.primary = bwt->primary;
	memcpy(ix.L2, bwt->L2, sizeof ix.L2);
	ix.seq_len = bwt->seq_len;
	ix.bwt_size = bwt->bwt_size;
	ix.bwt = bwt->bwt;
	{
		/* How many launches in flight, and how many reads each can carry, follows from the batch size the reference calls with
		 * (-b, seen here for the first time): tiny batches are latency-bound -- one handle per worker thread keeps every worker's
		 * wait at one launch latency; large batches fill the GPU by themselves -- a few handles with room to combine four
		 * requests.  SMEM_GPU_ADAPTER_HANDLES / SMEM_GPU_MAX_BATCH override. */
		const char *eh = getenv("SMEM_GPU_ADAPTER_HANDLES"), *eb = getenv("SMEM_GPU_MAX_BATCH");
		int use = first_batch <= 256 ? 16 : first_batch <= 2048 ? 8 : 4;
		int64_t per = first_batch <= 1024 ? FIRST_BATCH : 4ll * first_batch;
		if (per > 65536) per = first_batch > 65536 ? first_batch : 65536;
		if (eb) per = atoll(eb);
		if (per < first_batch) per = first_batch;
		if (eh) use = atoi(eh);
		if (use > g_n_created) use = g_n_created;
		if (use < 1) use = 1;
		for (k = use; k < g_n_created; ++k) {          /* surplus handles go (they were cheap: small buffers) */
			svc_handle_t *s = &g_hd[k];
			smem_gpu_destroy(s->h);
			smem_gpu_host_free(s->seq); smem_gpu_host_free(s->offs); smem_gpu_host_free(s->roff); smem_gpu_host_free(s->x); smem_gpu_host_free(s->mi);
			smem_gpu_host_free(s->ret); smem_gpu_host_free(s->out); smem_gpu_host_free(s->tag); smem_gpu_host_free(s->ret16);
			memset(s, 0, sizeof *s);
		}
		g_n_created = use;
		g_max_batch = per;
		if (per > FIRST_BATCH)
			for (k = 0; k < g_n_created; ++k) {
				svc_handle_t *s = &g_hd[k];
				if ((rc = smem_gpu_resize(s->h, per, s->max_len)) != 0) return rc;
				s->max_batch = per;
				if ((rc = stage_grow(s, (size_t)per, (size_t)per * 110)) != 0) return rc;
				if ((rc = out_grow(s, 24 * (size_t)per + 4096)) != 0) return rc;
			}
	}
	for (k = 0; k < g_n_created; ++k) {     /* one index copy per GPU */
		svc_handle_t *s = &g_hd[k];
		int j;
		for (j = 0; j < k; ++j) if (g_hd[j].dev == s->dev) break;
		rc = j < k ? smem_gpu_share_index(s->h, g_hd[j].h) : smem_gpu_upload_index(s->h, &ix);
		if (rc != 0) return rc;
	}
	for (k = 0; k < g_n_created; ++k) {     /* one tiny launch of each kind per handle: nothing is first-time for a worker */
		svc_handle_t *s = &g_hd[k];
		smem_seed_opt_t opt;
		int64_t total = 0;
		memset(s->seq, 0, 64); s->seq[1] = 1; s->seq[2] = 2; s->seq[3] = 3;
		s->offs[0] = 0; s->offs[1] = 40; s->x[0] = 0; s->mi[0] = 1;
		opt.min_seed_len = g_split_len; opt.split_factor = 1.0; opt.split_width = g_split_width; opt.start_width = 1;
		if ((rc = smem_gpu_trace(s->h, 1, s->seq, s->offs, &opt, s->out, (int64_t)s->out_cap, s->roff, s->tag, s->ret16, &total)) != 0) return rc;
		if ((rc = smem_gpu_smem1(s->h, 1, s->seq, s->offs, s->x, s->mi, s->out, (int64_t)s->out_cap, s->roff, s->ret, &total)) != 0) return rc;
	}
	g_n_handles = g_n_created;
	if (g_bwt == 0 && getenv("SMEM_GPU_ADAPTER_STATS")) atexit(print_stats);
	g_bwt = bwt;
	return 0;
}

int harp_gpu_service_start(const harp_bwt_t *bwt)
{
	int rc;
	pthread_mutex_lock(&g_mu);
	rc = service_start_locked(bwt, 65536);
	pthread_mutex_unlock(&g_mu);
	return rc;
}

void harp_gpu_service_stop(void)
{
	int k;
	pthread_mutex_lock(&g_mu);
	while (g_pre_state == 1) pthread_cond_wait(&g_cv, &g_mu);
	for (k = 0; k < MAX_HANDLES; ++k) {
		svc_handle_t *s = &g_hd[k];
		if (s->h) smem_gpu_destroy(s->h);
		smem_gpu_host_free(s->seq); smem_gpu_host_free(s->offs); smem_gpu_host_free(s->roff); smem_gpu_host_free(s->x); smem_gpu_host_free(s->mi);
		smem_gpu_host_free(s->ret); smem_gpu_host_free(s->out); smem_gpu_host_free(s->tag); smem_gpu_host_free(s->ret16);
	}
	memset(g_hd, 0, sizeof g_hd);
	g_n_handles = 0; g_n_created = 0; g_bwt = 0;
	pthread_mutex_unlock(&g_mu);
}

void harp_gpu_service_stats(uint64_t out[4]) { memcpy(out, g_stats, 4 * sizeof(uint64_t)); }
void harp_gpu_service_stats8(uint64_t out[8]) { memcpy(out, g_stats, 8 * sizeof(uint64_t)); }

/* ---- pinned staging of a service handle ---------------------------------------------------------------------- */
static int pin_grow(void **p, size_t *cap, size_t need, size_t elem)
{
	void *q = 0;
	size_t c;
	if (need <= *cap) return 0;
	c = need + need / 2 + 1024;
	if (smem_gpu_host_alloc(&q, c * elem) != 0) return SMEM_GPU_E_NOMEM;
	if (*p) smem_gpu_host_free(*p);
	*p = q; *cap = c;
	return 0;
}

static int stage_grow(svc_handle_t *s, size_t n, size_t nbytes)
{
	int rc = 0;
	size_t c;
	if (n > s->n_cap) {
		c = 0; rc |= pin_grow((void **)&s->offs, &c, n + 1, sizeof(int64_t));
		c = 0; rc |= pin_grow((void **)&s->roff, &c, n + 1, sizeof(int64_t));
		c = 0; rc |= pin_grow((void **)&s->x, &c, n, sizeof(int32_t));
		c = 0; rc |= pin_grow((void **)&s->mi, &c, n, sizeof(int32_t));
		c = 0; rc |= pin_grow((void **)&s->ret, &c, n, sizeof(int32_t));
		s->n_cap = rc ? 0 : n + n / 2 + 1024;
	}
	if (!rc) rc = pin_grow((void **)&s->seq, &s->seq_cap, nbytes + 64, 1);
	return rc;
}

static int out_grow(svc_handle_t *s, size_t need)
{
	int rc = 0;
	size_t c;
	if (need <= s->out_cap) return 0;
	c = 0; rc |= pin_grow((void **)&s->out, &c, need, sizeof(smem_intv_t));
	c = 0; rc |= pin_grow((void **)&s->tag, &c, need, sizeof(uint16_t));
	c = 0; rc |= pin_grow((void **)&s->ret16, &c, need, sizeof(uint16_t));
	s->out_cap = rc ? 0 : need + need / 2 + 1024;
	return rc;
}

/* ---- the leader: every compatible pending request in one launch ----------------------------------------------- */
static void lead(svc_handle_t *s, req_t *list)
{
	req_t *r;
	size_t n = 0, nbytes = 0, k, pos;
	int max_len = 0, rc = 0, n_req = 0;
	int64_t total = 0;
	const uint64_t t0 = now_ns();
	uint64_t t_back, t_grown = 0, t_launched = 0, t1 = 0;
	static int trace = -1;
	if (trace < 0) trace = getenv("SMEM_GPU_ADAPTER_TRACE") != 0;
	for (r = list; r; r = r->next) {
		++n_req;
		for (k = 0; k < r->n; ++k) { const int l = r->itr[r->idx[k]]->len; nbytes += (size_t)l; if (l > max_len) max_len = l; }
		n += r->n;
	}
	if (max_len > s->max_len || (int64_t)n > s->max_batch) {    /* grow the handle in place; the index stays */
		int want_len = s->max_len;
		int64_t want_batch = s->max_batch;
		while (want_len < max_len) want_len = want_len * 2 > 65535 ? 65535 : want_len * 2;
		while (want_batch < (int64_t)n) want_batch *= 2;
		if (max_len > 65535) rc = SMEM_GPU_E_CAPACITY;
		else if ((rc = smem_gpu_resize(s->h, want_batch, want_len)) == 0) { s->max_len = want_len; s->max_batch = want_batch; stat_add(11, 1); }
	}
	{
		const size_t c0 = s->n_cap + s->seq_cap + s->out_cap;
		if (!rc) rc = stage_grow(s, n, nbytes);
		if (!rc) rc = out_grow(s, 24 * n + 4096);
		if (s->n_cap + s->seq_cap + s->out_cap != c0) stat_add(11, 1);
	}
	t_grown = now_ns();
	if (!rc) {
		for (r = list, pos = 0, nbytes = 0; r; r = r->next)
			for (k = 0; k < r->n; ++k, ++pos) {
				const harp_smem_i *it = r->itr[r->idx[k]];
				s->offs[pos] = (int64_t)nbytes;
				memcpy(s->seq + nbytes, it->query, (size_t)it->len);
				nbytes += (size_t)it->len;
				if (list->kind == 0) { s->x[pos] = r->x[k]; s->mi[pos] = r->mi[k]; }
			}
		s->offs[n] = (int64_t)nbytes;
		t1 = now_ns(); stat_add(8, t1 - t0);
		for (;;) {
			if (list->kind == 0)
				rc = smem_gpu_smem1(s->h, (int64_t)n, s->seq, s->offs, s->x, s->mi, s->out, (int64_t)s->out_cap, s->roff, s->ret, &total);
			else {
				smem_seed_opt_t opt;
				opt.min_seed_len = g_split_len; opt.split_factor = 1.0; opt.split_width = g_split_width; opt.start_width = list->start_width;
				rc = smem_gpu_trace(s->h, (int64_t)n, s->seq, s->offs, &opt, s->out, (int64_t)s->out_cap, s->roff, s->tag, s->ret16, &total);
			}
			if (rc == SMEM_GPU_E_CAPACITY && (size_t)total > s->out_cap) {   /* more intervals than guessed: size exactly and repeat */
				stat_add(11, 1);
				if ((rc = out_grow(s, (size_t)total)) == 0) continue;
			}
			break;
		}
		t_launched = now_ns();
		stat_add(9, t_launched - t1);
	}
	t_back = now_ns();
	if (rc) fail(list->kind ? "smem_gpu_trace" : "smem_gpu_smem1", rc, s->h);
	/* hand the slices back */
	for (r = list, pos = 0; r; r = r->next) {
		r->rc = rc;
		if (!rc) {
			const int64_t base = s->roff[pos];
			const size_t cnt = (size_t)(s->roff[pos + r->n] - base);
			if (cnt > *r->cap) {
				const size_t c = cnt + cnt / 2 + 256;
				*r->intv = (smem_intv_t *)realloc(*r->intv, c * sizeof(smem_intv_t));
				if (r->tag) { *r->tag = (uint16_t *)realloc(*r->tag, c * 2); *r->ret16 = (uint16_t *)realloc(*r->ret16, c * 2); }
				*r->cap = c;
			}
			memcpy(*r->intv, s->out + base, cnt * sizeof(smem_intv_t));
			if (r->tag) { memcpy(*r->tag, s->tag + base, cnt * 2); memcpy(*r->ret16, s->ret16 + base, cnt * 2); }
			for (k = 0; k <= r->n; ++k) r->roff[k] = s->roff[pos + k] - base;
			if (r->ret) memcpy(r->ret, s->ret + pos, r->n * sizeof(int32_t));
		} else
			for (k = 0; k <= r->n; ++k) r->roff[k] = 0;
		pos += r->n;
	}
	stat_add(0, 1); stat_add(1, n); stat_add(2, (uint64_t)total); stat_add(4, (uint64_t)n_req); stat_add(5, now_ns() - t0); stat_add(10, now_ns() - t_back);
	if (trace) fprintf(stderr, "[adapter trace] handle %d kind %d: %d requests, %zu reads, %lld intervals; grow %.2f gather %.2f launch %.2f hand-back %.2f ms\n", (int)(s - g_hd),
	                   list->kind, n_req, n, (long long)total, (t_grown - t0) * 1e-6, t1 ? (t1 - t_grown) * 1e-6 : 0.0,
	                   t_launched ? (t_launched - t1) * 1e-6 : 0.0, (now_ns() - t_back) * 1e-6);
}

/* post a request and wait for it; while waiting, lead launches whenever a service handle is free */
static int submit(req_t *me)
{
	int k;
	me->done = 0; me->rc = 0; me->next = 0;
	pthread_mutex_lock(&g_mu);
	if (g_tail) g_tail->next = me; else g_head = me;
	g_tail = me;
	while (!me->done) {
		svc_handle_t *s = 0;
		if (g_head) for (k = 0; k < g_n_handles; ++k) if (!g_hd[k].busy) { s = &g_hd[k]; break; }
		if (s) {
			/* take the head request and every later one that can share its launch */
			req_t *list = g_head, *lt = list, **pp, *r;
			int64_t n = (int64_t)list->n;
			g_head = list->next; if (!g_head) g_tail = 0;
			lt->next = 0;
			for (pp = &g_head, g_tail = 0; (r = *pp) != 0;) {
				if (r->kind == list->kind && r->start_width == list->start_width && n + (int64_t)r->n <= g_max_batch) {
					*pp = r->next; r->next = 0; lt->next = r; lt = r; n += (int64_t)r->n;
				} else { g_tail = r; pp = &r->next; }
			}
			s->busy = 1;
			pthread_mutex_unlock(&g_mu);
			lead(s, list);
			pthread_mutex_lock(&g_mu);
			s->busy = 0;
			for (r = list; r;) { req_t *nx = r->next; r->next = 0; r->done = 1; r = nx; }   /* (owners may return as soon as done is set) */
			pthread_cond_broadcast(&g_cv);
		} else
			pthread_cond_wait(&g_cv, &g_mu);
	}
	pthread_mutex_unlock(&g_mu);
	return me->rc;
}

/* ---- thread-local state --------------------------------------------------------------------------------------- */
typedef struct {
	/* whole-batch cache: the per-call lists of every round of the batch, from one smem_gpu_trace */
	int n;
	const harp_smem_i **itr; const uint8_t **query; int *len;
	int *next_pos;        /* where the next pass-1 call of the read must start (before skipping N) */
	int *step;            /* last step served by pass 1, -1 = none */
	int64_t *cur;         /* cursor into the read's trace entries */
	int64_t *roff;        /* [n+1] trace CSR, batch order (empty for reads that were done) */
	smem_intv_t *intv; uint16_t *tag, *ret; size_t cap;
	size_t n_cap;
	/* direct path / request scratch */
	int32_t *list, *lx, *lmi, *lret, *gidx;
	int64_t *groff;
	smem_intv_t *dintv; size_t dcap;
} tls_t;
static __thread tls_t tc;

static int skip_n(const uint8_t *q, int len, int pos) { while (pos < len && q[pos] > 3) ++pos; return pos; }

static void tls_grow(int n)
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
	tc.groff = (int64_t *)realloc(tc.groff, (tc.n_cap + 1) * sizeof(int64_t));
	tc.list = (int32_t *)realloc(tc.list, tc.n_cap * sizeof(int32_t));
	tc.lx = (int32_t *)realloc(tc.lx, tc.n_cap * sizeof(int32_t));
	tc.lmi = (int32_t *)realloc(tc.lmi, tc.n_cap * sizeof(int32_t));
	tc.lret = (int32_t *)realloc(tc.lret, tc.n_cap * sizeof(int32_t));
	tc.gidx = (int32_t *)realloc(tc.gidx, tc.n_cap * sizeof(int32_t));
}

static void tls_free(void)
{
	free((void *)tc.itr); free((void *)tc.query); free(tc.len); free(tc.next_pos); free(tc.step); free(tc.cur); free(tc.roff); free(tc.groff);
	free(tc.intv); free(tc.tag); free(tc.ret); free(tc.list); free(tc.lx); free(tc.lmi); free(tc.lret); free(tc.gidx); free(tc.dintv);
	memset(&tc, 0, sizeof tc);
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

/* ---- direct path: one bwt_smem1 per listed read ------------------------------------------------------------------
 * idx[k] = position in the batch; x/mi per listed read.  Results go to matches (pass 1, + start) or sub (pass 2). */
static void direct_smem1(harp_smem_i **itr, const int32_t *idx, size_t n, const int32_t *x, const int32_t *mi, int is_middle)
{
	req_t rq;
	size_t k;
	if (n == 0) return;
	memset(&rq, 0, sizeof rq);
	rq.kind = 0; rq.n = n; rq.itr = itr; rq.idx = idx; rq.x = x; rq.mi = mi;
	rq.intv = &tc.dintv; rq.cap = &tc.dcap; rq.roff = tc.groff; rq.ret = tc.lret;
	submit(&rq);
	for (k = 0; k < n; ++k) {   /* scatter into the caller-owned kvecs (bwt.c:719-749); after a failure every list is empty */
		harp_smem_i *it = itr[idx[k]];
		put_list(is_middle ? it->sub : it->matches, tc.dintv + tc.groff[k], (size_t)(tc.groff[k + 1] - tc.groff[k]));
		if (!is_middle) it->start = rq.rc ? it->len : tc.lret[k];
	}
}

/* ---- whole-batch cache: all rounds of the active reads of this batch in one (shared) launch ---------------------- */
static void cache_fill(harp_smem_i **itr, int batch_size, const int *done, int start_width)
{
	req_t rq;
	size_t n = 0, k;
	int i;
	for (i = 0; i < batch_size; ++i) {
		tc.itr[i] = itr[i]; tc.query[i] = done[i] ? 0 : itr[i]->query; tc.len[i] = done[i] ? 0 : itr[i]->len;
		tc.next_pos[i] = 0; tc.step[i] = -1;
		if (!done[i]) tc.gidx[n++] = i;
	}
	memset(&rq, 0, sizeof rq);
	rq.kind = 1; rq.start_width = start_width; rq.n = n; rq.itr = itr; rq.idx = tc.gidx;
	rq.intv = &tc.intv; rq.tag = &tc.tag; rq.ret16 = &tc.ret; rq.cap = &tc.cap; rq.roff = tc.groff;
	if (submit(&rq) != 0) { tc.n = 0; return; }      /* failure: nothing cached, the direct path reports it again and returns empty lists */
	/* trace CSR is in gathered order; spread it to batch order (reads that were done own an empty range) */
	for (i = 0, k = 0; i < batch_size; ++i) {
		tc.roff[i] = tc.groff[k]; tc.cur[i] = tc.groff[k];
		if (!done[i]) ++k;
	}
	tc.roff[batch_size] = tc.groff[n];
	tc.n = batch_size;
}

void bwt_smem1_batched(harp_smem_i **itr, int *ori_start, int *max_i, int start_width, int is_middle,
                       int batch_size, const int *done, int bwt_batched_status)
{
	int i;
	size_t n = 0, nf = 0;
	uint64_t t0;
	if (bwt_batched_status == HARP_BWT_BATCHED_FREE) { tls_free(); return; }
	if (bwt_batched_status != HARP_BWT_BATCHED_DO) return;   /* INIT: nothing thread-local to set up */
	if (batch_size <= 0) return;
	t0 = now_ns();

	if ((g_n_handles == 0 && !g_start_failed) || g_use_cache < 0) {               /* first call: bring the service up (bwa.c:289-301 has no hook to do it earlier) */
		int rc = 0;
		pthread_mutex_lock(&g_mu);
		if (g_use_cache < 0) {
			const char *e = getenv("SMEM_GPU_ADAPTER_CACHE");
			if ((e = getenv("SMEM_GPU_SPLIT_LEN")) != 0) g_split_len = atoi(e);
			if ((e = getenv("SMEM_GPU_SPLIT_WIDTH")) != 0) g_split_width = atoi(e);
			e = getenv("SMEM_GPU_ADAPTER_CACHE");
			g_use_cache = e ? atoi(e) != 0 : 1;
		}
		if (!g_start_failed)
			for (i = 0; i < batch_size; ++i)
				if (!done[i]) { rc = service_start_locked(itr[i]->bwt, batch_size); break; }
		pthread_mutex_unlock(&g_mu);
		if (rc) fail("service start", rc, 0);
		stat_add(7, now_ns() - t0);
	}
	tls_grow(batch_size);

	if (!is_middle) {
		/* pass 1 (bwt.c:514-546): x = ori_start[i], min_intv = max(start_width, 1) */
		int fresh = 0, active = 0;
		for (i = 0; i < batch_size; ++i) {
			if (done[i]) continue;
			++active;
			fresh += ori_start[i] == skip_n(itr[i]->query, itr[i]->len, 0);
		}
		if (g_use_cache && g_n_handles && active > 0 && fresh == active) cache_fill(itr, batch_size, done, start_width);   /* first round of a batch */
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
		if (g_n_handles) direct_smem1(itr, tc.list, n, tc.lx, tc.lmi, 0);
		else for (i = 0; i < (int)n; ++i) { itr[tc.list[i]]->matches->n = 0; itr[tc.list[i]]->start = itr[tc.list[i]]->len; }
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
		if (g_n_handles) direct_smem1(itr, tc.list, n, tc.lx, tc.lmi, 1);
		else for (i = 0; i < (int)n; ++i) itr[tc.list[i]]->sub->n = 0;
	}
	stat_add(3, nf);
	stat_add(6, now_ns() - t0);
}
