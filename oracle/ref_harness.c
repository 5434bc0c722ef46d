/* TEST INFRASTRUCTURE ONLY -- not part of the shipped product path.
 *
 * Flat-array harness around the UNMODIFIED reference objects compiled from
 * /root/reference/software (see oracle/Makefile, target _ref/libbwaref.so).
 * Everything here *calls* the reference's own functions:
 *   bwt_restore_bwt  (bwt.c:899)    bwt_smem1   (bwt.c:776)
 *   smem_itr_init    (bwamem.c:80)  smem_next2  (bwamem.c:244)
 *   smem_set_query   (bwamem.c:101) kt_for      (kthread.c:41)
 * and re-states only the enumeration loop of mem_insert_seed
 * (bwamem.c:453-460: split_len clamp, start_width, `while (smem_next2(...))`).
 *
 * Used (a) to pin the C restatement in oracle/smem_oracle.c, (b) to generate the
 * golden fixtures under tests/golden/, (c) as the `--impl reference` / cpu_baseline
 * arm of bench.py ("kind": "reference").  Never linked into the CUDA library.
 *
 * Read batch layout (shared with oracle/smem_oracle.c and include/smem_gpu.h):
 *   seq  : uint8, one base per byte, 0..3 = A,C,G,T, >3 ambiguous  (bwamem.c:1403-1406)
 *   offs : int64[n+1], read i occupies seq[offs[i] .. offs[i+1])
 * Result layout: intv uint64[total][4] = {x0,x1,x2,info} (bwt.h:60-62), read_off int64[n+1],
 *   step uint16[total] = index of the smem_next2 call (per read) that produced the interval.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <pthread.h>
#include <time.h>
#include "bwt.h"
#include "bwamem.h"

smem_i *smem_itr_init(const bwt_t *bwt);
void smem_itr_destroy(smem_i *itr);
void smem_set_query(smem_i *itr, int len, const uint8_t *query);
const bwtintv_v *smem_next2(smem_i *itr, int split_len, int split_width, int start_width);
void kt_for(int n_threads, void (*func)(void *, int, int), void *data, int n);

typedef struct {
	int min_seed_len;     /* mem_opt_t::min_seed_len, default 19 (bwamem.c:45-75) */
	double split_factor;  /* default 1.5 */
	int split_width;      /* default 10 */
	int start_width;      /* 1, or 2 with MEM_F_NO_EXACT (bwamem.c:457) */
} ref_seed_opt_t;

void *ref_bwt_load(const char *path) { return bwt_restore_bwt(path); }

/* Build a bwt_t around caller-owned words (no copy); free with ref_bwt_free_view. */
void *ref_bwt_view(uint64_t primary, const uint64_t L2[5], uint64_t seq_len, uint64_t bwt_size, uint32_t *words)
{
	bwt_t *b = (bwt_t *)calloc(1, sizeof(bwt_t));
	int i;
	b->primary = primary;
	for (i = 0; i < 5; ++i) b->L2[i] = L2[i];
	b->seq_len = seq_len;
	b->bwt_size = bwt_size;
	b->bwt = words;
	bwt_gen_cnt_table(b);
	return b;
}
void ref_bwt_free_view(void *p) { free(p); }
void ref_bwt_free(void *p) { bwt_destroy((bwt_t *)p); }

void ref_bwt_info(const void *p, uint64_t out[8])
{
	const bwt_t *b = (const bwt_t *)p;
	int i;
	out[0] = b->primary;
	for (i = 0; i < 5; ++i) out[1 + i] = b->L2[i];
	out[6] = b->seq_len;
	out[7] = b->bwt_size;
}
const uint32_t *ref_bwt_words(const void *p) { return ((const bwt_t *)p)->bwt; }

void ref_occ4(const void *p, uint64_t k, uint64_t cnt[4]) { bwt_occ4((const bwt_t *)p, k, cnt); }

/* ik = {x0,x1,x2}; ok = 4 x {x0,x1,x2} */
void ref_extend(const void *p, const uint64_t ik3[3], int is_back, uint64_t ok12[12])
{
	bwtintv_t ik, ok[4];
	int i;
	ik.x[0] = ik3[0]; ik.x[1] = ik3[1]; ik.x[2] = ik3[2]; ik.info = 0;
	bwt_extend((const bwt_t *)p, &ik, ok, is_back);
	for (i = 0; i < 4; ++i) { ok12[3*i] = ok[i].x[0]; ok12[3*i+1] = ok[i].x[1]; ok12[3*i+2] = ok[i].x[2]; }
}

/* Attach caller-owned SA samples to a bwt_t (view or loaded) and call the reference's bwt_sa (bwt.c:104). */
void ref_bwt_set_sa(void *p, int sa_intv, uint64_t n_sa, uint64_t *sa)
{
	bwt_t *b = (bwt_t *)p;
	b->sa_intv = sa_intv; b->n_sa = n_sa; b->sa = sa;
}
void ref_bwt_clear_sa(void *p) { ((bwt_t *)p)->sa = 0; ((bwt_t *)p)->n_sa = 0; }
void ref_sa(const void *p, int64_t n, const uint64_t *k, uint64_t *out)
{
	int64_t i;
	for (i = 0; i < n; ++i) out[i] = bwt_sa((const bwt_t *)p, k[i]);
}

/* ---------------- growable flat result ---------------- */
typedef struct { uint64_t *v; uint16_t *step; int64_t n, m; } flat_t;
static void flat_push(flat_t *f, const bwtintv_t *p, int step)
{
	if (f->n == f->m) {
		f->m = f->m ? f->m << 1 : 1024;
		f->v = (uint64_t *)realloc(f->v, (size_t)f->m * 32);
		f->step = (uint16_t *)realloc(f->step, (size_t)f->m * 2);
	}
	f->v[4*f->n] = p->x[0]; f->v[4*f->n+1] = p->x[1]; f->v[4*f->n+2] = p->x[2]; f->v[4*f->n+3] = p->info;
	f->step[f->n++] = (uint16_t)step;
}

typedef struct {
	const bwt_t *bwt; const uint8_t *seq; const int64_t *offs; ref_seed_opt_t opt;
	int64_t lo, hi;                 /* read range of this worker */
	flat_t out; int64_t *cnt;       /* cnt[i-lo] = intervals of read i */
	int32_t *n_steps;               /* optional, global array */
	int32_t *last_start;            /* optional, global array: itr->start when the loop ended */
} collect_job_t;

/* mem_insert_seed's enumeration loop, bwamem.c:453-460 (chaining body omitted: not on the path). */
static void collect_one(smem_i *itr, const uint8_t *q, int len, const ref_seed_opt_t *o,
                        flat_t *out, int64_t *cnt, int32_t *n_steps, int32_t *last_start)
{
	const bwtintv_v *a;
	int split_len = (int)(o->min_seed_len * (float)o->split_factor + .499);   /* float product as with mem_opt_t (bwamem.c:456) */
	int step = 0;
	int64_t n0 = out->n;
	split_len = split_len < len ? split_len : len;
	smem_set_query(itr, len, q);
	while ((a = smem_next2(itr, split_len, o->split_width, o->start_width)) != 0) {
		size_t i;
		for (i = 0; i < a->n; ++i) flat_push(out, &a->a[i], step);
		++step;
	}
	*cnt = out->n - n0;
	if (n_steps) *n_steps = step;
	if (last_start) *last_start = itr->start;
}

static void *collect_worker(void *data)
{
	collect_job_t *j = (collect_job_t *)data;
	smem_i *itr = smem_itr_init(j->bwt);
	int64_t i;
	for (i = j->lo; i < j->hi; ++i)
		collect_one(itr, j->seq + j->offs[i], (int)(j->offs[i+1] - j->offs[i]), &j->opt, &j->out, &j->cnt[i - j->lo],
		            j->n_steps ? &j->n_steps[i] : 0, j->last_start ? &j->last_start[i] : 0);
	smem_itr_destroy(itr);
	return 0;
}

/* Returns total interval count; if > cap nothing is written except read_off. */
int64_t ref_collect(const void *bwt, int64_t n, const uint8_t *seq, const int64_t *offs, const ref_seed_opt_t *opt,
                    int nthreads, uint64_t *intv, int64_t cap, int64_t *read_off, uint16_t *step,
                    int32_t *n_steps, int32_t *last_start)
{
	collect_job_t *jobs;
	pthread_t *tid;
	int t;
	int64_t total = 0, i, per;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
	jobs = (collect_job_t *)calloc(nthreads, sizeof(collect_job_t));
	tid = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
	per = (n + nthreads - 1) / nthreads;
	for (t = 0; t < nthreads; ++t) {
		collect_job_t *j = &jobs[t];
		j->bwt = (const bwt_t *)bwt; j->seq = seq; j->offs = offs; j->opt = *opt;
		j->lo = t * per < n ? t * per : n; j->hi = (t + 1) * per < n ? (t + 1) * per : n;
		j->cnt = (int64_t *)calloc(j->hi - j->lo + 1, sizeof(int64_t));
		j->n_steps = n_steps; j->last_start = last_start;
		pthread_create(&tid[t], 0, collect_worker, j);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(tid[t], 0);
	read_off[0] = 0;
	for (t = 0; t < nthreads; ++t)
		for (i = jobs[t].lo; i < jobs[t].hi; ++i) { total += jobs[t].cnt[i - jobs[t].lo]; read_off[i+1] = total; }
	if (total <= cap) {
		int64_t pos = 0;
		for (t = 0; t < nthreads; ++t) {
			if (jobs[t].out.n) {
				memcpy(intv + 4 * pos, jobs[t].out.v, (size_t)jobs[t].out.n * 32);
				if (step) memcpy(step + pos, jobs[t].out.step, (size_t)jobs[t].out.n * 2);
			}
			pos += jobs[t].out.n;
		}
	}
	for (t = 0; t < nthreads; ++t) { free(jobs[t].out.v); free(jobs[t].out.step); free(jobs[t].cnt); }
	free(jobs); free(tid);
	return total;
}

/* One raw bwt_smem1 call per read (bwt.c:776): x[i], min_intv[i] -> ret[i] + intervals. Single thread. */
int64_t ref_smem1(const void *bwt, int64_t n, const uint8_t *seq, const int64_t *offs, const int32_t *x,
                  const int32_t *min_intv, uint64_t *intv, int64_t cap, int64_t *read_off, int32_t *ret)
{
	bwtintv_v mem = {0, 0, 0};
	int64_t i, total = 0;
	size_t k;
	read_off[0] = 0;
	for (i = 0; i < n; ++i) {
		int len = (int)(offs[i+1] - offs[i]);
		ret[i] = bwt_smem1((const bwt_t *)bwt, len, seq + offs[i], x[i], min_intv[i], &mem, 0);
		for (k = 0; k < mem.n; ++k, ++total)
			if (total < cap) {
				intv[4*total] = mem.a[k].x[0]; intv[4*total+1] = mem.a[k].x[1];
				intv[4*total+2] = mem.a[k].x[2]; intv[4*total+3] = mem.a[k].info;
			}
		read_off[i+1] = total;
	}
	free(mem.a);
	return total;
}

/* ---------------- timing arm (reference's own kt_for, one smem_i per thread) ---------------- */
#define FNV_BASIS 0xcbf29ce484222325ull
#define FNV_PRIME 0x100000001b3ull

typedef struct {
	const bwt_t *bwt; const uint8_t *seq; const int64_t *offs; ref_seed_opt_t opt;
	smem_i **itr; uint64_t *sum; int64_t *cnt; int chunk; int64_t n;
} time_job_t;

static void time_worker(void *data, int ci, int tid)
{
	time_job_t *j = (time_job_t *)data;
	int64_t i, lo = (int64_t)ci * j->chunk, hi = lo + j->chunk < j->n ? lo + j->chunk : j->n;
	smem_i *itr = j->itr[tid];
	for (i = lo; i < hi; ++i) {
		const bwtintv_v *a;
		int len = (int)(j->offs[i+1] - j->offs[i]);
		int split_len = (int)(j->opt.min_seed_len * (float)j->opt.split_factor + .499);
		uint64_t h = FNV_BASIS ^ (uint64_t)i;
		split_len = split_len < len ? split_len : len;
		smem_set_query(itr, len, j->seq + j->offs[i]);
		while ((a = smem_next2(itr, split_len, j->opt.split_width, j->opt.start_width)) != 0) {
			size_t k;
			for (k = 0; k < a->n; ++k) {
				h = (h ^ a->a[k].x[0]) * FNV_PRIME; h = (h ^ a->a[k].x[1]) * FNV_PRIME;
				h = (h ^ a->a[k].x[2]) * FNV_PRIME; h = (h ^ a->a[k].info) * FNV_PRIME;
			}
			j->cnt[tid] += a->n;
		}
		j->sum[tid] += h;
	}
}

/* Wall seconds for seeding n reads with nthreads host threads; file I/O and index load excluded. */
double ref_time_collect(const void *bwt, int64_t n, const uint8_t *seq, const int64_t *offs, const ref_seed_opt_t *opt,
                        int nthreads, uint64_t *checksum, int64_t *n_intervals)
{
	time_job_t j;
	struct timespec t0, t1;
	int t;
	if (nthreads < 1) nthreads = 1;
	j.bwt = (const bwt_t *)bwt; j.seq = seq; j.offs = offs; j.opt = *opt; j.chunk = 256; j.n = n;
	j.itr = (smem_i **)calloc(nthreads, sizeof(smem_i *));
	j.sum = (uint64_t *)calloc(nthreads, sizeof(uint64_t));
	j.cnt = (int64_t *)calloc(nthreads, sizeof(int64_t));
	for (t = 0; t < nthreads; ++t) j.itr[t] = smem_itr_init(j.bwt);
	clock_gettime(CLOCK_MONOTONIC, &t0);
	kt_for(nthreads, time_worker, &j, (int)((n + j.chunk - 1) / j.chunk));
	clock_gettime(CLOCK_MONOTONIC, &t1);
	*checksum = 0; *n_intervals = 0;
	for (t = 0; t < nthreads; ++t) { *checksum += j.sum[t]; *n_intervals += j.cnt[t]; smem_itr_destroy(j.itr[t]); }
	free(j.itr); free(j.sum); free(j.cnt);
	return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---------------- chaining: the reference's own mem_chain + mem_chain_flt ---------------- */
/* Calls mem_chain (bwamem.c:593-615: seeding, bwt_sa, kbtree insertion with test_and_merge bwamem.c:334-356, traversal in
 * pos order) and optionally mem_chain_flt (bwamem.c:629-700) for every read; `bwt` must carry SA samples
 * (ref_bwt_set_sa).  Flat result: chain_off int64[n+1] (CSR per read), chain int64[total][2] = {pos, n_seeds},
 * seeds = mem_seed_t[] {int64 rbeg; int32 qbeg, len} grouped by chain in chain order.  Returns the chain total;
 * *n_seeds_out = seed total; nothing but chain_off / the totals is complete when a cap is exceeded. */
typedef struct { int64_t rbeg; int32_t qbeg, len; } h_seed_t;
typedef struct { int n, m; int64_t pos; h_seed_t *seeds; } h_chain_t;
typedef struct { size_t n, m; h_chain_t *a; } h_chain_v;
h_chain_v mem_chain(const mem_opt_t *opt, const bwt_t *bwt, int64_t l_pac, int len, const uint8_t *seq);
int mem_chain_flt(const mem_opt_t *opt, int n_chn, h_chain_t *chains);

typedef struct {
	ref_seed_opt_t seed;
	int w, max_chain_gap, max_occ;
	float mask_level, chain_drop_ratio;
} ref_chain_opt_t;

int64_t ref_chains(const void *bwt, int64_t l_pac, int64_t n, const uint8_t *seq, const int64_t *offs, const ref_chain_opt_t *o,
                   int do_flt, int64_t *chain_off, int64_t *chain, int64_t chain_cap, h_seed_t *seeds, int64_t seed_cap,
                   int64_t *n_seeds_out)
{
	mem_opt_t *opt = mem_opt_init();
	int64_t i, nc = 0, ns = 0;
	opt->min_seed_len = o->seed.min_seed_len; opt->split_factor = (float)o->seed.split_factor; opt->split_width = o->seed.split_width;
	if (o->seed.start_width == 2) opt->flag |= MEM_F_NO_EXACT;
	opt->w = o->w; opt->max_chain_gap = o->max_chain_gap; opt->max_occ = o->max_occ;
	opt->mask_level = o->mask_level; opt->chain_drop_ratio = o->chain_drop_ratio;
	chain_off[0] = 0;
	for (i = 0; i < n; ++i) {
		h_chain_v c = mem_chain(opt, (const bwt_t *)bwt, l_pac, (int)(offs[i+1] - offs[i]), seq + offs[i]);
		size_t k;
		if (do_flt) c.n = (size_t)mem_chain_flt(opt, (int)c.n, c.a);
		for (k = 0; k < c.n; ++k) {
			int t;
			if (nc < chain_cap) { chain[2*nc] = c.a[k].pos; chain[2*nc+1] = c.a[k].n; }
			++nc;
			for (t = 0; t < c.a[k].n; ++t, ++ns) if (ns < seed_cap) seeds[ns] = c.a[k].seeds[t];
			free(c.a[k].seeds);
		}
		free(c.a);
		chain_off[i+1] = nc;
	}
	free(opt);
	*n_seeds_out = ns;
	return nc;
}
