/* TEST INFRASTRUCTURE ONLY -- see smem_oracle.h.  Parity status: PINNED against oracle/_ref.
 *
 * A from-scratch CPU restatement of BWA-MEM 0.7.8's SMEM seeding as found in the reference:
 *   orc_occ4     <- bwt_occ4        bwt.c:187-204   (popcount on bit planes instead of the LUT)
 *   extend       <- bwt_extend      bwt.c:416-429   (bwt_2occ4 = two occ4, bwt.c:213-214)
 *   smem1        <- bwt_smem1       bwt.c:776-835
 *   next2        <- smem_next2      bwamem.c:244-305 (0.7.8 re-seed + ordered merge)
 *   collect_read <- mem_insert_seed bwamem.c:453-460 (enumeration loop only)
 */
#include "smem_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

typedef struct { uint64_t x0, x1, x2, info; } iv_t;
typedef struct { iv_t *a; int n, m; } ivv_t;

static void ivv_push(ivv_t *v, iv_t e)
{
	if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (iv_t *)realloc(v->a, sizeof(iv_t) * (size_t)v->m); }
	v->a[v->n++] = e;
}
static void ivv_reverse(ivv_t *v)
{
	int i, j;
	for (i = 0, j = v->n - 1; i < j; ++i, --j) { iv_t t = v->a[i]; v->a[i] = v->a[j]; v->a[j] = t; }
}

/* Occurrences of A,C,G,T in BWT[0..k] (inclusive); k counts the '$' row, hence the primary shift
 * (bwt.c:194).  Block = 16 words per 128 symbols: 4 x uint64 counts, then 8 words of 16 symbols,
 * first symbol in bits 31:30 (bwtindex.c:101,137-143). */
void orc_occ4(const orc_index_t *ix, uint64_t k, uint64_t cnt[4])
{
	const uint32_t *blk;
	uint64_t base[4];
	unsigned rem, w;
	uint32_t c1 = 0, c2 = 0, c3 = 0;
	if (k == (uint64_t)-1) { cnt[0] = cnt[1] = cnt[2] = cnt[3] = 0; return; }
	k -= (k >= ix->primary);
	blk = ix->bwt + ((k >> 7) << 4);
	memcpy(base, blk, 32);
	rem = (unsigned)(k & 127) + 1; /* symbols of this block to count */
	for (w = 0; rem > 0; ++w) {
		unsigned take = rem < 16 ? rem : 16;
		uint32_t v = blk[8 + w];
		uint32_t keep = take == 16 ? 0x55555555u : (0x55555555u & ~((1u << (2 * (16 - take))) - 1));
		uint32_t hi = (v >> 1) & keep, lo = v & keep;
		c3 += (uint32_t)__builtin_popcount(hi & lo);
		c2 += (uint32_t)__builtin_popcount(hi & ~lo);
		c1 += (uint32_t)__builtin_popcount(~hi & lo);
		rem -= take;
	}
	cnt[1] = base[1] + c1; cnt[2] = base[2] + c2; cnt[3] = base[3] + c3;
	cnt[0] = base[0] + ((unsigned)(k & 127) + 1 - c1 - c2 - c3);
}

static inline void extend(const orc_index_t *ix, const iv_t *ik, int is_back, iv_t ok[4], orc_stats_t *st)
{
	uint64_t tk[4], tl[4], a = is_back ? ik->x0 : ik->x1, b = is_back ? ik->x1 : ik->x0; /* a = x[!is_back] */
	uint64_t k = a - 1, l = a - 1 + ik->x2, acc;
	int c;
	orc_occ4(ix, k, tk);
	orc_occ4(ix, l, tl);
	if (st) {
		uint64_t kk = k - (k >= ix->primary), ll = l - (l >= ix->primary);
		st->extends++;
		st->blocks += (kk >> 7) == (ll >> 7) ? 1 : 2;
	}
	/* the other strand's start: bases are laid out T,G,C,A after the (possible) '$' (bwt.c:425-428) */
	acc = b + (a <= ix->primary && a + ik->x2 - 1 >= ix->primary);
	for (c = 3; c >= 0; --c) {
		uint64_t na = ix->L2[c] + 1 + tk[c], sz = tl[c] - tk[c];
		ok[c].x2 = sz;
		if (is_back) { ok[c].x0 = na; ok[c].x1 = acc; } else { ok[c].x1 = na; ok[c].x0 = acc; }
		acc += sz;
	}
}

typedef struct { ivv_t prev, curr, mem, sub, merged; } scratch_t;

/* MODEL of the device kernel's second shortcut, off unless orc_set_spec_walk(1) -- a test hook.
 * When a pass-1 bwt_smem1 at x directly follows one at x_prev whose forward sweep ended at ret == x because the interval
 * fell below min_intv, the pattern q[x_prev..x+1) occurs fewer than min_intv times; every candidate of the new call's
 * backward sweep with start <= x_prev contains it, so the sweep ends at start x_prev with everything dead.  While the
 * LONGEST candidate is alive nothing is emitted (bwt.c:815: curr->n != 0), so if it survives down to start x_prev + 1 (or to
 * an ambiguous base / the read start, where everything dies as well) the call's result is that one candidate -- the
 * shorter candidates never needed to be extended.  If it dies earlier by count, the full sweep is run after all. */
static int orc_spec_walk = 0;
static __thread int spec_lim = -1;          /* x_prev of the call about to be made, or -1 */
uint64_t orc_spec_hits = 0, orc_spec_misses = 0;
void orc_set_spec_walk(int on) { orc_spec_walk = on; orc_spec_hits = orc_spec_misses = 0; }
uint64_t orc_get_spec_hits(void) { return orc_spec_hits; }
uint64_t orc_get_spec_misses(void) { return orc_spec_misses; }

/* COUNTING MODEL of the device kernel's unique forward walk, off unless orc_set_unique_walk(1) -- a
 * hook of bench.py's `executed` figure; results are untouched.  Once the forward sweep of a pass-1 call (min_intv == 1) has
 * extended an interval of size 1 `run` times and at least `left` read bases remain, the device replaces the rest of the
 * sweep by three gathers (suffix array, text, inverse suffix array); they are counted as three extends of one block each
 * and the extends they replace are not counted (a walk longer than the 128-base text window costs one more gather per
 * window, which this model leaves out: 101 bp reads never need it). */
static int orc_uw_model = 0, orc_uw_run = 3, orc_uw_left = 8;
void orc_set_unique_walk(int on) { orc_uw_model = on; }

/* EXECUTABLE MODEL of the same rule (tests/test_repeat_filter.py), off unless orc_set_unique_walk_tables is given tables:
 * text = T = forward + reverse complement, one base per byte, n = seq_len bases; fsa[r] = text position of row r
 * (fsa[0] = n for the '$' row), isa = its inverse.  With them the rest of a forward sweep that holds an interval of size 1
 * is computed the way PH_UW_SA / PH_UW_TEXT / PH_UW_ISA of the device kernel do it -- t = fsa[x0], compare the read with
 * the text behind the match, x1 = isa[n - t - length] -- instead of by bwt_extend. */
static const uint8_t *orc_uw_text = 0;
static const uint64_t *orc_uw_fsa = 0, *orc_uw_isa = 0;
static uint64_t orc_uw_n = 0;
uint64_t orc_uw_walks = 0;
void orc_set_unique_walk_tables(const uint8_t *text, const uint64_t *fsa, const uint64_t *isa, uint64_t n, int run, int left)
{
	orc_uw_text = text; orc_uw_fsa = fsa; orc_uw_isa = isa; orc_uw_n = n; orc_uw_walks = 0;
	orc_uw_run = run > 0 ? run : 3; orc_uw_left = left > 0 ? left : 8;
}
uint64_t orc_get_unique_walks(void) { return orc_uw_walks; }

/* MODEL of the device kernel's split backward sweep, off unless orc_set_split_pairs(1) -- a test hook.  The
 * entries of one backward round are independent but for two things: an entry whose new size equals the last KEPT entry's is
 * dropped, and the first entry to die while nothing has been kept yet may be recorded.  The device therefore extends prev[j] and
 * prev[j + 1] in the same trip (one lane each), exchanges the two new sizes and replays both decisions in order.  This model does
 * exactly that on the CPU -- both extends first, from the untouched entries, then the decisions from the sizes alone, the second
 * emission of a pair suppressed the way the device suppresses it -- so that the suite checks the rule against the sequential loop of
 * bwt.c:812-825 without a GPU. */
static int orc_split_pairs = 0;
uint64_t orc_split_trips = 0;
void orc_set_split_pairs(int on) { orc_split_pairs = on; orc_split_trips = 0; }
uint64_t orc_get_split_trips(void) { return orc_split_trips; }

/* bwt_smem1, bwt.c:776-835.  Writes the SMEM candidates through x into `mem`, returns ret. */
static int smem1(const orc_index_t *ix, int len, const uint8_t *q, int x, int min_intv, ivv_t *mem,
                 ivv_t *prev, ivv_t *curr, orc_stats_t *st)
{
	iv_t ik, ok[4];
	int i, j, ret, uw_run = 0, uw_walking = 0;
	ivv_t *t;
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	if (st) st->smem1_calls++;
	ik.x0 = ix->L2[q[x]] + 1; ik.x1 = ix->L2[3 - q[x]] + 1; ik.x2 = ix->L2[q[x] + 1] - ix->L2[q[x]];
	ik.info = (uint64_t)(x + 1);
	curr->n = 0;
	for (i = x + 1; i < len; ++i) { /* forward: extend with the complement on the reverse strand */
		int c;
		if (q[i] > 3) { ivv_push(curr, ik); break; }
		c = 3 - q[i];
		extend(ix, &ik, 0, ok, uw_walking ? 0 : st);
		if (ok[c].x2 != ik.x2) {
			ivv_push(curr, ik);
			if (ok[c].x2 < (uint64_t)min_intv) break;
		}
		ik = ok[c]; ik.info = (uint64_t)(i + 1);
		if (orc_uw_text && min_intv == 1) {
			uw_run = ik.x2 == 1 ? uw_run + 1 : 0;
			if (uw_run >= orc_uw_run && len - (i + 1) >= orc_uw_left && q[i + 1] <= 3) {
				const uint64_t t = orc_uw_fsa[ik.x0];                 /* the pattern q[x..i] occurs only here */
				uint64_t plen = (uint64_t)(i + 1 - x);
				int e = i + 1;
				while (e < len && q[e] <= 3 && t + plen < orc_uw_n && orc_uw_text[t + plen] == q[e]) { ++e; ++plen; }
				ik.x1 = orc_uw_isa[orc_uw_n - (t + plen)];            /* rc(pattern) sits mirrored in T */
				ik.info = (uint64_t)e;
				/* read end, ambiguous base, text end or a differing base: the reference's next bwt_extend finds nothing
				 * (bwt.c:796-798) or the loop ends (bwt.c:800-806); either way ik is pushed and the sweep is over */
				ivv_push(curr, ik);
				__sync_fetch_and_add(&orc_uw_walks, 1);
				i = -2;                                               /* (not len: the push after the loop is skipped) */
				break;
			}
		} else if (orc_uw_model && st && !uw_walking && min_intv == 1) {
			uw_run = ik.x2 == 1 ? uw_run + 1 : 0;
			if (uw_run >= orc_uw_run && len - (i + 1) >= orc_uw_left && q[i + 1] <= 3) { uw_walking = 1; st->extends += 3; st->blocks += 3; }
		}
	}
	if (i == len) ivv_push(curr, ik);
	ivv_reverse(curr); /* longest match first */
	ret = (int)curr->a[0].info;
	if (st && (uint64_t)curr->n > st->max_curr) st->max_curr = (uint64_t)curr->n;
	t = prev; prev = curr; curr = t;
	if (orc_spec_walk && spec_lim >= 0 && spec_lim < x) {
		iv_t p = prev->a[0];
		int ok_ = 1;
		for (i = x - 1; i > spec_lim; --i) {
			int c = q[i] < 4 ? q[i] : -1;
			if (c < 0) break;                                   /* everything dies here: (i+1, end of the longest) */
			extend(ix, &p, 1, ok, st);
			if (ok[c].x2 < (uint64_t)min_intv) { ok_ = 0; break; } /* died by count before the limit: the others matter */
			ok[c].info = p.info; p = ok[c];
		}
		if (ok_) {
			p.info |= (uint64_t)(i + 1) << 32;
			ivv_push(mem, p);
			__sync_fetch_and_add(&orc_spec_hits, 1);
			return ret;
		}
		__sync_fetch_and_add(&orc_spec_misses, 1);
	}
	for (i = x - 1; i >= -1; --i) { /* backward */
		int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		curr->n = 0;
		if (orc_split_pairs && c >= 0) {
			uint64_t last_s = 0;
			for (j = 0; j < prev->n; j += 2) {
				const int two = j + 1 < prev->n;
				iv_t okA[4], okB[4], *pA = &prev->a[j], *pB = two ? &prev->a[j + 1] : 0;
				uint64_t sA, sB = 0;
				int smallA, smallB = 0, emitA, pushA, emitB, pushB;
				extend(ix, pA, 1, okA, st);
				if (two) extend(ix, pB, 1, okB, st);
				__sync_fetch_and_add(&orc_split_trips, 1);
				sA = okA[c].x2; if (two) sB = okB[c].x2;
				smallA = sA < (uint64_t)min_intv; smallB = two && sB < (uint64_t)min_intv;
				emitA = smallA && curr->n == 0;
				pushA = !smallA && (curr->n == 0 || sA != last_s);
				if (emitA && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) {
					iv_t e = *pA; e.info |= (uint64_t)(i + 1) << 32; ivv_push(mem, e);
				}
				if (pushA) { okA[c].info = pA->info; ivv_push(curr, okA[c]); last_s = sA; }
				emitB = two && !emitA && smallB && curr->n == 0;
				pushB = two && !(sB < (uint64_t)min_intv) && (curr->n == 0 || sB != last_s);
				if (emitB && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) {
					iv_t e = *pB; e.info |= (uint64_t)(i + 1) << 32; ivv_push(mem, e);
				}
				if (pushB) { okB[c].info = pB->info; ivv_push(curr, okB[c]); last_s = sB; }
			}
		} else
		for (j = 0; j < prev->n; ++j) {
			iv_t *p = &prev->a[j];
			extend(ix, p, 1, ok, st);
			if (c < 0 || ok[c].x2 < (uint64_t)min_intv) {
				if (curr->n == 0 && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) {
					iv_t e = *p;
					e.info |= (uint64_t)(i + 1) << 32;
					ivv_push(mem, e);
				}
			} else if (curr->n == 0 || ok[c].x2 != curr->a[curr->n - 1].x2) {
				ok[c].info = p->info;
				ivv_push(curr, ok[c]);
			}
		}
		if (curr->n == 0) break;
		t = prev; prev = curr; curr = t;
	}
	ivv_reverse(mem); /* ascending start */
	if (st && (uint64_t)mem->n > st->max_mem) st->max_mem = (uint64_t)mem->n;
	return ret;
}

static inline int iv_len(const iv_t *p) { return (int)((uint32_t)p->info - (uint32_t)(p->info >> 32)); }

/* MODEL of the device kernel's re-seeding shortcut, off unless orc_set_skip_kmer(D > 0) -- a test hook
 * that lets the CPU suite check the argument behind it on real data:
 *   the re-seeding pass (bwamem.c:272-300) only contributes entries of length >= max>>1 (the KEEP test of the merge), every
 *   such entry is a pattern through `mid` occurring >= min_intv >= 2 times, and a pattern of length >= D through `mid`
 *   contains a D-mer window through `mid`; so if max>>1 >= D and every D-mer window of the read through `mid` occurs at most
 *   once in the text, the pass cannot contribute and smem_next2's result is its pass-1 list.
 * Window counts here are exact (backward search); the device asks a conservative bit table ("may occur twice"). */
static int orc_skip_kmer = 0;
uint64_t orc_pass2_skipped = 0;
void orc_set_skip_kmer(int D) { orc_skip_kmer = D; orc_pass2_skipped = 0; }
uint64_t orc_get_pass2_skipped(void) { return orc_pass2_skipped; }

static int pass2_is_void(const orc_index_t *ix, int len, const uint8_t *q, int mid, int max)
{
	const int D = orc_skip_kmer;
	int a, lo, hi, k;
	if (D <= 0 || (max >> 1) < D || len < D) return 0;
	lo = mid - D + 1 < 0 ? 0 : mid - D + 1;
	hi = mid < len - D ? mid : len - D;
	for (k = lo; k < hi + D; ++k) if (q[k] > 3) return 0;          /* (the device does not bother with ambiguous bases either) */
	for (a = lo; a <= hi; ++a) {
		iv_t ik, ok[4];
		int c = q[a + D - 1];
		ik.x0 = ix->L2[c] + 1; ik.x2 = ix->L2[c + 1] - ix->L2[c]; ik.x1 = ix->L2[3 - c] + 1; ik.info = 0;
		for (k = a + D - 2; k >= a && ik.x2 > 0; --k) { extend(ix, &ik, 1, ok, 0); ik = ok[q[k]]; }
		if (ik.x2 > 1) return 0;
	}
	return 1;
}

static __thread int last_x = -1, last_ret = -1;   /* previous pass-1 call of the read being seeded (model of the second shortcut) */

/* smem_next2, bwamem.c:244-305.  Returns 0 when exhausted, else 1 with the step's list in s->mem. */
static int next2(const orc_index_t *ix, int len, const uint8_t *q, int *start, int split_len, int split_width,
                 int start_width, scratch_t *s, orc_stats_t *st)
{
	int i, max = 0, max_i = 0, ori_start;
	s->mem.n = s->sub.n = 0;
	if (*start >= len || *start < 0) return 0;
	while (*start < len && q[*start] > 3) ++*start;
	if (*start == len) return 0;
	ori_start = *start;
	spec_lim = (last_x >= 0 && last_ret == ori_start) ? last_x : -1;      /* see the model above */
	*start = smem1(ix, len, q, ori_start, start_width, &s->mem, &s->prev, &s->curr, st);
	spec_lim = -1;
	last_x = ori_start; last_ret = *start;
	if (s->mem.n == 0) return 1;
	for (i = 0; i < s->mem.n; ++i) { /* first maximum wins (bwamem.c:266-270) */
		int l = iv_len(&s->mem.a[i]);
		if (max < l) { max = l; max_i = i; }
	}
	if (split_len > 0 && max >= split_len && s->mem.a[max_i].x2 <= (uint64_t)split_width) {
		const iv_t *p = &s->mem.a[max_i];
		int j, mid = (int)(((uint32_t)p->info + (uint32_t)(p->info >> 32)) >> 1);
		ivv_t *a = &s->merged;
		if (pass2_is_void(ix, len, q, mid, max)) { __sync_fetch_and_add(&orc_pass2_skipped, 1); return 1; }
		smem1(ix, len, q, mid, (int)(p->x2 + 1), &s->sub, &s->prev, &s->curr, st);
		a->n = 0;
		i = j = 0;
#define KEY(e) ((int64_t)(((e).info >> 32 << 32) | (uint32_t)(len - (int)(uint32_t)(e).info)))
#define KEEP(e) (iv_len(&(e)) >= (max >> 1) && (int)(uint32_t)(e).info > ori_start)
		while (i < s->mem.n && j < s->sub.n) {
			if (KEY(s->mem.a[i]) < KEY(s->sub.a[j])) ivv_push(a, s->mem.a[i++]);
			else { if (KEEP(s->sub.a[j])) ivv_push(a, s->sub.a[j]); ++j; }
		}
		for (; i < s->mem.n; ++i) ivv_push(a, s->mem.a[i]);
		for (; j < s->sub.n; ++j) if (KEEP(s->sub.a[j])) ivv_push(a, s->sub.a[j]);
#undef KEY
#undef KEEP
		s->mem.n = 0;
		for (i = 0; i < a->n; ++i) ivv_push(&s->mem, a->a[i]);
	}
	return 1;
}

typedef void (*emit_fn)(void *ctx, const iv_t *e, int step);

static void collect_read(const orc_index_t *ix, int len, const uint8_t *q, const orc_seed_opt_t *o, scratch_t *s,
                         orc_stats_t *st, emit_fn emit, void *ctx, int32_t *n_steps, int32_t *last_start)
{
	/* mem_opt_t::split_factor is a float (bwamem.h:47): float product, then the double addition of bwamem.c:456 */
	int split_len = (int)((double)((float)o->min_seed_len * (float)o->split_factor) + .499), start = 0, step = 0, i;
	if (split_len > len) split_len = len;
	last_x = last_ret = -1;
	while (next2(ix, len, q, &start, split_len, o->split_width, o->start_width, s, st)) {
		for (i = 0; i < s->mem.n; ++i) emit(ctx, &s->mem.a[i], step);
		if (st) { st->steps++; st->intervals += (uint64_t)s->mem.n; }
		++step;
	}
	if (n_steps) *n_steps = step;
	if (last_start) *last_start = start;
}

static void scratch_free(scratch_t *s) { free(s->prev.a); free(s->curr.a); free(s->mem.a); free(s->sub.a); free(s->merged.a); }

/* ------------------------------------------------------------------ flat APIs */
void orc_extend(const orc_index_t *ix, const uint64_t ik3[3], int is_back, uint64_t ok12[12])
{
	iv_t ik, ok[4];
	int c;
	ik.x0 = ik3[0]; ik.x1 = ik3[1]; ik.x2 = ik3[2]; ik.info = 0;
	extend(ix, &ik, is_back, ok, 0);
	for (c = 0; c < 4; ++c) { ok12[3*c] = ok[c].x0; ok12[3*c+1] = ok[c].x1; ok12[3*c+2] = ok[c].x2; }
}

int64_t orc_smem1(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs, const int32_t *x,
                  const int32_t *min_intv, uint64_t *intv, int64_t cap, int64_t *read_off, int32_t *ret)
{
	scratch_t s;
	int64_t i, total = 0;
	int k;
	memset(&s, 0, sizeof(s));
	read_off[0] = 0;
	for (i = 0; i < n; ++i) {
		ret[i] = smem1(ix, (int)(offs[i+1] - offs[i]), seq + offs[i], x[i], min_intv[i], &s.mem, &s.prev, &s.curr, 0);
		for (k = 0; k < s.mem.n; ++k, ++total)
			if (total < cap) memcpy(intv + 4 * total, &s.mem.a[k], 32);
		read_off[i+1] = total;
	}
	scratch_free(&s);
	return total;
}

typedef struct { uint64_t *v; uint16_t *step; int64_t n, m; } flat_t;
static void flat_emit(void *ctx, const iv_t *e, int step)
{
	flat_t *f = (flat_t *)ctx;
	if (f->n == f->m) {
		f->m = f->m ? f->m * 2 : 1024;
		f->v = (uint64_t *)realloc(f->v, (size_t)f->m * 32);
		f->step = (uint16_t *)realloc(f->step, (size_t)f->m * 2);
	}
	memcpy(f->v + 4 * f->n, e, 32);
	f->step[f->n++] = (uint16_t)step;
}

typedef struct {
	const orc_index_t *ix; const uint8_t *seq; const int64_t *offs; orc_seed_opt_t opt;
	int64_t lo, hi; flat_t out; int64_t *cnt; int32_t *n_steps, *last_start; orc_stats_t st; int want_stats;
} job_t;

static void *collect_worker(void *data)
{
	job_t *j = (job_t *)data;
	scratch_t s;
	int64_t i;
	memset(&s, 0, sizeof(s));
	for (i = j->lo; i < j->hi; ++i) {
		int64_t n0 = j->out.n;
		collect_read(j->ix, (int)(j->offs[i+1] - j->offs[i]), j->seq + j->offs[i], &j->opt, &s, j->want_stats ? &j->st : 0,
		             flat_emit, &j->out, j->n_steps ? &j->n_steps[i] : 0, j->last_start ? &j->last_start[i] : 0);
		j->cnt[i - j->lo] = j->out.n - n0;
	}
	scratch_free(&s);
	return 0;
}

int64_t orc_collect(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs, const orc_seed_opt_t *opt,
                    int nthreads, uint64_t *intv, int64_t cap, int64_t *read_off, uint16_t *step,
                    int32_t *n_steps, int32_t *last_start, orc_stats_t *stats)
{
	job_t *jobs;
	pthread_t *tid;
	int t;
	int64_t total = 0, i, per, pos = 0;
	if (nthreads < 1) nthreads = 1;
	if (nthreads > n) nthreads = n > 0 ? (int)n : 1;
	jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
	tid = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	per = (n + nthreads - 1) / nthreads;
	for (t = 0; t < nthreads; ++t) {
		job_t *j = &jobs[t];
		j->ix = ix; j->seq = seq; j->offs = offs; j->opt = *opt; j->want_stats = stats != 0;
		j->lo = t * per < n ? t * per : n; j->hi = (t + 1) * per < n ? (t + 1) * per : n;
		j->cnt = (int64_t *)calloc((size_t)(j->hi - j->lo + 1), sizeof(int64_t));
		j->n_steps = n_steps; j->last_start = last_start;
		pthread_create(&tid[t], 0, collect_worker, j);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(tid[t], 0);
	read_off[0] = 0;
	if (stats) memset(stats, 0, sizeof(*stats));
	for (t = 0; t < nthreads; ++t) {
		for (i = jobs[t].lo; i < jobs[t].hi; ++i) { total += jobs[t].cnt[i - jobs[t].lo]; read_off[i+1] = total; }
		if (stats) {
			stats->extends += jobs[t].st.extends; stats->blocks += jobs[t].st.blocks;
			stats->smem1_calls += jobs[t].st.smem1_calls; stats->steps += jobs[t].st.steps;
			stats->intervals += jobs[t].st.intervals;
			if (jobs[t].st.max_curr > stats->max_curr) stats->max_curr = jobs[t].st.max_curr;
			if (jobs[t].st.max_mem > stats->max_mem) stats->max_mem = jobs[t].st.max_mem;
		}
	}
	if (total <= cap)
		for (t = 0; t < nthreads; ++t) {
			if (jobs[t].out.n) {
				memcpy(intv + 4 * pos, jobs[t].out.v, (size_t)jobs[t].out.n * 32);
				if (step) memcpy(step + pos, jobs[t].out.step, (size_t)jobs[t].out.n * 2);
			}
			pos += jobs[t].out.n;
		}
	for (t = 0; t < nthreads; ++t) { free(jobs[t].out.v); free(jobs[t].out.step); free(jobs[t].cnt); }
	free(jobs); free(tid);
	return total;
}

/* ------------------------------------------------------------------ seed -> reference position */
/* bwt_occ, bwt.c:125-147: occurrences of base c in BWT[0..k] */
uint64_t orc_occ(const orc_index_t *ix, uint64_t k, int c)
{
	uint64_t cnt[4];
	if (k == ix->seq_len) return ix->L2[c + 1] - ix->L2[c];
	if (k == (uint64_t)-1) return 0;
	orc_occ4(ix, k, cnt);
	return cnt[c];
}

/* bwt_invPsi, bwt.c:71-77 */
static uint64_t inv_psi(const orc_index_t *ix, uint64_t k)
{
	uint64_t x = k - (k > ix->primary);
	int c = (int)((ix->bwt[((x >> 7) << 4) + 8 + ((x & 127) >> 4)] >> ((~x & 15) << 1)) & 3);   /* bwt_B0, bwt.h:78 */
	uint64_t r = ix->L2[c] + orc_occ(ix, k, c);
	return k == ix->primary ? 0 : r;
}

/* bwt_sa, bwt.c:104-114 */
void orc_sa(const orc_index_t *ix, int sa_intv, const uint64_t *sa, int64_t n, const uint64_t *k, uint64_t *out)
{
	int64_t i;
	const uint64_t mask = (uint64_t)sa_intv - 1;
	for (i = 0; i < n; ++i) {
		uint64_t kk = k[i], steps = 0;
		while (kk & mask) { ++steps; kk = inv_psi(ix, kk); }
		out[i] = steps + sa[kk / (uint64_t)sa_intv];
	}
}

/* ------------------------------------------------------------------ checksum + timing arm */
#define FNV_BASIS 0xcbf29ce484222325ull
#define FNV_PRIME 0x100000001b3ull

uint64_t orc_checksum(int64_t n, const uint64_t *intv, const int64_t *read_off)
{
	uint64_t sum = 0;
	int64_t i, k;
	for (i = 0; i < n; ++i) {
		uint64_t h = FNV_BASIS ^ (uint64_t)i;
		for (k = 4 * read_off[i]; k < 4 * read_off[i+1]; ++k) h = (h ^ intv[k]) * FNV_PRIME;
		sum += h;
	}
	return sum;
}

typedef struct { uint64_t h; int64_t n; } hash_ctx_t;
static void hash_emit(void *ctx, const iv_t *e, int step)
{
	hash_ctx_t *c = (hash_ctx_t *)ctx;
	(void)step;
	c->h = (c->h ^ e->x0) * FNV_PRIME; c->h = (c->h ^ e->x1) * FNV_PRIME;
	c->h = (c->h ^ e->x2) * FNV_PRIME; c->h = (c->h ^ e->info) * FNV_PRIME;
	c->n++;
}

typedef struct {
	const orc_index_t *ix; const uint8_t *seq; const int64_t *offs; orc_seed_opt_t opt;
	int64_t n; volatile int64_t *next; uint64_t sum; int64_t cnt;
} tjob_t;

static void *time_worker(void *data)
{
	tjob_t *j = (tjob_t *)data;
	scratch_t s;
	memset(&s, 0, sizeof(s));
	for (;;) { /* dynamic chunks of 256 reads, like kt_for's stealing (kthread.c:17-38) */
		int64_t lo = __sync_fetch_and_add(j->next, 256), hi = lo + 256 < j->n ? lo + 256 : j->n, i;
		if (lo >= j->n) break;
		for (i = lo; i < hi; ++i) {
			hash_ctx_t c;
			c.h = FNV_BASIS ^ (uint64_t)i; c.n = 0;
			collect_read(j->ix, (int)(j->offs[i+1] - j->offs[i]), j->seq + j->offs[i], &j->opt, &s, 0, hash_emit, &c, 0, 0);
			j->sum += c.h; j->cnt += c.n;
		}
	}
	scratch_free(&s);
	return 0;
}

double orc_time_collect(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs,
                        const orc_seed_opt_t *opt, int nthreads, uint64_t *checksum, int64_t *n_intervals)
{
	tjob_t *jobs;
	pthread_t *tid;
	volatile int64_t next = 0;
	struct timespec t0, t1;
	int t;
	if (nthreads < 1) nthreads = 1;
	jobs = (tjob_t *)calloc((size_t)nthreads, sizeof(tjob_t));
	tid = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
	clock_gettime(CLOCK_MONOTONIC, &t0);
	for (t = 0; t < nthreads; ++t) {
		jobs[t].ix = ix; jobs[t].seq = seq; jobs[t].offs = offs; jobs[t].opt = *opt; jobs[t].n = n; jobs[t].next = &next;
		pthread_create(&tid[t], 0, time_worker, &jobs[t]);
	}
	for (t = 0; t < nthreads; ++t) pthread_join(tid[t], 0);
	clock_gettime(CLOCK_MONOTONIC, &t1);
	*checksum = 0; *n_intervals = 0;
	for (t = 0; t < nthreads; ++t) { *checksum += jobs[t].sum; *n_intervals += jobs[t].cnt; }
	free(jobs); free(tid);
	return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

/* ------------------------------------------------------------------ seeds and chains */
/* Intervals -> seeds: the seed loop of mem_insert_seed (bwamem.c:462-476): an interval of seed length >= min_seed_len with
 * x[2] <= max_occ yields its x[2] seeds {bwt_sa(x[0] + k), qbeg, len}, in interval order then k.  (The forward/reverse
 * boundary test of bwamem.c:478 belongs to chaining below.) */
int64_t orc_seeds(const orc_index_t *ix, int sa_intv, const uint64_t *sa, int64_t n, const uint64_t *intv, const int64_t *read_off,
                  int min_seed_len, int64_t max_occ, orc_seed_t *seeds, int64_t cap, int64_t *seed_off)
{
	int64_t r, e, tot = 0;
	seed_off[0] = 0;
	for (r = 0; r < n; ++r) {
		for (e = read_off[r]; e < read_off[r + 1]; ++e) {
			const uint64_t *p = intv + 4 * e;
			const int qbeg = (int)(p[3] >> 32), slen = (int)(uint32_t)p[3] - qbeg;
			uint64_t k;
			if (slen < min_seed_len || p[2] > (uint64_t)max_occ) continue;
			for (k = 0; k < p[2]; ++k, ++tot) {
				if (tot < cap) {
					uint64_t row = p[0] + k, pos;
					orc_sa(ix, sa_intv, sa, 1, &row, &pos);
					seeds[tot].rbeg = (int64_t)pos; seeds[tot].qbeg = qbeg; seeds[tot].len = slen;
				}
			}
		}
		seed_off[r + 1] = tot;
	}
	return tot;
}

typedef struct { int n, m; int64_t pos; orc_seed_t *seeds; } ochain_t;

/* test_and_merge, bwamem.c:334-356: 1 = seed absorbed (contained, or appended to the chain), 0 = start a new chain */
static int chain_absorbs(const orc_chain_opt_t *o, int64_t l_pac, ochain_t *c, const orc_seed_t *p)
{
	const orc_seed_t *first = &c->seeds[0], *last = &c->seeds[c->n - 1];
	const int64_t qend = last->qbeg + last->len, rend = last->rbeg + last->len;
	int64_t x, y;
	if (p->qbeg >= first->qbeg && p->qbeg + p->len <= qend && p->rbeg >= first->rbeg && p->rbeg + p->len <= rend) return 1;
	if ((last->rbeg < l_pac || first->rbeg < l_pac) && p->rbeg >= l_pac) return 0;      /* other strand */
	x = p->qbeg - last->qbeg;
	y = p->rbeg - last->rbeg;
	if (y >= 0 && x - y <= o->w && y - x <= o->w && x - last->len < o->max_chain_gap && y - last->len < o->max_chain_gap) {
		if (c->n == c->m) { c->m <<= 1; c->seeds = (orc_seed_t *)realloc(c->seeds, (size_t)c->m * sizeof(orc_seed_t)); }
		c->seeds[c->n++] = *p;
		return 1;
	}
	return 0;
}

/* mem_chain_weight, bwamem.c:502-521, including its second loop that keeps advancing `end` by QUERY coordinates */
static int chain_weight(const ochain_t *c)
{
	int64_t end = 0;
	int j, w = 0, tmp;
	for (j = 0; j < c->n; ++j) {
		const orc_seed_t *s = &c->seeds[j];
		if (s->qbeg >= end) w += s->len;
		else if (s->qbeg + s->len > end) w += (int)(s->qbeg + s->len - end);
		if (s->qbeg + s->len > end) end = s->qbeg + s->len;
	}
	tmp = w;
	for (j = 0, end = 0; j < c->n; ++j) {
		const orc_seed_t *s = &c->seeds[j];
		if (s->rbeg >= end) w += s->len;
		else if (s->rbeg + s->len > end) w += (int)(s->rbeg + s->len - end);
		if (s->qbeg + s->len > end) end = s->qbeg + s->len;
	}
	return w < tmp ? w : tmp;
}

typedef struct { int beg, end, w, p, p2; } flt_t;
#define FLT_LT(a, b) ((a).w > (b).w)      /* bwamem.c:626: heavier chains first */

static void flt_insertsort(flt_t *s, flt_t *t)   /* ksort.h __ks_insertsort: stable */
{
	flt_t *i, *j, tmp;
	for (i = s + 1; i < t; ++i)
		for (j = i; j > s && FLT_LT(*j, *(j - 1)); --j) { tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

static void flt_combsort(size_t n, flt_t *a)     /* ksort.h ks_combsort (introsort's depth-limit fallback) */
{
	const double shrink = 1.2473309501039786540366528676643;
	int swapped;
	size_t gap = n;
	flt_t tmp, *i, *j;
	do {
		if (gap > 2) { gap = (size_t)((double)gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
		swapped = 0;
		for (i = a; i < a + n - gap; ++i) {
			j = i + gap;
			if (FLT_LT(*j, *i)) { tmp = *i; *i = *j; *j = tmp; swapped = 1; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) flt_insertsort(a, a + n);
}

/* ks_introsort of ksort.h as instantiated at bwamem.c:627: the order of chains of EQUAL weight is whatever this exact
 * procedure leaves (median-of-3 quicksort down to runs of <= 16, depth limit 2*ceil(log2 n) with combsort, one final
 * insertion sort), so it is restated step by step. */
static void flt_introsort(size_t n, flt_t *a)
{
	typedef struct { flt_t *left, *right; int depth; } frame_t;
	frame_t stack[2 * 8 * sizeof(size_t) + 2], *top = stack;
	flt_t *s, *t, *i, *j, *k, rp, tmp;
	int d;
	if (n < 1) return;
	if (n == 2) { if (FLT_LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
	for (d = 2; (1ul << d) < n; ++d) {}
	s = a; t = a + (n - 1); d <<= 1;
	for (;;) {
		if (s < t) {
			if (--d == 0) { flt_combsort((size_t)(t - s) + 1, s); t = s; continue; }
			i = s; j = t; k = i + ((j - i) >> 1) + 1;
			if (FLT_LT(*k, *i)) { if (FLT_LT(*k, *j)) k = j; }
			else k = FLT_LT(*j, *i) ? i : j;
			rp = *k;
			if (k != t) { tmp = *k; *k = *t; *t = tmp; }
			for (;;) {
				do ++i; while (FLT_LT(*i, rp));
				do --j; while (i <= j && FLT_LT(rp, *j));
				if (j <= i) break;
				tmp = *i; *i = *j; *j = tmp;
			}
			tmp = *i; *i = *t; *t = tmp;
			if (i - s > t - i) {
				if (i - s > 16) { top->left = s; top->right = i - 1; top->depth = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { top->left = i + 1; top->right = t; top->depth = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == stack) { flt_insertsort(a, a + n); return; }
			--top; s = top->left; t = top->right; d = top->depth;
		}
	}
}

/* mem_chain_flt, bwamem.c:629-700, on chain indices: order[] receives the kept chains in output order; returns how many */
static int chain_filter(const orc_chain_opt_t *o, int n_chn, const ochain_t *chains, int *order)
{
	flt_t *a;
	char *keep;
	int i, j, n, n_out = 0;
	if (n_chn <= 1) { for (i = 0; i < n_chn; ++i) order[i] = i; return n_chn; }
	a = (flt_t *)malloc(sizeof(flt_t) * (size_t)n_chn);
	keep = (char *)calloc((size_t)n_chn, 1);
	for (i = 0; i < n_chn; ++i) {
		const ochain_t *c = &chains[i];
		a[i].beg = c->seeds[0].qbeg;
		a[i].end = c->seeds[c->n - 1].qbeg + c->seeds[c->n - 1].len;
		a[i].w = chain_weight(c); a[i].p = i; a[i].p2 = -1;
	}
	flt_introsort((size_t)n_chn, a);
	for (i = 0; i < n_chn; ++i) { order[i] = a[i].p; a[i].p = i; }          /* best chain first; p = rank from here on */
	for (i = 1, n = 1; i < n_chn; ++i) {
		for (j = 0; j < n; ++j) {
			const int b_max = a[j].beg > a[i].beg ? a[j].beg : a[i].beg;
			const int e_min = a[j].end < a[i].end ? a[j].end : a[i].end;
			if (e_min > b_max) {
				const int li = a[i].end - a[i].beg, lj = a[j].end - a[j].beg, min_l = li < lj ? li : lj;
				if (e_min - b_max >= min_l * o->mask_level) {                /* int vs float, as in the reference */
					if (a[j].p2 < 0) a[j].p2 = a[i].p;
					if (a[i].w < a[j].w * o->chain_drop_ratio && a[j].w - a[i].w >= o->min_seed_len << 1) break;
				}
			}
		}
		if (j == n) a[n++] = a[i];
	}
	for (i = 0; i < n; ++i) { keep[a[i].p] = 1; if (a[i].p2 >= 0) keep[a[i].p2] = 1; }
	for (i = 0; i < n_chn; ++i) if (keep[i]) order[n_out++] = order[i];
	free(a); free(keep);
	return n_out;
}

/* The chains of a read live in the reference in a klib B-tree keyed by pos (bwamem.c:328-331: KBTREE_INIT(chn, mem_chain_t,
 * chain_cmp); kb_init(chn, KB_DEFAULT_SIZE = 512) with 24-byte keys and 8-byte pointers gives t = ((512 - 4 - 8) / 32 + 1) >> 1
 * = 8, i.e. at most 15 keys per node, kbtree.h:55-70).  For chains with EQUAL pos -- the same reference segment twice in a
 * read, more than `w` apart -- which of them kb_intervalp returns as the lower bound, where kb_putp puts the next one and
 * the order __kb_traverse leaves them in all depend on the shape of that tree (which node a key was split into), so the
 * tree is restated here with the same node capacity, search, pre-emptive split and descent rules, over chain ids:
 *   bt_find   <- __kb_getp_aux (kbtree.h:117-131): lower-bound search in one node; index of the FIRST key equal to pos if there
 *                is one (*r = 0), else of the last smaller key (-1 if none)
 *   bt_lower  <- kb_intervalp's `lower` (kbtree.h:151-169): an equal key ends the descent at once; otherwise the deepest
 *                "last smaller key" seen on the way down
 *   bt_put    <- kb_putp / __kb_putp_aux / __kb_split (kbtree.h:174-226): a full root is split first; on the way down a full
 *                child is split before it is entered and the descent moves right only if pos is GREATER than the median
 *                moved up; in a leaf the key goes right behind the slot bt_find names
 *   bt_walk   <- __kb_traverse (kbtree.h:336-361): in order */
#define BT_T 8
#define BT_MAX (2 * BT_T - 1)
typedef struct { int n, internal; int key[BT_MAX]; int ptr[BT_MAX + 1]; } btnode_t;
typedef struct { btnode_t *nd; int n_nodes, cap, root; const ochain_t *ch; } btree_t;

static int bt_new(btree_t *t, int internal)
{
	if (t->n_nodes == t->cap) { t->cap = t->cap ? t->cap << 1 : 8; t->nd = (btnode_t *)realloc(t->nd, sizeof(btnode_t) * (size_t)t->cap); }
	memset(&t->nd[t->n_nodes], 0, sizeof(btnode_t));
	t->nd[t->n_nodes].internal = internal;
	return t->n_nodes++;
}

static int bt_find(const btree_t *t, const btnode_t *x, int64_t pos, int *r)
{
	int lo = 0, hi = x->n;
	if (x->n == 0) return -1;
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (t->ch[x->key[mid]].pos < pos) lo = mid + 1; else hi = mid; }
	if (lo == x->n) { *r = 1; return x->n - 1; }
	*r = t->ch[x->key[lo]].pos == pos ? 0 : -1;
	return *r < 0 ? lo - 1 : lo;
}

static int bt_lower(const btree_t *t, int64_t pos)        /* chain id or -1 */
{
	int x = t->root, lower = -1, r = 0;
	for (;;) {
		const btnode_t *nd = &t->nd[x];
		const int i = bt_find(t, nd, pos, &r);
		if (i >= 0 && r == 0) return nd->key[i];
		if (i >= 0) lower = nd->key[i];
		if (!nd->internal) return lower;
		x = nd->ptr[i + 1];
	}
}

static void bt_split(btree_t *t, int xi, int i, int yi)   /* child yi = ptr[i] of xi is full */
{
	const int zi = bt_new(t, t->nd[yi].internal);       /* (may move t->nd) */
	btnode_t *x = &t->nd[xi], *y = &t->nd[yi], *z = &t->nd[zi];
	z->n = BT_T - 1;
	memcpy(z->key, y->key + BT_T, sizeof(int) * (BT_T - 1));
	if (y->internal) memcpy(z->ptr, y->ptr + BT_T, sizeof(int) * BT_T);
	y->n = BT_T - 1;
	memmove(x->ptr + i + 2, x->ptr + i + 1, sizeof(int) * (size_t)(x->n - i));
	x->ptr[i + 1] = zi;
	memmove(x->key + i + 1, x->key + i, sizeof(int) * (size_t)(x->n - i));
	x->key[i] = y->key[BT_T - 1];
	++x->n;
}

static void bt_put(btree_t *t, int id)
{
	const int64_t pos = t->ch[id].pos;
	int x = t->root, r = 0;
	if (t->nd[x].n == BT_MAX) {
		const int s = bt_new(t, 1);
		t->nd[s].ptr[0] = x;
		bt_split(t, s, 0, x);
		t->root = x = s;
	}
	while (t->nd[x].internal) {
		int i = bt_find(t, &t->nd[x], pos, &r) + 1;
		if (t->nd[t->nd[x].ptr[i]].n == BT_MAX) {
			bt_split(t, x, i, t->nd[x].ptr[i]);
			if (pos > t->ch[t->nd[x].key[i]].pos) ++i;
		}
		x = t->nd[x].ptr[i];
	}
	{
		btnode_t *nd = &t->nd[x];
		const int i = bt_find(t, nd, pos, &r);
		if (i != nd->n - 1) memmove(nd->key + i + 2, nd->key + i + 1, sizeof(int) * (size_t)(nd->n - i - 1));
		nd->key[i + 1] = id;
		++nd->n;
	}
}

static int bt_walk(const btree_t *t, int x, int *out, int n)
{
	const btnode_t *nd = &t->nd[x];
	int i;
	for (i = 0; i <= nd->n; ++i) {
		if (nd->internal) n = bt_walk(t, nd->ptr[i], out, n);
		if (i < nd->n) out[n++] = nd->key[i];
	}
	return n;
}

/* mem_chain (bwamem.c:593-615) from the seed loop on (bwamem.c:478-496): per seed, the closest chain at or below its
 * reference position (kb_intervalp lower bound) absorbs it or a new chain keyed by rbeg is inserted; chains leave in
 * the tree's order (__kb_traverse).  Then, optionally, mem_chain_flt.  Flat result as oracle/ref_harness.c:ref_chains. */
int64_t orc_chains(int64_t n, const orc_seed_t *seeds, const int64_t *seed_off, int64_t l_pac, const orc_chain_opt_t *o, int do_flt,
                   int64_t *chain_off, int64_t *chain, int64_t chain_cap, orc_seed_t *out_seeds, int64_t seed_cap, int64_t *n_seeds_out)
{
	int64_t r, nc = 0, ns = 0;
	ochain_t *ch = 0, *sorted = 0;        /* ch: creation order (the tree's keys are indices into it); sorted: traversal order */
	int *order = 0, *walk = 0, cap = 0;
	btree_t t;
	memset(&t, 0, sizeof t);
	chain_off[0] = 0;
	for (r = 0; r < n; ++r) {
		int nch = 0, k, n_out;
		int64_t e;
		t.n_nodes = 0; t.root = bt_new(&t, 0);
		for (e = seed_off[r]; e < seed_off[r + 1]; ++e) {
			const orc_seed_t *s = &seeds[e];
			int lower;
			if (s->rbeg < l_pac && l_pac < s->rbeg + s->len) continue;      /* bridges the forward/reverse boundary, bwamem.c:478 */
			t.ch = ch;
			lower = nch ? bt_lower(&t, s->rbeg) : -1;
			if (lower >= 0 && chain_absorbs(o, l_pac, &ch[lower], s)) continue;
			if (nch == cap) {
				cap = cap ? cap << 1 : 16;
				ch = (ochain_t *)realloc(ch, sizeof(ochain_t) * (size_t)cap);
				sorted = (ochain_t *)realloc(sorted, sizeof(ochain_t) * (size_t)cap);
				order = (int *)realloc(order, sizeof(int) * (size_t)cap);
				walk = (int *)realloc(walk, sizeof(int) * (size_t)cap);
			}
			ch[nch].n = 1; ch[nch].m = 4; ch[nch].pos = s->rbeg;
			ch[nch].seeds = (orc_seed_t *)calloc(4, sizeof(orc_seed_t));
			ch[nch].seeds[0] = *s;
			t.ch = ch;
			bt_put(&t, nch);
			++nch;
		}
		if (nch) { k = bt_walk(&t, t.root, walk, 0); (void)k; }
		for (k = 0; k < nch; ++k) sorted[k] = ch[walk[k]];
		if (do_flt) n_out = chain_filter(o, nch, sorted, order);
		else for (n_out = 0; n_out < nch; ++n_out) order[n_out] = n_out;
		for (k = 0; k < n_out; ++k) {
			const ochain_t *c = &sorted[order[k]];
			int q;
			if (nc < chain_cap) { chain[2 * nc] = c->pos; chain[2 * nc + 1] = c->n; }
			++nc;
			for (q = 0; q < c->n; ++q, ++ns) if (ns < seed_cap) out_seeds[ns] = c->seeds[q];
		}
		for (k = 0; k < nch; ++k) free(ch[k].seeds);
		chain_off[r + 1] = nc;
	}
	free(ch); free(sorted); free(order); free(walk); free(t.nd);
	*n_seeds_out = ns;
	return nc;
}
