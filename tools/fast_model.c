/* TEST INFRASTRUCTURE ONLY -- executable specification of the "k-mer count pyramid" fast path of the CUDA kernel.
 *
 * The reference algorithm (bwt_smem1, bwt.c:776-835) spends most of its bwt_extend calls on SHORT patterns: the
 * forward sweep and the triangular backward sweep over patterns q[s..e) of length <= ~17 on a 3.1 Gbp index, whose
 * intervals are only ever looked at for their SIZE (the push / drop / stop decisions of bwt.c:794-799,813-824).
 * The size of the bi-interval of a pattern P is the number of occurrences of P in the indexed text T (forward +
 * reverse complement), its x[0] is 1 + the number of suffixes of T that sort before P, and its x[1] is the x[0] of
 * revcomp(P) -- functions of P alone, so they can be tabulated:
 *
 *   cnt[L][code]   L = 1..DL        occurrences of every L-mer                           (uint32, direct)
 *   cum[L][code]   L = 1..DL+1      x[0] of every L-mer (rows sorting before it, + 1)    (uint64, direct)
 *   pyr[code]      L = DL+4         occurrences of every (DL+4)-mer, saturating uint8; 64 consecutive entries share
 *                                   a (DL+1)-mer prefix, and the counts of levels DL+1..DL+3 are sums of 64/16/4 of them
 *   top[code]      L = DL+5         occurrences of every (DL+5)-mer, saturating uint8
 *   (255 = "unknown": saturated, or one of the few entries touched by a suffix of T shorter than the level)
 *
 * An entry of the prev/curr lists is VIRTUAL while its pattern is no longer than D = DL+5 (only its end and size are
 * kept), and REAL beyond (a bi-interval extended with the FM index exactly like the reference).  A virtual entry is
 * MATERIALISED (x[0], x[1] computed from cum/pyr/top) when it is emitted or when it outgrows the tables.
 * Whenever a table says "unknown" the whole read is handed to the plain FM path (return value < 0).
 *
 * This file mirrors that logic on the CPU so that tests can check it bit for bit against oracle/smem_oracle.c, and
 * counts the memory requests the device kernel will issue.  It links against liboracle.so for bwt_extend.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../oracle/smem_oracle.h"

#define DLMAX 13

typedef struct {
	int DL, has_top;
	const uint32_t *cnt[DLMAX + 1];
	const uint64_t *cum[DLMAX + 2];
	const uint8_t *pyr, *top;
} fm_tables_t;

typedef struct {
	uint64_t calls, rounds, fm_fwd, fm_bwd, mat_emit, mat_grow, lookups_direct, lookups_pyr, lookups_top, escapes, reads;
} fm_stats_t;

typedef struct { uint64_t x0, x1, s; int end, real; } ent_t;
typedef struct { uint64_t x0, x1, x2, info; } iv_t;
typedef struct { ent_t *a; int n, m; } entv_t;
typedef struct { iv_t *a; int n, m; } ivv_t;

#define ESC ((int64_t)-1)

static void entv_push(entv_t *v, ent_t e)
{
	if (v->n == v->m) { v->m = v->m ? v->m * 2 : 32; v->a = (ent_t *)realloc(v->a, sizeof(ent_t) * (size_t)v->m); }
	v->a[v->n++] = e;
}
static void ivv_push(ivv_t *v, iv_t e)
{
	if (v->n == v->m) { v->m = v->m ? v->m * 2 : 32; v->a = (iv_t *)realloc(v->a, sizeof(iv_t) * (size_t)v->m); }
	v->a[v->n++] = e;
}

static uint64_t code_of(const uint8_t *q, int s, int e) { uint64_t c = 0; int i; for (i = s; i < e; ++i) c = (c << 2) | q[i]; return c; }
static uint64_t rc_code_of(const uint8_t *q, int s, int e) { uint64_t c = 0; int i; for (i = e - 1; i >= s; --i) c = (c << 2) | (uint64_t)(3 - q[i]); return c; }

/* occurrences of the L-mer `code`, or ESC */
static int64_t tab_cnt(const fm_tables_t *t, int L, uint64_t code, fm_stats_t *st)
{
	const int K = t->DL + 1, LP = t->DL + 4;
	if (L <= t->DL) { st->lookups_direct++; return (int64_t)t->cnt[L][code]; }
	if (L <= LP) {
		const uint64_t lo = code << (2 * (LP - L)), span = 1ull << (2 * (LP - L));
		uint64_t k, sum = 0;
		(void)K;
		for (k = 0; k < span; ++k) { uint8_t v = t->pyr[lo + k]; if (v == 255) return ESC; sum += v; }
		return (int64_t)sum;
	}
	{ uint8_t v = t->top[code]; st->lookups_top++; return v == 255 ? ESC : (int64_t)v; }
}

/* x[0] of the L-mer `code` (1 + number of suffixes sorting before it), or ESC */
static int64_t tab_x0(const fm_tables_t *t, int L, uint64_t code)
{
	const int K = t->DL + 1, LP = t->DL + 4;
	if (L <= K) return (int64_t)t->cum[L][code];
	{
		const int Lb = L < LP ? L : LP;                                  /* level inside the pyramid block */
		const uint64_t c16 = L <= LP ? code : code >> (2 * (L - LP));
		const uint64_t pre = c16 >> (2 * (Lb - K)), blk = pre << (2 * (LP - K));
		const uint64_t lo = c16 << (2 * (LP - Lb));
		uint64_t k, sum = t->cum[K][pre];
		for (k = blk; k < lo; ++k) { uint8_t v = t->pyr[k]; if (v == 255) return ESC; sum += v; }
		/* the pattern's own range must be trustworthy too: a suffix of T shorter than L that is a prefix of the pattern sorts before it */
		for (k = lo; k < lo + (1ull << (2 * (LP - Lb))); ++k) if (t->pyr[k] == 255) return ESC;
		if (L > LP) {                                                    /* smaller siblings at the top level */
			uint64_t sib = code & ~3ull;
			/* the block prefix must be trustworthy as a whole: a suffix of T of length LP equal to c16 sorts first */
			for (k = sib; k < code; ++k) { uint8_t v = t->top[k]; if (v == 255) return ESC; sum += v; }
			if (t->top[code] == 255) return ESC;
		}
		return (int64_t)sum;
	}
}

/* bi-interval of q[s..e) from the tables; 0 on success */
static int materialise(const fm_tables_t *t, const uint8_t *q, int s, int e, ent_t *p)
{
	const int L = e - s;
	int64_t a = tab_x0(t, L, code_of(q, s, e)), b = tab_x0(t, L, rc_code_of(q, s, e));
	if (a == ESC || b == ESC) return -1;
	p->x0 = (uint64_t)a; p->x1 = (uint64_t)b; p->real = 1;
	return 0;
}

typedef struct { entv_t prev, curr; ivv_t mem, sub, merged; } scr_t;

/* bwt_smem1 (bwt.c:776-835) with virtual entries.  Returns ret, or -1 when a table answered "unknown". */
static int fast_smem1(const orc_index_t *ix, const fm_tables_t *t, int len, const uint8_t *q, int x, int min_intv, ivv_t *mem,
                      entv_t *prev, entv_t *curr, fm_stats_t *st)
{
	const int D = t->DL + (t->has_top ? 5 : 4);
	ent_t ik;
	entv_t *tmp;
	int i, j, ret;
	uint64_t ok12[12], ik3[3];
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	st->calls++;
	memset(&ik, 0, sizeof(ik));
	ik.s = t->cnt[1][q[x]]; ik.end = x + 1; ik.real = 0;
	curr->n = 0;
	for (i = x + 1; i < len; ++i) {
		uint64_t oks;
		ent_t ok;
		const int L = i + 1 - x;
		if (q[i] > 3) { entv_push(curr, ik); break; }
		memset(&ok, 0, sizeof(ok));
		if (L <= D) {
			int64_t c = tab_cnt(t, L, code_of(q, x, i + 1), st);
			if (c == ESC) return -1;
			oks = (uint64_t)c; ok.s = oks; ok.real = 0;
		} else {
			if (!ik.real) { if (materialise(t, q, x, i, &ik)) return -1; st->mat_grow++; }
			ik3[0] = ik.x0; ik3[1] = ik.x1; ik3[2] = ik.s;
			orc_extend(ix, ik3, 0, ok12);
			st->fm_fwd++;
			ok.x0 = ok12[3 * (3 - q[i])]; ok.x1 = ok12[3 * (3 - q[i]) + 1]; ok.s = ok12[3 * (3 - q[i]) + 2]; ok.real = 1;
			oks = ok.s;
		}
		if (oks != ik.s) {
			entv_push(curr, ik);
			if (oks < (uint64_t)min_intv) break;
		}
		ik = ok; ik.end = i + 1;
	}
	if (i == len) entv_push(curr, ik);
	for (i = 0, j = curr->n - 1; i < j; ++i, --j) { ent_t e = curr->a[i]; curr->a[i] = curr->a[j]; curr->a[j] = e; }
	ret = curr->a[0].end;
	tmp = prev; prev = curr; curr = tmp;
	for (i = x - 1; i >= -1; --i) {
		const int c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		curr->n = 0;
		st->rounds++;
		for (j = 0; j < prev->n; ++j) {
			ent_t *p = &prev->a[j], ok;
			uint64_t n = 0;
			memset(&ok, 0, sizeof(ok));
			if (c >= 0) {
				const int L = p->end - i;
				if (L <= D) {
					int64_t v = tab_cnt(t, L, code_of(q, i, p->end), st);
					if (v == ESC) return -1;
					n = (uint64_t)v; ok.s = n; ok.real = 0;
				} else {
					if (!p->real) { if (materialise(t, q, i + 1, p->end, p)) return -1; st->mat_grow++; }
					ik3[0] = p->x0; ik3[1] = p->x1; ik3[2] = p->s;
					orc_extend(ix, ik3, 1, ok12);
					st->fm_bwd++;
					ok.x0 = ok12[3 * c]; ok.x1 = ok12[3 * c + 1]; ok.s = ok12[3 * c + 2]; ok.real = 1;
					n = ok.s;
				}
			}
			if (c < 0 || n < (uint64_t)min_intv) {
				if (curr->n == 0 && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) {
					iv_t e;
					if (!p->real) { if (materialise(t, q, i + 1, p->end, p)) return -1; st->mat_emit++; }
					e.x0 = p->x0; e.x1 = p->x1; e.x2 = p->s; e.info = (uint64_t)p->end | ((uint64_t)(i + 1) << 32);
					ivv_push(mem, e);
				}
			} else if (curr->n == 0 || n != curr->a[curr->n - 1].s) {
				ok.end = p->end;
				entv_push(curr, ok);
			}
		}
		if (curr->n == 0) break;
		tmp = prev; prev = curr; curr = tmp;
	}
	for (i = 0, j = mem->n - 1; i < j; ++i, --j) { iv_t e = mem->a[i]; mem->a[i] = mem->a[j]; mem->a[j] = e; }
	return ret;
}


/* ------------------------------------------------------------------------------------------------------------------
 * The same algorithm in the form the device kernel runs it: the virtual entries of one bwt_smem1 call are a BIT MASK
 * over their ends (bit b <-> end x+1+b; at most D of them, all within D of x), the real entries a short list.
 * Everything a round needs from the tables is the vector n[L] = occurrences of the length-L prefix of the window
 * starting at the round's position.  Sizes are monotone in L, which is what makes the mask form possible:
 *   - the entries that fail (n < min_intv) are the longest ones, i.e. the top bits;
 *   - "differs from the previously pushed size" (bwt.c:822) is "some length in between changed the count".           */
typedef struct { uint64_t x0, x1, s; int end; } real_t;
typedef struct { real_t *a; int n, m; } realv_t;
static void realv_push(realv_t *v, real_t e)
{
	if (v->n == v->m) { v->m = v->m ? v->m * 2 : 16; v->a = (real_t *)realloc(v->a, sizeof(real_t) * (size_t)v->m); }
	v->a[v->n++] = e;
}

/* n[L], L = lo..hi, of the window q[pos..); returns -1 on "unknown" */
static int window_sizes(const fm_tables_t *t, const uint8_t *q, int pos, int lo, int hi, uint64_t *n, fm_stats_t *st)
{
	int L;
	for (L = lo; L <= hi; ++L) {
		int64_t v = tab_cnt(t, L, code_of(q, pos, pos + L), st);
		if (v == ESC) return -1;
		n[L] = (uint64_t)v;
	}
	st->lookups_pyr++;
	return 0;
}

static int real_from_tables(const fm_tables_t *t, const uint8_t *q, int s, int e, uint64_t size, real_t *p)
{
	ent_t tmp;
	if (materialise(t, q, s, e, &tmp)) return -1;
	p->x0 = tmp.x0; p->x1 = tmp.x1; p->s = size; p->end = e;
	return 0;
}

static int fast_smem1_mask(const orc_index_t *ix, const fm_tables_t *t, int len, const uint8_t *q, int x, int min_intv, ivv_t *mem,
                           realv_t *prevR, realv_t *currR, fm_stats_t *st)
{
	const int D = t->DL + (t->has_top ? 5 : 4);
	uint64_t n[DLMAX + 8], ok12[12], ik3[3], s_top = 0, m;
	uint32_t E = 0;
	int avail, Lend, L, stop = 0, ret, r, j;
	realv_t *tmp;
	mem->n = 0;
	if (q[x] > 3) return x + 1;
	if (min_intv < 1) min_intv = 1;
	m = (uint64_t)min_intv;
	st->calls++;
	for (avail = 0; avail <= D && x + avail < len && q[x + avail] <= 3; ++avail) {}     /* valid bases from x, capped at D + 1 */
	Lend = avail < D ? avail : D;
	if (window_sizes(t, q, x, 1, Lend, n, st)) return -1;
	/* forward sweep, bwt.c:790-806 */
	for (L = 1; L < Lend; ++L)
		if (n[L + 1] != n[L]) { E |= 1u << (L - 1); if (n[L + 1] < m) { stop = 1; break; } }
	currR->n = 0;
	if (!stop) {
		if (avail <= D) E |= 1u << (Lend - 1);               /* ambiguous base or end of read: the last interval is pushed */
		else {                                               /* the pattern outgrows the tables: FM from length D on */
			real_t ik;
			int i;
			if (real_from_tables(t, q, x, x + D, n[D], &ik)) return -1;
			st->mat_grow++;
			for (i = x + D; i < len; ++i) {
				real_t ok;
				if (q[i] > 3) { realv_push(currR, ik); break; }
				ik3[0] = ik.x0; ik3[1] = ik.x1; ik3[2] = ik.s;
				orc_extend(ix, ik3, 0, ok12);
				st->fm_fwd++;
				ok.x0 = ok12[3 * (3 - q[i])]; ok.x1 = ok12[3 * (3 - q[i]) + 1]; ok.s = ok12[3 * (3 - q[i]) + 2]; ok.end = i + 1;
				if (ok.s != ik.s) { realv_push(currR, ik); if (ok.s < m) break; }
				ik = ok;
			}
			if (i == len) realv_push(currR, ik);
		}
	}
	/* longest first */
	for (r = 0, j = currR->n - 1; r < j; ++r, --j) { real_t e = currR->a[r]; currR->a[r] = currR->a[j]; currR->a[j] = e; }
	tmp = prevR; prevR = currR; currR = tmp;
	if (E) { int b = 31 - __builtin_clz(E); s_top = n[b + 1]; }
	ret = prevR->n ? prevR->a[0].end : x + 1 + (31 - __builtin_clz(E));
	/* backward sweep, bwt.c:810-828 */
	for (r = 1; ; ++r) {
		const int i = x - r, c = i < 0 ? -1 : (q[i] < 4 ? q[i] : -1);
		uint64_t last_s = 0;
		int pushed = 0;                                      /* curr->n != 0 */
		uint32_t E2 = 0;
		currR->n = 0;
		st->rounds++;
#define TRY_EMIT(X0, X1, S, END, VIRT) do { \
		if (!pushed && (mem->n == 0 || (uint64_t)(i + 1) < (mem->a[mem->n - 1].info >> 32))) { \
			iv_t e_; e_.x0 = (X0); e_.x1 = (X1); e_.x2 = (S); e_.info = (uint64_t)(END) | ((uint64_t)(i + 1) << 32); \
			if (VIRT) { ent_t t_; if (materialise(t, q, i + 1, (END), &t_)) return -1; e_.x0 = t_.x0; e_.x1 = t_.x1; st->mat_emit++; } \
			ivv_push(mem, e_); } } while (0)
		for (j = 0; j < prevR->n; ++j) {
			real_t *p = &prevR->a[j], ok;
			ok.s = 0;
			if (c >= 0) {
				ik3[0] = p->x0; ik3[1] = p->x1; ik3[2] = p->s;
				orc_extend(ix, ik3, 1, ok12);
				st->fm_bwd++;
				ok.x0 = ok12[3 * c]; ok.x1 = ok12[3 * c + 1]; ok.s = ok12[3 * c + 2]; ok.end = p->end;
			}
			if (c < 0 || ok.s < m) TRY_EMIT(p->x0, p->x1, p->s, p->end, 0);
			else if (!pushed || ok.s != last_s) { realv_push(currR, ok); pushed = 1; last_s = ok.s; }
		}
		if (c < 0) {
			if (E) TRY_EMIT(0, 0, s_top, x + 1 + (31 - __builtin_clz(E)), 1);
			break;                                           /* nothing can be pushed: curr is empty */
		}
		{
			const int bg = D - r;                            /* the virtual entry that reaches length D + 1 in this round */
			if (bg >= 0 && (E >> bg & 1)) {
				real_t p, ok;
				if (real_from_tables(t, q, i + 1, x + 1 + bg, s_top, &p)) return -1;
				st->mat_grow++;
				ik3[0] = p.x0; ik3[1] = p.x1; ik3[2] = p.s;
				orc_extend(ix, ik3, 1, ok12);
				st->fm_bwd++;
				ok.x0 = ok12[3 * c]; ok.x1 = ok12[3 * c + 1]; ok.s = ok12[3 * c + 2]; ok.end = p.end;
				if (ok.s < m) TRY_EMIT(p.x0, p.x1, p.s, p.end, 0);
				else if (!pushed || ok.s != last_s) { realv_push(currR, ok); pushed = 1; last_s = ok.s; }
				E &= ~(1u << bg);
				if (E) s_top = 0;                            /* (the next top's previous size is fetched below when needed) */
			}
		}
		if (E) {
			const int btop = 31 - __builtin_clz(E), blow = __builtin_ctz(E);
			int b, first = 1;
			uint64_t prev_top_size = s_top;
			if (window_sizes(t, q, i, blow + r + 1, btop + r + 1, n, st)) return -1;
			if (prev_top_size == 0) {                        /* top changed by the grow step: its size of the previous round */
				int64_t v = tab_cnt(t, btop + r, code_of(q, i + 1, x + 1 + btop), st);
				if (v == ESC) return -1;
				prev_top_size = (uint64_t)v;
			}
			for (b = btop; b >= 0; --b) {
				if (!(E >> b & 1)) continue;
				L = b + r + 1;
				if (n[L] < m) { if (first) TRY_EMIT(0, 0, prev_top_size, x + 1 + b, 1); }
				else if (!pushed || n[L] != last_s) { E2 |= 1u << b; if (!(E2 & (E2 - 1))) s_top = n[L]; pushed = 1; last_s = n[L]; }
				first = 0;
			}
		}
		E = E2;
		if (currR->n == 0 && E == 0) break;
		tmp = prevR; prevR = currR; currR = tmp;
	}
#undef TRY_EMIT
	for (r = 0, j = mem->n - 1; r < j; ++r, --j) { iv_t e = mem->a[r]; mem->a[r] = mem->a[j]; mem->a[j] = e; }
	return ret;
}

int fast_model_variant = 1;      /* 0: entry form, 1: mask form */
static realv_t g_r1, g_r2;

static inline int iv_len(const iv_t *p) { return (int)((uint32_t)p->info - (uint32_t)(p->info >> 32)); }

/* smem_next2 (bwamem.c:244-305); returns 0 when exhausted, 1 with a list, -1 on a table escape */
static int fast_next2(const orc_index_t *ix, const fm_tables_t *t, int len, const uint8_t *q, int *start, int split_len, int split_width,
                      int start_width, scr_t *s, fm_stats_t *st)
{
	int i, max = 0, max_i = 0, ori_start, r;
	s->mem.n = s->sub.n = 0;
	if (*start >= len || *start < 0) return 0;
	while (*start < len && q[*start] > 3) ++*start;
	if (*start == len) return 0;
	ori_start = *start;
	r = fast_model_variant ? fast_smem1_mask(ix, t, len, q, ori_start, start_width, &s->mem, &g_r1, &g_r2, st)
	                       : fast_smem1(ix, t, len, q, ori_start, start_width, &s->mem, &s->prev, &s->curr, st);
	if (r < 0) return -1;
	*start = r;
	if (s->mem.n == 0) return 1;
	for (i = 0; i < s->mem.n; ++i) { int l = iv_len(&s->mem.a[i]); if (max < l) { max = l; max_i = i; } }
	if (split_len > 0 && max >= split_len && s->mem.a[max_i].x2 <= (uint64_t)split_width) {
		const iv_t *p = &s->mem.a[max_i];
		int j, mid = (int)(((uint32_t)p->info + (uint32_t)(p->info >> 32)) >> 1);
		ivv_t *a = &s->merged;
		if ((fast_model_variant ? fast_smem1_mask(ix, t, len, q, mid, (int)(p->x2 + 1), &s->sub, &g_r1, &g_r2, st)
		                        : fast_smem1(ix, t, len, q, mid, (int)(p->x2 + 1), &s->sub, &s->prev, &s->curr, st)) < 0) return -1;
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

/* Whole batch, single thread.  escaped[i] = 1 when read i needs the plain FM path (its intervals are then absent).
 * Same flat output conventions as orc_collect. */
int64_t fast_collect(const orc_index_t *ix, const fm_tables_t *t, int64_t n, const uint8_t *seq, const int64_t *offs,
                     const orc_seed_opt_t *o, uint64_t *intv, int64_t cap, int64_t *read_off, uint16_t *step, uint8_t *escaped,
                     fm_stats_t *st)
{
	scr_t s;
	int64_t r, total = 0;
	memset(&s, 0, sizeof(s));
	memset(st, 0, sizeof(*st));
	read_off[0] = 0;
	for (r = 0; r < n; ++r) {
		const int len = (int)(offs[r + 1] - offs[r]);
		const uint8_t *q = seq + offs[r];
		int split_len = (int)(o->min_seed_len * o->split_factor + .499), start = 0, stp = 0, rc, k;
		const int64_t total0 = total;
		if (split_len > len) split_len = len;
		escaped[r] = 0;
		st->reads++;
		while ((rc = fast_next2(ix, t, len, q, &start, split_len, o->split_width, o->start_width, &s, st)) > 0) {
			for (k = 0; k < s.mem.n; ++k, ++total)
				if (total < cap) { memcpy(intv + 4 * total, &s.mem.a[k], 32); step[total] = (uint16_t)stp; }
			++stp;
		}
		if (rc < 0) { escaped[r] = 1; st->escapes++; total = total0; }
		read_off[r + 1] = total;
	}
	free(s.prev.a); free(s.curr.a); free(s.mem.a); free(s.sub.a); free(s.merged.a);
	return total;
}
