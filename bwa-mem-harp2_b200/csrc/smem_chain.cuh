// smem_chain.cuh -- seeds -> chains on the device (sm_100a).
//
// What is computed (bit-exact with the reference's CPU code, see oracle/smem_oracle.c:orc_chains):
//   chain_build_kernel  <- the seed loop of mem_insert_seed from the boundary test on (bwamem.c:478-496):
//                          closest chain at or below the seed's reference position (kb_intervalp lower bound on the
//                          kbtree keyed by chain pos), test_and_merge (bwamem.c:334-356), else a new chain;
//                          then mem_chain_flt (bwamem.c:629-700): mem_chain_weight (bwamem.c:502-521), ks_introsort
//                          by weight (ksort.h, restated step by step because it fixes the order of equal weights),
//                          the overlap / drop-ratio filter, the kept "shadowed" chain p2.
//   chain_emit_kernel   <- __kb_traverse (bwamem.c:608-610) / the squeeze at bwamem.c:690-697: chains in output order
//                          with their seeds in insertion order.
//
// Shape of the work: ~2 seeds and ~1 chain per read on unique sequence, thousands of seeds for a read inside a repeat
// family (max_occ = 10000, bwamem.c:63).  One WARP per read: lane 0 runs the insertion loop (it is sequential by
// definition: every seed sees the chains its predecessors made), all lanes do the parts that grow with the number of
// chains -- the ordered insert (shift of the sorted chain list), the chain weights, the permutation after the sort and
// the copy-out -- so that a repeat read costs O(n^2 / 32) instead of stalling the grid behind one thread.
// A chain's id is the index of its first seed (pos = that seed's rbeg); the seeds of a chain are a linked list through s_next in
// insertion order.  While a read's chains are being built they live in a B-tree of chain ids with the node capacity, search,
// pre-emptive split and descent rules of the reference's klib kbtree (bwamem.c:328-331 / kbtree.h: t = 8, at most 15 keys per
// node): for chains with EQUAL pos -- the same reference segment twice in a read, more than w apart -- which one is found as the
// lower bound, where the next one goes and the order they leave in depend on the shape of that tree, so a sorted array (round 1)
// was not enough (oracle/smem_oracle.c bt_*; tests/test_chains.py crafted equal-key reads against the reference's mem_chain).
// Lane 0 runs the insertion loop -- it is sequential by definition, and O(log n) per seed now; the tree is walked in order into
// `ord`, from where the filter and the copy-out use all lanes.
// Provenance note: flt_introsort / flt_combsort / flt_insertsort and the two predicates chain_weight_dev / the test_and_merge test are
// deliberate step-by-step restatements of klib ksort.h:146-224 and bwamem.c:334-356, 502-521 -- an unstable sort's order of equal
// weights (and the reference's never-reset `w` in the weight loop) cannot be reproduced any other way.  They are confined to those
// ~70 lines; everything around them (one warp per read, the device B-tree, the scans, the emit) is this library's own design.
#pragma once
#include "smem_kernels.cuh"

struct ChainOpt {            // mem_opt_t fields (bwamem.h:33-60)
	int w, max_chain_gap, min_seed_len;
	float mask_level, chain_drop_ratio;
	long long l_pac;
	int do_flt;
};

struct __align__(8) Chain { long long pos; long long seed_first; int n_seeds, weight; };   // == smem_chain_t

struct FltRec { int beg, end, w, p, p2; };

struct ChainWork {
	const Seed *seeds;           // resident seeds of the last smem_gpu_seeds run, grouped by read
	const long long *off;        // [n + 1] interval CSR of the reads (d_off)
	const long long *soff;       // [total + 1] seed offset of every interval (d_soff): read r owns seeds soff[off[r]] .. soff[off[r+1]]
	long long n;
	int *ord, *ord2;             // per seed slot: chain ids in tree order / scratch for the permutation
	int *bt;                     // B-tree nodes, BT_INTS ints each; read r owns nodes soff[off[r]] / 7 + 2 r .. (at most n_seeds / 7 + 2 of them)
	int *c_last, *c_n, *s_next;  // per chain (at its first seed's slot): last seed, seed count; per seed: next seed of its chain
	FltRec *flt;
	unsigned char *keep;
	int *n_chains, *n_kept;      // [n + 1] per read: chains / seeds in the output (entry n = 0 for the scan total)
};

#define FLT_LT(a, b) ((a).w > (b).w)      // bwamem.c:626

__device__ __forceinline__ Seed ld_seed(const Seed *p)
{
	const int4 v = *reinterpret_cast<const int4 *>(p);
	Seed s;
	s.rbeg = (long long)(((unsigned long long)(unsigned)v.y << 32) | (unsigned)v.x); s.qbeg = v.z; s.len = v.w;
	return s;
}

// mem_chain_weight, bwamem.c:502-521 (its second loop advances `end` by QUERY coordinates; kept as is)
__device__ __forceinline__ int chain_weight_dev(const Seed *sd, const int *s_next, int first)
{
	long long end = 0;
	int w = 0;
	for (int e = first; e >= 0; e = s_next[e]) {
		const Seed s = ld_seed(&sd[e]);
		if (s.qbeg >= end) w += s.len;
		else if (s.qbeg + s.len > end) w += (int)(s.qbeg + s.len - end);
		if (s.qbeg + s.len > end) end = s.qbeg + s.len;
	}
	const int tmp = w;
	end = 0;
	for (int e = first; e >= 0; e = s_next[e]) {
		const Seed s = ld_seed(&sd[e]);
		if (s.rbeg >= end) w += s.len;
		else if (s.rbeg + s.len > end) w += (int)(s.rbeg + s.len - end);
		if (s.qbeg + s.len > end) end = s.qbeg + s.len;
	}
	return w < tmp ? w : tmp;
}

__device__ __forceinline__ void flt_insertsort(FltRec *s, FltRec *t)      // ksort.h __ks_insertsort
{
	for (FltRec *i = s + 1; i < t; ++i)
		for (FltRec *j = i; j > s && FLT_LT(*j, *(j - 1)); --j) { const FltRec tmp = *j; *j = *(j - 1); *(j - 1) = tmp; }
}

__device__ __noinline__ void flt_combsort(size_t n, FltRec *a)            // ksort.h ks_combsort
{
	const double shrink = 1.2473309501039786540366528676643;
	bool swapped;
	size_t gap = n;
	do {
		if (gap > 2) { gap = (size_t)((double)gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
		swapped = false;
		for (FltRec *i = a; i < a + n - gap; ++i) {
			FltRec *j = i + gap;
			if (FLT_LT(*j, *i)) { const FltRec tmp = *i; *i = *j; *j = tmp; swapped = true; }
		}
	} while (swapped || gap > 2);
	if (gap != 1) flt_insertsort(a, a + n);
}

// ks_introsort (ksort.h) as instantiated at bwamem.c:627, one thread
__device__ __noinline__ void flt_introsort(size_t n, FltRec *a)
{
	struct Frame { FltRec *left, *right; int depth; };
	Frame stack[72];
	Frame *top = stack;
	FltRec *s, *t, *i, *j, *k, rp, tmp;
	int d;
	if (n < 1) return;
	if (n == 2) { if (FLT_LT(a[1], a[0])) { tmp = a[0]; a[0] = a[1]; a[1] = tmp; } return; }
	for (d = 2; (1ull << d) < n; ++d) {}
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

// ---- the reference's kbtree over chain ids (see the header; restated from kbtree.h:117-226, 336-361) ----
#define BT_T 8
#define BT_MAX (2 * BT_T - 1)
#define BT_INTS 32           // [0] number of keys | internal << 8, [1..15] keys, [16..31] children
#define BT_LEAF 0
#define BT_INTERNAL 256
struct BTree { int *nd; int n_nodes, root; const Seed *sd; };

__device__ __forceinline__ int *bt_node(const BTree &t, int x) { return t.nd + (size_t)x * BT_INTS; }
__device__ __forceinline__ int bt_new(BTree &t, int kind)
{
	bt_node(t, t.n_nodes)[0] = kind;
	return t.n_nodes++;
}
// __kb_getp_aux: index of the FIRST key equal to pos (r = 0), else of the last smaller key (-1 if none)
__device__ __forceinline__ int bt_find(const BTree &t, const int *nd, long long pos, int &r)
{
	const int n = nd[0] & 255;
	int lo = 0, hi = n;
	if (n == 0) return -1;
	while (lo < hi) { const int mid = (lo + hi) >> 1; if (t.sd[nd[1 + mid]].rbeg < pos) lo = mid + 1; else hi = mid; }
	if (lo == n) { r = 1; return n - 1; }
	r = t.sd[nd[1 + lo]].rbeg == pos ? 0 : -1;
	return r < 0 ? lo - 1 : lo;
}
// kb_intervalp's lower bound: an equal key ends the descent at once, otherwise the deepest "last smaller key" on the way down
__device__ __forceinline__ int bt_lower(const BTree &t, long long pos)
{
	int x = t.root, lower = -1, r = 0;
	for (;;) {
		const int *nd = bt_node(t, x);
		const int i = bt_find(t, nd, pos, r);
		if (i >= 0 && r == 0) return nd[1 + i];
		if (i >= 0) lower = nd[1 + i];
		if (!(nd[0] & BT_INTERNAL)) return lower;
		x = nd[16 + i + 1];
	}
}
// __kb_split: child y = children[i] of x is full; its upper half moves to a new node z, its median up into x
__device__ __forceinline__ void bt_split(BTree &t, int xi, int i, int yi)
{
	int *y = bt_node(t, yi);
	const int kind = y[0] & BT_INTERNAL;
	const int zi = bt_new(t, kind);
	int *x = bt_node(t, xi), *z = bt_node(t, zi);
	z[0] = kind | (BT_T - 1);
	for (int k = 0; k < BT_T - 1; ++k) z[1 + k] = y[1 + BT_T + k];
	if (kind) for (int k = 0; k < BT_T; ++k) z[16 + k] = y[16 + BT_T + k];
	y[0] = kind | (BT_T - 1);
	const int xn = x[0] & 255;
	for (int k = xn; k > i; --k) x[16 + k + 1] = x[16 + k];
	x[16 + i + 1] = zi;
	for (int k = xn - 1; k >= i; --k) x[1 + k + 1] = x[1 + k];
	x[1 + i] = y[1 + BT_T - 1];
	x[0] = BT_INTERNAL | (xn + 1);
}
// kb_putp / __kb_putp_aux: a full root is split first; a full child is split before it is entered and the descent moves
// right only if pos is GREATER than the median moved up; in a leaf the key goes right behind the slot bt_find names
__device__ __forceinline__ void bt_put(BTree &t, int id)
{
	const long long pos = t.sd[id].rbeg;
	int x = t.root, r = 0;
	if ((bt_node(t, x)[0] & 255) == BT_MAX) {
		const int s = bt_new(t, BT_INTERNAL);
		bt_node(t, s)[16] = x;
		bt_split(t, s, 0, x);
		t.root = x = s;
	}
	for (;;) {
		int *nd = bt_node(t, x);
		if (!(nd[0] & BT_INTERNAL)) break;
		int i = bt_find(t, nd, pos, r) + 1;
		if ((bt_node(t, nd[16 + i])[0] & 255) == BT_MAX) {
			bt_split(t, x, i, nd[16 + i]);
			if (pos > t.sd[nd[1 + i]].rbeg) ++i;
		}
		x = nd[16 + i];
	}
	int *nd = bt_node(t, x);
	const int n = nd[0] & 255, i = bt_find(t, nd, pos, r);
	for (int k = n - 1; k > i; --k) nd[1 + k + 1] = nd[1 + k];
	nd[1 + i + 1] = id;
	nd[0] = n + 1;                                   // (a leaf)
}
// __kb_traverse: in order, iteratively (depth <= 11 for 2^31 keys at >= 8 children per internal node).
// A frame holds 2 * i + phase: phase 0 = child i is still to be visited, 1 = child i is done and key i comes next.
__device__ __forceinline__ int bt_walk(const BTree &t, int *out)
{
	int sx[12], si[12], top = 0, n = 0;
	sx[0] = t.root; si[0] = 0;
	while (top >= 0) {
		const int *nd = bt_node(t, sx[top]);
		const int nk = nd[0] & 255;
		if (!(nd[0] & BT_INTERNAL)) {
			for (int k = 0; k < nk; ++k) out[n++] = nd[1 + k];
			--top;
			continue;
		}
		const int i = si[top] >> 1;
		if (!(si[top] & 1)) { si[top] |= 1; ++top; sx[top] = nd[16 + i]; si[top] = 0; continue; }
		if (i < nk) { out[n++] = nd[1 + i]; si[top] = 2 * (i + 1); continue; }
		--top;
	}
	return n;
}

#define CHAIN_TPB 128

__global__ void __launch_bounds__(CHAIN_TPB) chain_build_kernel(const ChainWork cw, const ChainOpt o)
{
	const int lane = threadIdx.x & 31;
	const long long r = ((long long)blockIdx.x * CHAIN_TPB + threadIdx.x) >> 5;
	if (r >= cw.n) { if (r == cw.n && lane == 0) { cw.n_chains[r] = 0; cw.n_kept[r] = 0; } return; }
	const long long s0 = cw.soff[cw.off[r]];
	const int ns = (int)(cw.soff[cw.off[r + 1]] - s0);
	const Seed *sd = cw.seeds + s0;
	int *ord = cw.ord + s0, *ord2 = cw.ord2 + s0, *c_last = cw.c_last + s0, *c_n = cw.c_n + s0, *s_next = cw.s_next + s0;
	int nch = 0;

	// ---- insertion loop, bwamem.c:478-496: sequential by definition (every seed sees the chains its predecessors made) -> lane 0
	if (lane == 0 && ns > 0) {
		BTree t;
		t.nd = cw.bt + ((size_t)(s0 / 7) + 2 * (size_t)r) * BT_INTS; t.n_nodes = 0; t.sd = sd;
		t.root = bt_new(t, BT_LEAF);
		for (int e = 0; e < ns; ++e) {
			const Seed s = ld_seed(&sd[e]);
			if (s.rbeg < o.l_pac && o.l_pac < s.rbeg + s.len) continue;            // bridging the forward/reverse boundary
			const int f = nch ? bt_lower(t, s.rbeg) : -1;                          // kb_intervalp: the closest chain at or below rbeg
			if (f >= 0) {
				// test_and_merge, bwamem.c:334-356
				const int l = c_last[f];
				const Seed first = ld_seed(&sd[f]), last = ld_seed(&sd[l]);
				const long long qend = last.qbeg + last.len, rend = last.rbeg + last.len;
				if (s.qbeg >= first.qbeg && s.qbeg + s.len <= qend && s.rbeg >= first.rbeg && s.rbeg + s.len <= rend) continue;          // contained
				if (!((last.rbeg < o.l_pac || first.rbeg < o.l_pac) && s.rbeg >= o.l_pac)) {                                          // same strand
					const long long x = s.qbeg - last.qbeg, y = s.rbeg - last.rbeg;
					if (y >= 0 && x - y <= o.w && y - x <= o.w && x - last.len < o.max_chain_gap && y - last.len < o.max_chain_gap) {
						s_next[l] = e; s_next[e] = -1; c_last[f] = e; c_n[f] += 1;
						continue;
					}
				}
			}
			c_last[e] = e; c_n[e] = 1; s_next[e] = -1;                             // a new chain, keyed by rbeg (bwamem.c:489-494)
			bt_put(t, e);
			++nch;
		}
		if (nch) bt_walk(t, ord);                                                  // __kb_traverse, bwamem.c:608-610
	}
	nch = __shfl_sync(FULL_MASK, nch, 0);
	__syncwarp();

	int n_out = nch;
	if (o.do_flt && nch > 1) {
		// ---- mem_chain_flt, bwamem.c:629-700
		FltRec *a = cw.flt + s0;
		unsigned char *keep = cw.keep + s0;
		for (int k = lane; k < nch; k += 32) {
			const int f = ord[k], l = c_last[f];
			const Seed first = ld_seed(&sd[f]), last = ld_seed(&sd[l]);
			FltRec t;
			t.beg = first.qbeg; t.end = last.qbeg + last.len; t.w = chain_weight_dev(sd, s_next, f); t.p = k; t.p2 = -1;
			a[k] = t;
			keep[k] = 0;
		}
		__syncwarp();
		if (lane == 0) flt_introsort((size_t)nch, a);
		__syncwarp();
		for (int k = lane; k < nch; k += 32) { ord2[k] = ord[a[k].p]; a[k].p = k; }          // best chain first; p = rank from here on
		__syncwarp();
		if (lane == 0) {
			int n = 1;
			for (int i = 1; i < nch; ++i) {
				const FltRec ai = a[i];
				int j;
				for (j = 0; j < n; ++j) {
					const FltRec aj = a[j];
					const int b_max = aj.beg > ai.beg ? aj.beg : ai.beg, e_min = aj.end < ai.end ? aj.end : ai.end;
					if (e_min > b_max) {
						const int li = ai.end - ai.beg, lj = aj.end - aj.beg, min_l = li < lj ? li : lj;
						if ((float)(e_min - b_max) >= __fmul_rn((float)min_l, o.mask_level)) {         // int vs float, as in the reference
							if (aj.p2 < 0) a[j].p2 = ai.p;
							if ((float)ai.w < __fmul_rn((float)aj.w, o.chain_drop_ratio) && aj.w - ai.w >= o.min_seed_len << 1) break;
						}
					}
				}
				if (j == n) a[n++] = ai;
			}
			for (int i = 0; i < n; ++i) { keep[a[i].p] = 1; if (a[i].p2 >= 0) keep[a[i].p2] = 1; }
			n_out = 0;
			for (int i = 0; i < nch; ++i) if (keep[i]) ord[n_out++] = ord2[i];
		}
		n_out = __shfl_sync(FULL_MASK, n_out, 0);
		__syncwarp();
	}
	int kept = 0;
	for (int k = lane; k < n_out; k += 32) kept += c_n[ord[k]];
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) kept += __shfl_xor_sync(FULL_MASK, kept, m);
	if (lane == 0) { cw.n_chains[r] = n_out; cw.n_kept[r] = kept; }
}

// chains of read r -> chains_out[coff[r] ..], their seeds -> seeds_out[koff[r] ..] (insertion order inside a chain)
__global__ void __launch_bounds__(CHAIN_TPB) chain_emit_kernel(const ChainWork cw, const long long *__restrict__ coff, const long long *__restrict__ koff,
                                                               long long seed_base, Chain *__restrict__ chains_out, Seed *__restrict__ seeds_out)
{
	const int lane = threadIdx.x & 31;
	const long long r = ((long long)blockIdx.x * CHAIN_TPB + threadIdx.x) >> 5;
	if (r >= cw.n) return;
	const long long s0 = cw.soff[cw.off[r]];
	const Seed *sd = cw.seeds + s0;
	const int *ord = cw.ord + s0, *c_n = cw.c_n + s0, *s_next = cw.s_next + s0;
	const int n_out = cw.n_chains[r];
	const long long c0 = coff[r];
	long long run = koff[r];
	for (int base = 0; base < n_out; base += 32) {
		const int k = base + lane;
		const int f = k < n_out ? ord[k] : -1;
		const int nk = f >= 0 ? c_n[f] : 0;
		int incl = nk;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const int t = __shfl_up_sync(FULL_MASK, incl, d); if (lane >= d) incl += t; }
		const long long first = run + incl - nk;
		run += __shfl_sync(FULL_MASK, incl, 31);
		if (f >= 0) {
			long long at = first;
			for (int e = f; e >= 0; e = s_next[e]) seeds_out[at++] = sd[e];
			Chain c;
			c.pos = sd[f].rbeg; c.seed_first = first + seed_base; c.n_seeds = nk; c.weight = chain_weight_dev(sd, s_next, f);
			chains_out[c0 + k] = c;
		}
	}
}
