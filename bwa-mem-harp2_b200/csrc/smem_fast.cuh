// smem_fast.cuh -- the "k-mer count pyramid" fast path of the seeding kernel (sm_100a), and the kernels that build
// its tables.  Executable specification: tools/fast_model.c (checked bit for bit against the oracle on the CPU).
//
// Why: bwt_smem1 (bwt.c:776-835) spends ~85 % of its bwt_extend calls on patterns no longer than ~17 bases (the
// forward sweep until the match becomes unique, and the triangular backward sweep over the prev/curr lists), and all
// it ever looks at for those is the SIZE of their intervals (bwt.c:794-799, 813-824).  The bi-interval of a pattern
// P is a function of P alone: x[2] = occurrences of P in the indexed text T, x[0] = 1 + number of suffixes of T
// sorting before P, x[1] = x[0] of revcomp(P).  With 180 GB of HBM per GPU these can be tabulated for every k-mer up
// to D = DL+5 bases (DL = 12 on a 3.1 Gbp index: 25 GB):
//
//   cnt[L][code]  L = 1..DL    uint32  occurrences                       (direct; L2-resident up to L ~ 11)
//   cum[L][code]  L = 1..DL+1  uint64  x[0]
//   pyr[code]     L = DL+4     uint8   occurrences, saturating; the 64 entries that share a (DL+1)-mer prefix form ONE
//                                      64-byte block = one request of a lane pair, and the counts of levels DL+1..DL+3
//                                      are sums of 64 / 16 / 4 of its bytes
//   top[code]     L = DL+5     uint8   occurrences, saturating
//   255 = "unknown" (saturated, or one of the ~100 entries next to a suffix of T shorter than the level).
//
// A bwt_smem1 call then keeps its short entries as a BIT MASK over their ends (sizes are monotone in the pattern
// length, so "fails min_intv" is a prefix of the list and "same size as the previously pushed entry" is "no length in
// between changed the count"); one backward round over all of them costs one table fetch (1-2 DRAM requests)
// instead of one bwt_extend (1-2 dependent DRAM requests, ~240 instructions) per entry.  Only entries longer than D are
// REAL bi-intervals extended with the FM index exactly like the reference; an entry that outgrows the tables is
// materialised from cum/pyr/top (GROW).  Emitted short matches leave the kernel with x[0], x[1] still unknown
// (info bit 31) and are resolved by resolve_kernel, one thread per interval.  Whenever a table says "unknown" the
// whole read is handed to the plain FM kernel (seed_kernel) through the overflow list: results never depend on the tables.
#pragma once
#include "smem_kernels.cuh"

struct FastTables {
	const u32 *cnt;          // levels 1..DL, level L at lvl_off(L)
	const u64 *cum;          // levels 1..DL+1, same offsets
	const uint8_t *pyr, *top;
	int DL;
};

struct FastParams {
	SeedParams s;
	FastTables t;
	int *esc;                // [n] 1 = the read is on the overflow / escape list (re-run by the FM kernel)
	int q2_words;            // 32-bit words of the 2-bit staged query per pair (incl. 2 words of padding)
	int n_slots;             // read slots per CTA (<= FAST_MAX_SLOTS)
	int esc_cap;             // slot capacity assumed for the FM re-run of escaped reads (exact after its first round)
};

#define FAST_FLAG 0x80000000ull      // info bit 31: x[0], x[1] of this interval are still to be resolved from the tables
#define ESCV 0xffffffffu

// entries of level L start at 4 + 16 + ... + 4^(L-1)
__host__ __device__ __forceinline__ u64 lvl_off(int L) { return 0x5555555555555554ull & ((1ull << (2 * L)) - 1); }

__device__ __forceinline__ u32 has_ff(u32 w) { return (~w - 0x01010101u) & w & 0x80808080u; }
__device__ __forceinline__ u32 bytesum(u32 w) { return __dp4a(w, 0x01010101u, 0u); }

// reverse complement of an L-mer code (2 bits per base, first base most significant)
__host__ __device__ __forceinline__ u64 revcomp_code(u64 c, int L)
{
	u64 r = ~c;
#ifdef __CUDA_ARCH__
	r = __brevll(r);
#else
	{ u64 x = r, y = 0; for (int k = 0; k < 64; ++k) { y = (y << 1) | (x & 1); x >>= 1; } r = y; }
#endif
	r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
	return r >> (64 - 2 * L);
}

// cold state of the fast kernel (per pair, shared memory)
enum { FS_RK = 0, FS_RID = 4, FS_NOUT = 8, FS_MAXS_LO = 12, FS_MAXS_HI = 16, FS_STOP = 20 /* u32: size of the top virtual entry */,
       FS_START = 24, FS_STEP = 26, FS_ORI = 28, FS_SPLIT = 30, FS_PASS = 32, FS_X = 34, FS_NMEM = 36, FS_LMS = 38, FS_NM1 = 40,
       FS_KEEP = 42, FS_MAXLEN = 44, FS_MAXSTART = 46, FS_MAXEND = 48, FS_RET = 50, FS_AVAIL = 52, FS_SIZES = 56 /* u32[20] */,
       FS_BYTES = 56 + 80 };

enum { FP_IDLE = 0, FP_FWD_FM, FP_BWD_FM, FP_GROWN_FM, FP_FWD_TAB, FP_RND_TAB, FP_FWD_GROW, FP_BWD_GROW, FP_HOT_MAX = FP_BWD_GROW,
       FP_NEED_READ, FP_NEXT_STEP, FP_CALL_BEGIN, FP_FWD_END, FP_FWD_DONE, FP_ROUND_BEGIN, FP_ROUND_VIRT, FP_ROUND_END, FP_CALL_DONE };

// 32 bases of the 2-bit staged query starting at base `pos`, first base in the top bits
__device__ __forceinline__ u64 window64(u32 sq2, int pos)
{
	const u32 a = sq2 + 4u * ((u32)pos >> 4);
	const u32 w0 = (u32)lds_i32(a), w1 = (u32)lds_i32(a + 4), w2 = (u32)lds_i32(a + 8);
	const u32 off = 2u * ((u32)pos & 15u);
	return ((u64)__funnelshift_l(w1, w0, off) << 32) | (u64)__funnelshift_l(w2, w1, off);
}

// Threads of a CTA are not tied to reads: the per-read state lives in shared-memory SLOTS, and a slot that waits for
// memory is posted on the ready bitmap of its kind of operation (FM extend forward / backward / grown, table fetch
// forward / round, GROW forward / backward).  Every warp, on its own and without CTA barriers, takes up to 16 slots
// of ONE kind (the fullest bitmap), so that its lane pairs run the same code, performs the operation and the rare
// transitions that follow, and posts the slots again.  (Measured on the way here: one read pinned to one lane pair
// ran at 6.9 active lanes of 32; sorting the slots CTA-wide before every trip at 15 lanes but waiting at barriers.)
#ifndef FAST_TPB
#define FAST_TPB 256
#endif
#define FAST_MAX_SLOTS 192
#define FAST_WORDS (FAST_MAX_SLOTS / 32)
#define FAST_HDR 256                 // CTA header: ready bitmaps u32 [7][6] | idle-slot count
enum { FH_BYTES = 64 };              // hot state of a slot (registers while a pair works on it)

template <int MIN_BLOCKS>
__global__ void __launch_bounds__(FAST_TPB, MIN_BLOCKS) fast_kernel(const FastParams fp)
{
	typedef BEntry<false> BE;
	const SeedParams &p = fp.s;
	const FastTables &T = fp.t;
	extern __shared__ uint4 smem_raw[];
	const int tid = threadIdx.x, lane = tid & 31, half = lane & 1;
	const u32 pm = 3u << (lane & ~1);                       // this pair's lanes
	const int pair = tid >> 1;
	const u32 s0 = (u32)__cvta_generic_to_shared(smem_raw);
	const u32 s_slots = s0 + FAST_HDR;
	u32 *const rdy = reinterpret_cast<u32 *>(smem_raw);                         // [FP_HOT_MAX][FAST_WORDS]
	volatile u32 *const vrdy = rdy;
	u32 *const n_idle = rdy + FP_HOT_MAX * FAST_WORDS + 2;
	const int S = fp.n_slots;
	u32 sh = 0, sb = 0, sc = 0, ss = 0, sq = 0, sq2 = 0;    // the slot this pair works on: hot state, real entries, cold state, sizes, query (x2)
	Intv *M1 = nullptr, *BX = nullptr;
	int slot = 0;
	auto bind = [&](int sl) {
		slot = sl;
		sh = s_slots + (u32)sl * (u32)p.pair_stride;
		sb = sh + FH_BYTES;
		sc = sb + (u32)p.b_cap * BE::BYTES;
		ss = sc + FS_SIZES;
		sq = sc + FS_BYTES;
		sq2 = sq + (u32)p.q_stride;
		M1 = p.scratch + ((size_t)blockIdx.x * S + sl) * 3 * p.scratch_cap;
		BX = M1 + 2 * p.scratch_cap;
	};
	const int DL = T.DL, K = DL + 1, LP = DL + 4, D = DL + 5;

	auto b_put = [&](int idx, u64 x0, u64 x1, u64 x2, u32 end) {
		if (__builtin_expect(idx < p.b_cap, 1)) BE::put(sb + (u32)idx * BE::BYTES, x0, x1, x2, end);
		else bx_put(&BX[idx], x0, x1, x2, end);
	};
	auto b_get = [&](int idx, u64 &x0, u64 &x1, u64 &x2, u32 &end) {
		if (__builtin_expect(idx < p.b_cap, 1)) BE::get(sb + (u32)idx * BE::BYTES, x0, x1, x2, end);
		else { const Intv t = bx_get(&BX[idx]); x0 = t.x0; x1 = t.x1; x2 = t.x2; end = (u32)t.info; }
	};
	// bwt.c:815-820; `flag` = FAST_FLAG for a virtual entry (x0 = x1 = 0, resolved later)
	auto emit = [&](u64 x0, u64 x1, u64 x2, u32 end, int st, u64 flag) {
		const int n_mem = lds_u16(sc + FS_NMEM);
		if (n_mem == 0 || st < lds_u16(sc + FS_LMS)) {
			Intv *M = M1 + (size_t)lds_u16(sc + FS_PASS) * p.scratch_cap;
			st_intv(&M[n_mem], x0, x1, x2, (u64)end | flag | ((u64)st << 32));
			sts_u16(sc + FS_NMEM, n_mem + 1);
			sts_u16(sc + FS_LMS, st);
			const int l = (int)end - st;
			if (l >= lds_u16(sc + FS_MAXLEN)) {
				sts_u16(sc + FS_MAXLEN, l); sts_u16(sc + FS_MAXSTART, st); sts_u16(sc + FS_MAXEND, (int)end);
				sts_i32(sc + FS_MAXS_LO, (int)(u32)x2); sts_i32(sc + FS_MAXS_HI, (int)(u32)(x2 >> 32));
			}
		}
	};

	// ---- hot state
	int phase = FP_NEED_READ;
	u64 a = 1, b = 1, s = 1;             // FM operand: FWD (a = x[1], b = x[0]); BWD (a = x[0], b = x[1])
	u32 end = 0;
	int c = 0;
	int i = 0, j = 0, n_prev = 0, n_curr = 0, len = 0, guard = 0, x = 0;
	int max_count = 0;
	u64 min_intv = 1, last_s = 0;
	u32 E = 0;                           // virtual entries: bit b <-> end x + 1 + b
	int tpos = 0, tlo = 0, thi = 0;      // pending table fetch: window start, lengths wanted

	auto load_state = [&]() {
		const uint4 h0 = lds_v4(sh), h1 = lds_v4(sh + 16), h2 = lds_v4(sh + 32), h3 = lds_v4(sh + 48);
		a = (u64)h0.x | ((u64)h0.y << 32); b = (u64)h0.z | ((u64)h0.w << 32);
		s = (u64)h1.x | ((u64)h1.y << 32); last_s = (u64)h1.z | ((u64)h1.w << 32);
		E = h2.x; min_intv = h2.y; phase = (int)(h2.z & 255u); c = (int)(signed char)((h2.z >> 8) & 255u); end = h2.z >> 16;
		i = (int)(short)(h2.w & 0xffffu); j = (int)(h2.w >> 16);
		n_prev = (int)(h3.x & 0xffffu); n_curr = (int)(h3.x >> 16); x = (int)(h3.y & 0xffffu); len = (int)(h3.y >> 16);
		tpos = (int)(h3.z & 0xffffu); tlo = (int)((h3.z >> 16) & 255u); thi = (int)(h3.z >> 24); guard = (int)h3.w;
	};
	auto store_state = [&]() {
		if (!half) {
			sts_v4(sh, make_uint4((u32)a, (u32)(a >> 32), (u32)b, (u32)(b >> 32)));
			sts_v4(sh + 16, make_uint4((u32)s, (u32)(s >> 32), (u32)last_s, (u32)(last_s >> 32)));
			sts_v4(sh + 32, make_uint4(E, (u32)min_intv, (u32)phase | (((u32)c & 255u) << 8) | (end << 16), ((u32)i & 0xffffu) | ((u32)j << 16)));
			sts_v4(sh + 48, make_uint4((u32)n_prev | ((u32)n_curr << 16), (u32)x | ((u32)len << 16), (u32)tpos | ((u32)tlo << 16) | ((u32)thi << 24), (u32)guard));
		}
	};

	auto escape = [&]() {                // a table said "unknown": this read goes to the FM kernel
		const int rk = lds_i32(sc + FS_RK);
		if (!half) {
			fp.esc[rk] = 1;
			p.overflow_list[atomicAdd(&p.status[1], 1)] = lds_i32(sc + FS_RID);
			atomicMax(&p.status[3], fp.esc_cap);
			atomicAdd(&p.status[5], 1);
		}
		p.counts[rk] = 0;
		phase = FP_NEED_READ;
	};

	// rare transitions of the per-read state machine; ends in a phase that waits for memory (or FP_IDLE)
	auto cold = [&]() {
		while (phase > FP_HOT_MAX) {
			switch (phase) {
			case FP_NEED_READ: {
				int rk = 0;
				if (!half) { rk = atomicAdd(&p.status[0], 1); sts_i32(sc + FS_RK, rk); }
				__syncwarp(pm);
				rk = lds_i32(sc + FS_RK);
				if ((long long)rk >= p.n) { if (!half && max_count > 0) atomicMax(&p.status[4], max_count); phase = FP_IDLE; break; }
				const int rid = rk;
				const long long o0 = p.offs[rid];
				const uint8_t *q = p.seq + o0;
				len = (int)(p.offs[rid + 1] - o0);
				// stage the query twice: two bases per byte (0..3, else 4) and 2 bits per base, 16 bases per chunk
				const int n_chunks = fp.q2_words;
				for (int t = half; t < n_chunks; t += 2) {
					u32 nib0 = 0, nib1 = 0, two = 0;
#pragma unroll
					for (int k = 0; k < 16; ++k) {
						const int pos = 16 * t + k;
						const u32 v = pos < len ? min((u32)q[pos], 4u) : 4u;
						if (k < 8) nib0 |= v << (4 * k); else nib1 |= v << (4 * (k - 8));
						two |= (v & 3u) << (30 - 2 * k);
					}
					if (8 * t < p.q_stride) { sts_i32(sq + 8 * t, (int)nib0); sts_i32(sq + 8 * t + 4, (int)nib1); }
					sts_i32(sq2 + 4 * t, (int)two);
				}
				__syncwarp(pm);
				sts_i32(sc + FS_RID, rid); sts_i32(sc + FS_NOUT, 0); sts_u16(sc + FS_START, 0); sts_u16(sc + FS_STEP, 0);
				sts_u16(sc + FS_SPLIT, p.split_len_init < len ? p.split_len_init : len);     // bwamem.c:458
				phase = FP_NEXT_STEP;
			} break;
			case FP_NEXT_STEP: {     // head of smem_next2, bwamem.c:249-258
				int start = lds_u16(sc + FS_START);
				while (start < len && qbase(sq, start) > 3) ++start;
				if (start >= len) {
					const int rk = lds_i32(sc + FS_RK), n_out = lds_i32(sc + FS_NOUT);
					p.counts[rk] = n_out;
					max_count = max(max_count, n_out);
					if (n_out > p.slot_cap && !half) { fp.esc[rk] = 1; p.overflow_list[atomicAdd(&p.status[1], 1)] = lds_i32(sc + FS_RID); atomicMax(&p.status[3], n_out); }
					phase = FP_NEED_READ;
					break;
				}
				sts_u16(sc + FS_START, start); sts_u16(sc + FS_ORI, start); sts_u16(sc + FS_X, start); sts_u16(sc + FS_PASS, 0);
				min_intv = p.start_width < 1 ? 1 : (u64)p.start_width;
				phase = FP_CALL_BEGIN;
			} break;
			case FP_CALL_BEGIN: {    // bwt.c:783-789; the forward sweep starts with one table fetch for the window at x
				x = lds_u16(sc + FS_X);
				int avail = 0;                                  // valid bases from x, capped at D + 1
				while (avail <= D && x + avail < len && qbase(sq, x + avail) <= 3) ++avail;
				sts_u16(sc + FS_AVAIL, avail);
				sts_u16(sc + FS_NMEM, 0); sts_u16(sc + FS_MAXLEN, 0); sts_u16(sc + FS_MAXSTART, 0); sts_u16(sc + FS_MAXEND, 0);
				sts_i32(sc + FS_MAXS_LO, 0); sts_i32(sc + FS_MAXS_HI, 0);
				E = 0; n_curr = 0;
				guard = 4 * (len + 2) + 64;                     // > every trip one bwt_smem1 can need... per real entry; refreshed per round
				guard += 2 * (len + 2) * (len + 2);
				tpos = x; tlo = 1; thi = min(avail, D);
				phase = FP_FWD_TAB;
			} break;
			case FP_FWD_END: {       // bwt.c:800-803 / :806: push the last (real) interval
				b_put(n_curr++, b, a, s, end);
				phase = FP_FWD_DONE;
			} break;
			case FP_FWD_DONE: {      // bwt.c:807-809: the real entries are reversed in place (there are seldom more than two)
				for (int lo = 0, hi = n_curr - 1; lo < hi; ++lo, --hi) {
					u64 a0, a1, a2, b0, b1, b2; u32 ae, be;
					b_get(lo, a0, a1, a2, ae); b_get(hi, b0, b1, b2, be);
					__syncwarp(pm);
					b_put(lo, b0, b1, b2, be); b_put(hi, a0, a1, a2, ae);
				}
				int ret;
				if (n_curr > 0) { u64 t0, t1, t2; u32 te; b_get(0, t0, t1, t2, te); ret = (int)te; }
				else ret = x + 1 + (31 - __clz(E));
				sts_u16(sc + FS_RET, ret);
				if (lds_u16(sc + FS_PASS) == 0) sts_u16(sc + FS_START, ret);          // bwamem.c:262
				sts_i32(sc + FS_STOP, E ? lds_i32(ss + 4 * (32 - __clz(E))) : 0);      // size of the top virtual entry: n[b + 1]
				n_prev = n_curr; i = x;
				phase = FP_ROUND_BEGIN;
			} break;
			case FP_ROUND_BEGIN: {   // bwt.c:810-812
				--i; n_curr = 0; j = 0;
				c = i < 0 ? -1 : (int)qbase(sq, i);
				if (c > 3) c = -1;
				if (c < 0) {
					// nothing can be extended: only the first entry of the round can pass the containment test (bwt.c:815-820)
					if (n_prev > 0) { b_get(0, a, b, s, end); emit(a, b, s, end, i + 1, 0); }
					else if (E) emit(0, 0, (u64)(u32)lds_i32(sc + FS_STOP), (u32)(x + 1 + (31 - __clz(E))), i + 1, FAST_FLAG);
					phase = FP_CALL_DONE;
				} else if (n_prev > 0) { b_get(0, a, b, s, end); phase = FP_BWD_FM; }
				else phase = FP_ROUND_VIRT;
			} break;
			case FP_ROUND_VIRT: {    // the real entries of this round are done; now the virtual ones
				const int r = x - i, bg = D - r;
				if (bg >= 0 && ((E >> bg) & 1u)) phase = FP_BWD_GROW;          // it reaches length D + 1: make it real
				else if (E) { tpos = i; tlo = __ffs(E) + r; thi = 32 - __clz(E) + r; phase = FP_RND_TAB; }
				else phase = FP_ROUND_END;
			} break;
			case FP_ROUND_END: {     // bwt.c:826-827
				if (n_curr == 0 && E == 0) phase = FP_CALL_DONE;
				else { n_prev = n_curr; phase = FP_ROUND_BEGIN; }
			} break;
			case FP_CALL_DONE: {
				const int rk = lds_i32(sc + FS_RK), n_mem = lds_u16(sc + FS_NMEM);
				Intv *const slot = p.slots + (size_t)rk * p.slot_cap;
				const Intv *const M2 = M1 + p.scratch_cap;
				const int step = lds_u16(sc + FS_STEP);
				int n_out = lds_i32(sc + FS_NOUT);
				const u64 tag = (u64)step << STEP_SHIFT;
				if (lds_u16(sc + FS_PASS) == 0) {
					const int max_len = lds_u16(sc + FS_MAXLEN), split_len = lds_u16(sc + FS_SPLIT);
					const u64 max_s = (u64)(u32)lds_i32(sc + FS_MAXS_LO) | ((u64)(u32)lds_i32(sc + FS_MAXS_HI) << 32);
					// bwamem.c:272: re-seed from the middle of the longest SMEM if it is long and (nearly) unique
					if (n_mem > 0 && split_len > 0 && max_len >= split_len && max_s <= (u64)p.split_width) {
						sts_u16(sc + FS_NM1, n_mem); sts_u16(sc + FS_KEEP, max_len); sts_u16(sc + FS_PASS, 1);
						sts_u16(sc + FS_X, (lds_u16(sc + FS_MAXEND) + lds_u16(sc + FS_MAXSTART)) >> 1);
						min_intv = max_s + 1;
						phase = FP_CALL_BEGIN;
						break;
					}
					for (int e = n_mem - 1; e >= 0; --e) {
						if (n_out < p.slot_cap) { const Intv t = ld_intv(&M1[e]); st_intv(&slot[n_out], t.x0, t.x1, t.x2, t.info | tag); }
						++n_out;
					}
				} else {
					// ordered merge, bwamem.c:281-301; both lists are walked in ascending start = reverse emission
					int ia = lds_u16(sc + FS_NM1) - 1, ib = n_mem - 1;
					const int half_len = lds_u16(sc + FS_KEEP) >> 1, ori_start = lds_u16(sc + FS_ORI);
					Intv va, vb;
					va.x0 = va.x1 = va.x2 = va.info = 0; vb = va;
					bool have_a = false, have_b = false;
					while (ia >= 0 || ib >= 0) {
						if (ia >= 0 && !have_a) { va = ld_intv(&M1[ia]); have_a = true; }
						if (ib >= 0 && !have_b) { vb = ld_intv(&M2[ib]); have_b = true; }
						bool take_a;
						if (ia >= 0 && ib >= 0) {
							const long long xa = (long long)((va.info >> 32 << 32) | (u32)(len - (int)(u32)(va.info & 0xffffu)));
							const long long xb = (long long)((vb.info >> 32 << 32) | (u32)(len - (int)(u32)(vb.info & 0xffffu)));
							take_a = xa < xb;
						} else take_a = ia >= 0;
						if (take_a) {
							if (n_out < p.slot_cap) st_intv(&slot[n_out], va.x0, va.x1, va.x2, va.info | tag);
							++n_out; --ia; have_a = false;
						} else {
							const int e_ = (int)(u32)(vb.info & 0xffffu), sl = e_ - (int)(vb.info >> 32);
							if (sl >= half_len && e_ > ori_start) {
								if (n_out < p.slot_cap) st_intv(&slot[n_out], vb.x0, vb.x1, vb.x2, vb.info | tag);
								++n_out;
							}
							--ib; have_b = false;
						}
					}
				}
				sts_i32(sc + FS_NOUT, n_out); sts_u16(sc + FS_STEP, step + 1);
				phase = FP_NEXT_STEP;
			} break;
			default: break;
			}
		}
	};

	// post the slot for its next operation (or retire it)
	auto post = [&]() {
		store_state();
		__threadfence_block();
		__syncwarp(pm);
		if (!half) {
			if (phase == FP_IDLE) atomicAdd(n_idle, 1u);
			else atomicOr(&rdy[(phase - 1) * FAST_WORDS + (slot >> 5)], 1u << (slot & 31));
		}
	};

	for (int t = tid; t < FAST_HDR / 4; t += FAST_TPB) rdy[t] = 0;
	__syncthreads();
	for (int sl = pair; sl < S; sl += FAST_TPB / 2) {
		bind(sl);
		phase = FP_NEED_READ;
		cold();
		post();
	}
	for (;;) {
		__syncwarp();
		// ---- the kind of operation with the most ready slots ...
		int cnt = 0;
		if (lane < FP_HOT_MAX) {
#pragma unroll
			for (int wd = 0; wd < FAST_WORDS; ++wd) cnt += __popc(vrdy[lane * FAST_WORDS + wd]);
		}
		int best = cnt, bk = lane;
#pragma unroll
		for (int off = 4; off >= 1; off >>= 1) {
			const int o = __shfl_down_sync(FULL_MASK, best, off), ok_ = __shfl_down_sync(FULL_MASK, bk, off);
			if (o > best) { best = o; bk = ok_; }
		}
		best = __shfl_sync(FULL_MASK, best, 0); bk = __shfl_sync(FULL_MASK, bk, 0);
		if (best == 0) {
			u32 idle = 0;
			if (lane == 0) idle = *(volatile u32 *)n_idle;
			idle = __shfl_sync(FULL_MASK, idle, 0);
			if (idle >= (u32)S) break;
			__nanosleep(200);
			continue;
		}
		// ---- ... up to 16 of them become this warp's (a bit seen set in the value atomicAnd returns is ours)
		const u32 m = lane < FAST_WORDS ? vrdy[bk * FAST_WORDS + lane] : 0u;
		const int c_ = __popc(m);
		int pre = c_;
#pragma unroll
		for (int off = 1; off < 8; off <<= 1) { const int t_ = __shfl_up_sync(FULL_MASK, pre, off); if (lane >= off) pre += t_; }
		pre -= c_;
		const int take = min(max(16 - pre, 0), c_);
		u32 rem = m;
		for (int k = 0; k < take; ++k) rem &= rem - 1u;        // drop the lowest `take` set bits
		const u32 bits = m ^ rem;
		u32 got = 0;
		if (bits) got = atomicAnd(&rdy[bk * FAST_WORDS + lane], ~bits) & bits;
		const int g_ = __popc(got);
		int gin = g_;
#pragma unroll
		for (int off = 1; off < 8; off <<= 1) { const int t_ = __shfl_up_sync(FULL_MASK, gin, off); if (lane >= off) gin += t_; }
		const int total = __shfl_sync(FULL_MASK, gin, FAST_WORDS - 1);
		const int pi = lane >> 1;
		int my = -1;
#pragma unroll
		for (int wd = 0; wd < FAST_WORDS; ++wd) {
			const u32 gw = __shfl_sync(FULL_MASK, got, wd);
			const int ge = __shfl_sync(FULL_MASK, gin, wd) - __popc(gw);
			if (pi >= ge && pi < ge + __popc(gw)) {
				u32 t_ = gw;
				for (int k = 0; k < pi - ge; ++k) t_ &= t_ - 1u;
				my = 32 * wd + __ffs(t_) - 1;
			}
		}
		if (total == 0 || my < 0) continue;
		__threadfence_block();
		bind(my);
		load_state();
		do {
		if (--guard < 0) { if (!half) atomicAdd(&p.status[2], 1); p.counts[lds_i32(sc + FS_RK)] = 0; phase = FP_NEED_READ; break; }

		// ============================================================== one memory round trip per pair
		// (the warp is divergent by kind of operation, but the loads of all kinds are issued before any result is used)
		const bool op_fm = phase <= FP_GROWN_FM, op_tab = phase == FP_FWD_TAB || phase == FP_RND_TAB;
		u32 w[8], v[8];                      // FM: K / L sectors; TAB: pyramid sector / direct counts + top word; GROW: the whole block
		u64 gcum = 0, W = 0;
		u32 gtop = 0;
		u64 kk = 0, ll = 0;
#pragma unroll
		for (int t = 0; t < 8; ++t) w[t] = v[t] = 0;
		if (op_fm) {
			// bwt_extend (bwt.c:416-429): one or two occ blocks, one request each (see smem_device.cuh)
			const u64 k = a - 1, l = a - 1 + s;
			kk = k - (k >= p.ix.primary); ll = l - (l >= p.ix.primary);
			const uint4 *pk = p.ix.blk + (kk >> 7) * 4 + half * 2, *pl = p.ix.blk + (ll >> 7) * 4 + half * 2;
			ld_sector(w, pk);
			if ((kk >> 7) != (ll >> 7)) ld_sector(v, pl);
		} else if (op_tab) {
			W = window64(sq2, tpos);
			const int hi_d = min(thi, DL);
#pragma unroll
			for (int t = 0; t < 7; ++t) {
				const int L = 2 * t + 2 - half;
				if (L >= tlo && L <= hi_d) v[t] = __ldg(T.cnt + lvl_off(L) + (W >> (64 - 2 * L)));
			}
			if (thi >= K) ld_sector(w, reinterpret_cast<const uint4 *>(T.pyr + ((W >> (64 - 2 * K)) << 6) + 32 * half));
			if (thi >= D) v[7] = __ldg(reinterpret_cast<const u32 *>(T.top + ((W >> (64 - 2 * D)) & ~3ull)));
		} else {
			// GROW: lane 0 computes x[0] of the D-mer, lane 1 x[0] of its reverse complement (= x[1])
			const u64 P = window64(sq2, phase == FP_FWD_GROW ? x : i + 1) >> (64 - 2 * D);
			W = half ? revcomp_code(P, D) : P;
			const u64 c16 = W >> 2;
			gcum = __ldg(T.cum + lvl_off(K) + (c16 >> 6));
			const uint4 *bp = reinterpret_cast<const uint4 *>(T.pyr + (c16 & ~63ull));
			ld_sector(w, bp); ld_sector(v, bp + 2);
			gtop = __ldg(reinterpret_cast<const u32 *>(T.top + (W & ~3ull)));
		}

		if (op_fm) {
			if ((kk >> 7) == (ll >> 7)) {
#pragma unroll
				for (int t = 0; t < 8; ++t) v[t] = w[t];
			}
			const int rk_ = min(max((int)(kk & 127) + 1 - 64 * half, 0), 64), rl_ = min(max((int)(ll & 127) + 1 - 64 * half, 0), 64);
			u32 ck = occ_half(w, rk_), cl = occ_half(v, rl_);
			ck += __shfl_xor_sync(pm, ck, 1);
			cl += __shfl_xor_sync(pm, cl, 1);
			const int j0 = 2 * half, cc = c & 3;
			const u64 tk0 = ((u64)w[0] | ((u64)w[1] << 32)) + ((ck >> (8 * j0)) & 0xffu), tk1 = ((u64)w[2] | ((u64)w[3] << 32)) + ((ck >> (8 * j0 + 8)) & 0xffu);
			const u64 tl0 = ((u64)v[0] | ((u64)v[1] << 32)) + ((cl >> (8 * j0)) & 0xffu), tl1 = ((u64)v[2] | ((u64)v[3] << 32)) + ((cl >> (8 * j0 + 8)) & 0xffu);
			const u64 sz0 = tl0 - tk0, sz1 = tl1 - tk1;
			u64 acc = (j0 > cc ? sz0 : 0) + (j0 + 1 > cc ? sz1 : 0);
			acc += __shfl_xor_sync(pm, acc, 1);
			const u64 tkc = (cc & 1) ? tk1 : tk0, szc = (cc & 1) ? sz1 : sz0;
			const int owner = (lane & ~1) | (cc >> 1);
			Ext ok;
			ok.a = p.ix.L2[cc] + 1 + __shfl_sync(pm, tkc, owner);
			ok.s = __shfl_sync(pm, szc, owner);
			ok.b = b + ((a <= p.ix.primary && a + s - 1 >= p.ix.primary) ? 1 : 0) + acc;

			const bool small = ok.s < min_intv;
			if (phase == FP_FWD_FM) {                            // bwt.c:794-799
				const bool diff = ok.s != s;
				if (diff) b_put(n_curr++, b, a, s, end);
				if (diff && small) { phase = FP_FWD_DONE; break; }
				a = ok.a; b = ok.b; s = ok.s; end = (u32)(i + 1);
				++i;
				const u32 qv = i < len ? qbase(sq, i) : 4u;
				if (qv > 3) phase = FP_FWD_END;
				else c = 3 - (int)qv;
			} else {                                             // bwt.c:813-824
				if (small) { if (n_curr == 0) emit(a, b, s, end, i + 1, 0); }
				else if (n_curr == 0 || ok.s != last_s) { b_put(n_curr++, ok.a, ok.b, ok.s, end); last_s = ok.s; }
				if (phase == FP_GROWN_FM) phase = FP_ROUND_VIRT;
				else if (++j == n_prev) phase = FP_ROUND_VIRT;
				else b_get(j, a, b, s, end);
			}
		} else if (op_tab) {
			// n[L] of the window -> shared memory
			const int hi_d = min(thi, DL);
#pragma unroll
			for (int t = 0; t < 7; ++t) {
				const int L = 2 * t + 2 - half;
				if (L >= tlo && L <= hi_d) sts_i32(ss + 4 * L, (int)v[t]);
			}
			if (thi >= K) {
				const u32 t6 = (u32)(W >> (64 - 2 * LP)) & 63u;
				u32 sm[8], ff[8];
#pragma unroll
				for (int t = 0; t < 8; ++t) { sm[t] = bytesum(w[t]); ff[t] = has_ff(w[t]); }
				const u32 g0 = sm[0] + sm[1] + sm[2] + sm[3], g1 = sm[4] + sm[5] + sm[6] + sm[7];
				const u32 f0 = ff[0] | ff[1] | ff[2] | ff[3], f1 = ff[4] | ff[5] | ff[6] | ff[7];
				u32 tot = g0 + g1, ftot = f0 | f1;
				tot += __shfl_xor_sync(pm, tot, 1); ftot |= __shfl_xor_sync(pm, ftot, 1);
				const u32 wi = (t6 >> 2) & 7u;
				u32 selw = w[0], sels = sm[0], self = ff[0];
#pragma unroll
				for (int t = 1; t < 8; ++t) if (wi == (u32)t) { selw = w[t]; sels = sm[t]; self = ff[t]; }
				const u32 selg = (t6 & 16u) ? g1 : g0, selgf = (t6 & 16u) ? f1 : f0;
				const u32 byte = (selw >> (8 * (t6 & 3u))) & 255u;
				const int owner = (lane & ~1) | (int)(t6 >> 5);
				const u32 c14 = __shfl_sync(pm, selgf ? ESCV : selg, owner), c15 = __shfl_sync(pm, self ? ESCV : sels, owner),
				          c16 = __shfl_sync(pm, byte == 255u ? ESCV : byte, owner);
				if (!half) {
					sts_i32(ss + 4 * K, (int)(ftot ? ESCV : tot)); sts_i32(ss + 4 * (K + 1), (int)c14);
					sts_i32(ss + 4 * (K + 2), (int)c15); sts_i32(ss + 4 * (K + 3), (int)c16);
				}
			}
			if (thi >= D && !half) {
				const u32 byte = (v[7] >> (8 * ((u32)(W >> (64 - 2 * D)) & 3u))) & 255u;
				sts_i32(ss + 4 * D, (int)(byte == 255u ? ESCV : byte));
			}
			__syncwarp(pm);
			if (phase == FP_FWD_TAB) {
				// forward sweep over the table levels (bwt.c:790-806): an entry is pushed where the count changes
				const int avail = lds_u16(sc + FS_AVAIL), Lend = thi;
				u32 prevn = (u32)lds_i32(ss + 4);
				bool stop = false, bad = prevn == ESCV;
				for (int L = 1; L < Lend; ++L) {
					const u32 nx = (u32)lds_i32(ss + 4 * (L + 1));
					if (nx == ESCV) { bad = true; break; }
					if (nx != prevn) { E |= 1u << (L - 1); if ((u64)nx < min_intv) { stop = true; break; } }
					prevn = nx;
				}
				if (bad) { escape(); break; }
				if (stop) phase = FP_FWD_DONE;
				else if (avail <= D) { E |= 1u << (Lend - 1); phase = FP_FWD_DONE; }
				else phase = FP_FWD_GROW;                 // the pattern outgrows the tables: FM from length D on
			} else {
				// one backward round over all virtual entries (bwt.c:813-824 with sizes from the table)
				const int r = x - i;
				u32 Et = E, E2 = 0, s_top = (u32)lds_i32(sc + FS_STOP), s_new = 0;
				bool first = true, bad = false;
				while (Et) {
					const int bb = 31 - __clz(Et);
					Et &= ~(1u << bb);
					const u32 n = (u32)lds_i32(ss + 4 * (bb + r + 1));
					if (n == ESCV) { bad = true; break; }
					if ((u64)n < min_intv) { if (first && n_curr == 0) emit(0, 0, (u64)s_top, (u32)(x + 1 + bb), i + 1, FAST_FLAG); }
					else if ((n_curr == 0 && E2 == 0) || (u64)n != last_s) { if (!E2) s_new = n; E2 |= 1u << bb; last_s = n; }
					first = false;
				}
				if (bad) { escape(); break; }
				E = E2;
				sts_i32(sc + FS_STOP, (int)s_new);
				phase = FP_ROUND_END;
			}
		} else {
			// GROW: x[0] = cum[K-mer prefix] + counts of the smaller (DL+4)-mers of the block + smaller siblings at the top level
			const u32 t6 = (u32)(W >> 2) & 63u, sib = (u32)W & 3u;
			u64 sum = gcum;
			u32 bad = 0;
#pragma unroll
			for (int t = 0; t < 16; ++t) {
				const int nb = min(max((int)t6 - 4 * t, 0), 4), nb1 = min(max((int)t6 + 1 - 4 * t, 0), 4);
				const u32 m0 = nb == 4 ? 0xffffffffu : ((1u << (8 * nb)) - 1u), m1 = nb1 == 4 ? 0xffffffffu : ((1u << (8 * nb1)) - 1u);
				const u32 word = t < 8 ? w[t & 7] : v[t & 7];
				sum += bytesum(word & m0);
				bad |= has_ff(word & m1);
			}
			sum += bytesum(gtop & ((1u << (8 * sib)) - 1u));
			bad |= has_ff(gtop & (sib == 3 ? 0xffffffffu : ((1u << (8 * (sib + 1))) - 1u)));
			const u32 cnt = (gtop >> (8 * sib)) & 255u;
			const u64 x0 = __shfl_sync(pm, sum, lane & ~1), x1 = __shfl_sync(pm, sum, lane | 1);
			const u32 sz = __shfl_sync(pm, cnt, lane & ~1);
			bad |= __shfl_xor_sync(pm, bad, 1);
			if (bad) { escape(); break; }
			if (phase == FP_FWD_GROW) {
				a = x1; b = x0; s = sz; end = (u32)(x + D); i = x + D;
				const u32 qv = i < len ? qbase(sq, i) : 4u;       // (avail > D: this base is valid)
				if (qv > 3) phase = FP_FWD_END;
				else { c = 3 - (int)qv; phase = FP_FWD_FM; }
			} else {
				const int r = x - i, bg = D - r;
				a = x0; b = x1; s = sz; end = (u32)(x + 1 + bg);
				E &= ~(1u << bg);
				// the next virtual entry is the top one now; its size in the previous round is still in sizes[]
				sts_i32(sc + FS_STOP, E ? lds_i32(ss + 4 * (32 - __clz(E) + r - 1)) : 0);
				phase = FP_GROWN_FM;
			}
		}
		} while (0);
		cold();
		post();
	}
}

// ---------------------------------------------------------------------------------------------
// x[0] of the L-mer `code` from the tables (tools/fast_model.c tab_x0); returns false on "unknown"
__device__ __forceinline__ bool table_x0(const FastTables &T, int L, u64 code, u64 &out)
{
	const int K = T.DL + 1, LP = T.DL + 4;
	if (L <= K) { out = T.cum[lvl_off(L) + code]; return true; }
	const int Lb = L < LP ? L : LP;
	const u64 c16 = L <= LP ? code : code >> (2 * (L - LP));
	const u64 pre = c16 >> (2 * (Lb - K)), blk = pre << (2 * (LP - K));
	const u64 lo = c16 << (2 * (LP - Lb)), span = 1ull << (2 * (LP - Lb));
	u64 sum = T.cum[lvl_off(K) + pre];
	for (u64 k = blk; k < lo + span; ++k) {
		const u32 v = T.pyr[k];
		if (v == 255u) return false;
		if (k < lo) sum += v;
	}
	if (L > LP) {
		const u64 sib = code & ~3ull;
		for (u64 k = sib; k <= code; ++k) {
			const u32 v = T.top[k];
			if (v == 255u) return false;
			if (k < code) sum += v;
		}
	}
	out = sum;
	return true;
}

// Intervals the fast kernel emitted from the tables still lack x[0] / x[1] (info bit 31): one thread per slot entry
// looks them up from the pattern q[start..end).  A table that answers "unknown" sends the read to the FM re-run.
__global__ void __launch_bounds__(128) resolve_kernel(const FastParams fp)
{
	const SeedParams &p = fp.s;
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t >> 3;
	if (r >= p.n) return;
	if (fp.esc[r]) return;
	const int cnt = min(p.counts[r], p.slot_cap);
	const uint8_t *q = p.seq + p.offs[r];
	for (int e = (int)(t & 7); e < cnt; e += 8) {
		Intv *sl = &p.slots[(size_t)r * p.slot_cap + e];
		const Intv v = ld_intv(sl);
		if (!(v.info & FAST_FLAG)) continue;
		const int st = (int)((v.info >> 32) & 0xffffu), en = (int)(v.info & 0xffffu), L = en - st;
		u64 code = 0;
		for (int k = st; k < en; ++k) code = (code << 2) | (u64)(q[k] & 3u);
		u64 x0 = 0, x1 = 0;
		if (!table_x0(fp.t, L, code, x0) || !table_x0(fp.t, L, revcomp_code(code, L), x1)) {
			if (atomicExch(&fp.esc[r], 1) == 0) {
				p.overflow_list[atomicAdd(&p.status[1], 1)] = (int)r;
				atomicMax(&p.status[3], fp.esc_cap);
				atomicAdd(&p.status[5], 1);
			}
			return;
		}
		st_intv(sl, x0, x1, v.x2, v.info & ~FAST_FLAG);
	}
}

// ---------------------------------------------------------------------------------------------
// Table construction from the 2-bit text (bntseq.c:268-273: T = forward + reverse complement; the forward half is
// the .pac array of the reference, 4 bases per byte, first base in the top bits, bntseq.h:_get_pac).

// doubled text, 32 bases per 64-bit word, first base in the top bits, zero padded
__global__ void pack_text_kernel(const uint8_t *__restrict__ pac, long long l_pac, u64 *__restrict__ tw, long long n_words)
{
	const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n_words) return;
	const long long n = 2 * l_pac;
	u64 v = 0;
	for (int k = 0; k < 32; ++k) {
		const long long pos = 32 * w + k;
		u32 base = 0;
		if (pos < l_pac) base = (pac[pos >> 2] >> ((~pos & 3) << 1)) & 3u;
		else if (pos < n) { const long long o = n - 1 - pos; base = 3u - ((pac[o >> 2] >> ((~o & 3) << 1)) & 3u); }
		v = (v << 2) | base;
	}
	tw[w] = v;
}

__device__ __forceinline__ void sat_inc_u8(uint8_t *tab, u64 idx)
{
	u32 *w = reinterpret_cast<u32 *>(tab + (idx & ~3ull));
	const u32 sh = (u32)(idx & 3ull) * 8u;
	u32 old = *w;
	for (;;) {
		if (((old >> sh) & 255u) == 255u) return;
		const u32 seen = atomicCAS(w, old, old + (1u << sh));
		if (seen == old) return;
		old = seen;
	}
}

// one thread per text position: count its K-mer (uint32), its (DL+4)-mer and its (DL+5)-mer (saturating uint8)
__global__ void __launch_bounds__(256) kmer_hist_kernel(const u64 *__restrict__ tw, long long n, int DL, u32 *__restrict__ cntK,
                                                        uint8_t *__restrict__ pyr, uint8_t *__restrict__ top)
{
	const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (pos >= n) return;
	const int K = DL + 1, LP = DL + 4, D = DL + 5;
	const u64 w0 = tw[pos >> 5], w1 = tw[(pos >> 5) + 1];
	const int o = 2 * (int)(pos & 31);
	const u64 W = o ? (w0 << o) | (w1 >> (64 - o)) : w0;
	if (pos + K <= n) atomicAdd(&cntK[W >> (64 - 2 * K)], 1u);
	if (pos + LP <= n) sat_inc_u8(pyr, W >> (64 - 2 * LP));
	if (pos + D <= n) sat_inc_u8(top, W >> (64 - 2 * D));
}

// level L-1 from level L: sum of the four children, plus the suffix of T of length L-1 (which no L-mer covers)
__global__ void kmer_reduce_kernel(const u32 *__restrict__ child, u32 *__restrict__ parent, long long n_parent, u64 tail_code)
{
	const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n_parent) return;
	const uint4 v = reinterpret_cast<const uint4 *>(child)[q];
	parent[q] = v.x + v.y + v.z + v.w + ((u64)q == tail_code ? 1u : 0u);
}

// exclusive scan of uint32 counts into uint64 x[0] values: local pass (1024 per block) ...
__global__ void __launch_bounds__(SCAN_TPB) cum_local_kernel(const u32 *__restrict__ cnt, long long n, u64 *__restrict__ cum, u64 *__restrict__ bsum)
{
	__shared__ u64 wsum[SCAN_TPB / 32];
	const long long base = (long long)blockIdx.x * SCAN_PER_BLOCK + (long long)threadIdx.x * SCAN_PER_THREAD;
	u32 v[SCAN_PER_THREAD];
	u64 t = 0;
#pragma unroll
	for (int k = 0; k < SCAN_PER_THREAD; ++k) { v[k] = base + k < n ? cnt[base + k] : 0; t += v[k]; }
	u64 incl = t;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const u64 o = __shfl_up_sync(FULL_MASK, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
	if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
	__syncthreads();
	u64 woff = 0, total = 0;
#pragma unroll
	for (int w = 0; w < SCAN_TPB / 32; ++w) { if (w < (int)(threadIdx.x >> 5)) woff += wsum[w]; total += wsum[w]; }
	u64 run = woff + incl - t;
#pragma unroll
	for (int k = 0; k < SCAN_PER_THREAD; ++k) { if (base + k < n) cum[base + k] = run; run += v[k]; }
	if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// ... block sums (scan_bsum_kernel of smem_kernels.cuh works on long long; same bits) ... and the final add:
// + block offset + 1 (row 0 is the '$' suffix) + the suffixes of T shorter than L that sort before the L-mer
struct TailCodes { u64 code[20]; };      // code[a] = the last a symbols of T
__global__ void cum_add_kernel(u64 *__restrict__ cum, long long n, const u64 *__restrict__ bsum, int L, TailCodes tails)
{
	const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= n) return;
	u64 v = cum[q] + bsum[q / SCAN_PER_BLOCK] + 1;
	for (int a = 1; a < L; ++a) v += (u64)q >= (tails.code[a] << (2 * (L - a))) ? 1 : 0;
	cum[q] = v;
}
