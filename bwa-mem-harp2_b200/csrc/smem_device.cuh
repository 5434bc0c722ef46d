// smem_device.cuh -- device side of the SMEM seeding path (sm_100a).
//
// What is computed (bit-exact with the reference's CPU code, see oracle/smem_oracle.c):
//   occ_in_block  <- bwt_occ4      bwt.c:187-204
//   extend        <- bwt_extend    bwt.c:416-429 (bwt_2occ4 = two occ4, bwt.c:213-214)
//   seed_kernel   <- bwt_smem1     bwt.c:776-835, smem_next2 bwamem.c:244-305,
//                    enumeration loop of mem_insert_seed bwamem.c:453-460
//
// How: one read per thread, but written as a *state machine whose every iteration performs
// exactly one bwt_extend*.  All lanes of a warp therefore re-converge on the expensive part
// (two 64-byte occ-block gathers + popcounts) no matter which read / pass / direction each lane
// is in; only the cheap bookkeeping between extends diverges.  Threads are persistent and pull
// read indices from a global counter (the AFU's free-PE dispatch, afu_core.v:3375-3440).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

struct DevIndex {
	const uint4 *blk;   // bwt_t::bwt viewed as 64-byte occ blocks (padded copy in HBM)
	u64 primary;
	u64 L2[5];
	u64 seq_len;
};

struct __align__(32) Intv { u64 x0, x1, x2, info; };

struct SeedParams {
	DevIndex ix;
	const uint8_t *seq;      // staged reads, one base per byte
	const long long *offs;   // [n+1]
	long long n;
	const int *list;         // optional indirection (overflow re-run): read id = list[k]
	const int *xs, *min_intvs; // smem1 mode inputs
	int *ret;                // smem1 mode output
	Intv *slots;             // [n][slot_cap] per-read result slots
	int slot_cap;
	int *counts;             // [n] exact interval count per read (even when > slot_cap)
	int *overflow_list;      // read ids with count > slot_cap
	int *status;             // [0] work counter, [1] n_overflow, [2] guard trips, [3] largest overflowing count
	Intv *scratch;           // per-thread: 4 arrays of scratch_cap entries
	int scratch_cap;
	int split_len_init, split_width, start_width;
	u64 hot_min_intv;        // 0 = off; L2 evict_last hint for occ blocks of intervals >= this size
};

// ---------------------------------------------------------------------------------------------
// memory helpers

// One 64-byte occ block as two 32-byte sectors (256-bit loads, new on sm_100), read-only path,
// no L1 allocation: these lines are never re-used by the same SM before eviction.
struct OccBlock { u32 w[16]; };

__device__ __forceinline__ void ld_block(OccBlock &b, const uint4 *p)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
	             : "l"(p));
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(b.w[8]), "=r"(b.w[9]), "=r"(b.w[10]), "=r"(b.w[11]), "=r"(b.w[12]), "=r"(b.w[13]), "=r"(b.w[14]), "=r"(b.w[15])
	             : "l"(p + 2));
}

__device__ __forceinline__ void ld_block_hot(OccBlock &b, const uint4 *p, u64 policy)
{
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
	             : "=r"(b.w[0]), "=r"(b.w[1]), "=r"(b.w[2]), "=r"(b.w[3]), "=r"(b.w[4]), "=r"(b.w[5]), "=r"(b.w[6]), "=r"(b.w[7])
	             : "l"(p), "l"(policy));
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
	             : "=r"(b.w[8]), "=r"(b.w[9]), "=r"(b.w[10]), "=r"(b.w[11]), "=r"(b.w[12]), "=r"(b.w[13]), "=r"(b.w[14]), "=r"(b.w[15])
	             : "l"(p + 2), "l"(policy));
}

__device__ __forceinline__ Intv ld_intv(const Intv *p)
{
	Intv v;
	const ulonglong2 *q = reinterpret_cast<const ulonglong2 *>(p);
	ulonglong2 a = q[0], b = q[1];
	v.x0 = a.x; v.x1 = a.y; v.x2 = b.x; v.info = b.y;
	return v;
}
__device__ __forceinline__ void st_intv(Intv *p, u64 x0, u64 x1, u64 x2, u64 info)
{
	ulonglong2 *q = reinterpret_cast<ulonglong2 *>(p);
	q[0] = make_ulonglong2(x0, x1);
	q[1] = make_ulonglong2(x2, info);
}

// ---------------------------------------------------------------------------------------------
// occ: counts of C,G,T (and by difference A) among the first r symbols (1..128) of a block.
// Two words are merged before each popcount (the bit planes only occupy even bit positions),
// so a block costs 12 POPC instead of the LUT walk of bwt.c:200-203.
struct Cnt4 { u32 c[4]; };

__device__ __forceinline__ Cnt4 occ_in_block(const OccBlock &b, u32 r)
{
	u32 nh = 0, nl = 0, nt = 0;
#pragma unroll
	for (int j = 0; j < 8; j += 2) {
		int ta = (int)r - 16 * j, tb = ta - 16;
		ta = max(min(ta, 16), 0); tb = max(min(tb, 16), 0);
		// mask keeping the top `t` symbols of a word (first symbol lives in bits 31:30)
		u32 ma = __funnelshift_rc(0u, 0xffffffffu, 2 * ta), mb = __funnelshift_rc(0u, 0xffffffffu, 2 * tb);
		u32 va = b.w[8 + j] & ma, vb = b.w[9 + j] & mb;
		u32 ha = (va >> 1) & 0x55555555u, la = va & 0x55555555u;
		u32 hb = vb & 0xaaaaaaaau, lb = (vb << 1) & 0xaaaaaaaau;
		nh += __popc(ha | hb);
		nl += __popc(la | lb);
		nt += __popc((ha & la) | (hb & lb));
	}
	Cnt4 o;
	o.c[3] = nt; o.c[2] = nh - nt; o.c[1] = nl - nt; o.c[0] = r - nh - nl + nt;
	return o;
}

__device__ __forceinline__ u64 blk_base(const OccBlock &b, int c) { return (u64)b.w[2 * c] | ((u64)b.w[2 * c + 1] << 32); }

// Result of extending an interval by base c (only the chosen base is materialised).
struct Ext { u64 a, b, s; };   // a = x[!is_back], b = x[is_back], s = x[2]

// bwt_extend for one chosen base c: in (a = x[!is_back], b = x[is_back], s = x[2]).
__device__ __forceinline__ Ext extend(const DevIndex &ix, u64 a, u64 b, u64 s, int c, u64 hot_min, u64 policy)
{
	u64 k = a - 1, l = a - 1 + s;
	u64 kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);   // '$' is not stored (bwt.c:194)
	u64 bk = kk >> 7, bl = ll >> 7;
	OccBlock K, L;
	const bool same = bk == bl;
	if (hot_min && s >= hot_min) {
		ld_block_hot(K, ix.blk + bk * 4, policy);
		if (!same) ld_block_hot(L, ix.blk + bl * 4, policy);
	} else {
		ld_block(K, ix.blk + bk * 4);
		if (!same) ld_block(L, ix.blk + bl * 4);
	}
	Cnt4 ck = occ_in_block(K, (u32)(kk & 127) + 1);
	if (same) L = K;
	Cnt4 cl = occ_in_block(L, (u32)(ll & 127) + 1);
	u64 sz[4], tk_c = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
		u64 tk = blk_base(K, j) + ck.c[j], tl = blk_base(L, j) + cl.c[j];
		sz[j] = tl - tk;
		if (j == c) tk_c = tk;
	}
	Ext o;
	// other strand: bases are laid out T,G,C,A after the optional '$' (bwt.c:425-428)
	u64 acc = b + ((a <= ix.primary && a + s - 1 >= ix.primary) ? 1 : 0);
	if (c < 3) acc += sz[3];
	if (c < 2) acc += sz[2];
	if (c < 1) acc += sz[1];
	o.b = acc;
	o.s = c == 0 ? sz[0] : c == 1 ? sz[1] : c == 2 ? sz[2] : sz[3];
	o.a = ix.L2[c] + 1 + tk_c;
	return o;
}
