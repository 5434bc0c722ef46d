// smem_device.cuh -- device side of the SMEM seeding path (sm_100a).
//
// What is computed (bit-exact with the reference's CPU code, see oracle/smem_oracle.c):
//   occ_half      <- bwt_occ4      bwt.c:187-204
//   extend_pair   <- bwt_extend    bwt.c:416-429 (bwt_2occ4 = two occ4, bwt.c:213-214)
//   seed_kernel   <- bwt_smem1     bwt.c:776-835, smem_next2 bwamem.c:244-305,
//                    enumeration loop of mem_insert_seed bwamem.c:453-460
//
// Hardware facts this design follows (measured with smem_gpu_gather_roofline):
//   * random gathers over the 3.1 GB index are bound by the NUMBER of DRAM-missing requests
//     (~38 G/s on a B200) and not by their size up to 128 B: a 64-byte occ block fetched as two
//     per-lane 32-byte loads costs two requests (1.23 TB/s), fetched by two adjacent lanes in one
//     instruction it costs one (2.45 TB/s);
//   * so one read is carried by a PAIR of lanes, each loading one 32-byte sector of the same
//     64-byte block in the same LDG.256: one request per block, one or two blocks per bwt_extend.
//
// Device layout of the index ("split bit-plane" block, produced by repack_kernel at upload from the
// unmodified bwt_t::bwt): per 128 symbols still one 64-byte block, but arranged so that both lanes of
// a pair do half of the popcount work and nothing has to be shifted into place:
//   sector h (h = 0,1) = { uint64 cnt[2h], uint64 cnt[2h+1],      checkpoints of bases 2h, 2h+1
//                          uint32 hi[2], uint32 lo[2] }            bit planes of symbols 64h .. 64h+63
//   (symbol i of the half lives in bit 31-(i&31) of word i>>5; hi = bit 1, lo = bit 0 of the 2-bit code).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

typedef unsigned long long u64;
typedef unsigned int u32;

#define FULL_MASK 0xffffffffu

struct DevIndex {
	const uint4 *blk;   // re-packed 64-byte occ blocks in HBM
	const uint4 *sec;   // the same index as self-contained 32-byte sectors (one lane per read, see extend_single); nullptr = not built
	u64 primary;
	u64 L2[5];
	u64 seq_len;
};

struct __align__(32) Intv { u64 x0, x1, x2, info; };

struct SeedParams {
	DevIndex ix;
	const int *rlen;         // [n] read lengths (pack pre-pass)
	const uint4 *qpack;      // reads re-packed two bases per byte (0..3, > 3 ambiguous; padded with 4), q_stride bytes per read (pack pre-pass)
	long long n;
	const int *list;         // optional indirection (overflow re-run): read id = list[k]
	const int *xs, *min_intvs; // smem1 mode inputs
	int *ret;                // smem1 mode output
	Intv *slots;             // [n][slot_cap] per-read result slots
	int slot_cap;
	int *counts;             // [n] exact interval count per read (even when > slot_cap)
	int *overflow_list;      // read ids with count > slot_cap
	int *status;             // [0] work counter, [1] n_overflow, [2] guard trips, [3] largest overflowing count, [5] bad-input flags of the pack pre-pass, [6] skipped re-seeding passes, [7] unique walks
	Intv *scratch;           // per lane pair: M1, M2, BX arrays of scratch_cap entries each (global, L2-resident)
	int scratch_cap;
	int b_cap;               // entries of the prev/curr array kept in shared memory per read (rest spills to BX)
	int q_stride;            // bytes of shared memory per staged query (two bases per byte)
	int pair_stride;         // bytes of shared memory per lane pair
	int split_len_init, split_width, start_width;
	// repeat filter (smem_repeat.cuh): bit hash(w) is set for every rf_k-mer w that occurs more than once in the indexed text
	const u32 *qflags;       // per read, per 32 window starts: bit set = that rf_k-mer window is not vouched for (pack_reads_kernel); nullptr = no filter
	int rf_k;                // k-mer length (<= 32)
	// unique-walk tables (smem_kernels.cuh PH_UW_*): text at one base per nibble (eight bases per 32-bit word, first base lowest, like the staged reads), the
	// full suffix array (row -> text position) and the inverse sampled every 2^uw_isa_shift positions from the end of the text, both as 33-bit
	// entries, seven to a 32-byte sector (seven low words + one word of high bits); nullptr = off
	const uint4 *uw_text; const u32 *uw_fsa, *uw_isa; int uw_isa_shift;
	int uw_min_run, uw_min_left;   // a walk starts after this many extends of an interval of size 1 and with at least this many read bases left (it costs three gathers)
	int spec_walk;           // 1 = pass-1 calls that directly follow another walk their longest candidate back alone first (PH_SPEC)
	int count_skips;         // debug: status[6] counts the re-seeding passes the filter proved void, status[7] the unique walks
	u64 hot_min_intv;        // 0 = off; occ blocks of intervals >= this size are "hot" (shallow levels, re-used across reads)
	int l2_mode;             // L2 eviction hints when hot_min_intv != 0: 0 = hot evict_last / cold normal,
	                         // 1 = hot normal / cold evict_first, 2 = hot evict_last / cold evict_first
};

// ---------------------------------------------------------------------------------------------
// memory helpers

// One 32-byte sector of an occ block per lane: 256-bit load (new on sm_100), read-only path, no L1
// allocation (a block is never re-used by the same SM before eviction).
__device__ __forceinline__ void ld_sector(u32 (&w)[8], const uint4 *p)
{
	asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p));
}
__device__ __forceinline__ void ld_sector_hot(u32 (&w)[8], const uint4 *p, u64 policy)
{
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
	             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7])
	             : "l"(p), "l"(policy));
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

// Shared memory through 32-bit shared-space addresses (keeps the generic->shared window arithmetic
// out of the hot loop).
__device__ __forceinline__ u32 lds_u8(u32 a) { u32 v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void sts_u8(u32 a, u32 v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ int lds_i32(u32 a) { int v; asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ int lds_u16(u32 a) { u32 v; asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a)); return (int)v; }
__device__ __forceinline__ void sts_u16(u32 a, int v) { asm volatile("st.shared.u16 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_i32(u32 a, int v) { asm volatile("st.shared.s32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ uint4 lds_v4(u32 a)
{
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
	return v;
}
__device__ __forceinline__ void sts_v4(u32 a, uint4 v)
{
	asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// prev/curr entries in shared memory: 16 bytes when every coordinate fits 36 bits (seq_len < 2^36,
// i.e. any genome bwa can index in practice) and the read is shorter than 2^20, else 32 bytes.
template <bool WIDE> struct BEntry;
template <> struct BEntry<false> {
	static const int BYTES = 16;
	static __device__ __forceinline__ void put(u32 addr, u64 x0, u64 x1, u64 x2, u32 end)
	{
		sts_v4(addr, make_uint4((u32)x0, (u32)x1, (u32)x2,
		                        ((u32)(x0 >> 32) & 15u) | (((u32)(x1 >> 32) & 15u) << 4) | (((u32)(x2 >> 32) & 15u) << 8) | (end << 12)));
	}
	static __device__ __forceinline__ void get(u32 addr, u64 &x0, u64 &x1, u64 &x2, u32 &end)
	{
		const uint4 t = lds_v4(addr);
		x0 = (u64)t.x | ((u64)(t.w & 15u) << 32);
		x1 = (u64)t.y | ((u64)((t.w >> 4) & 15u) << 32);
		x2 = (u64)t.z | ((u64)((t.w >> 8) & 15u) << 32);
		end = t.w >> 12;
	}
};
template <> struct BEntry<true> {
	static const int BYTES = 32;
	static __device__ __forceinline__ void put(u32 addr, u64 x0, u64 x1, u64 x2, u32 end)
	{
		sts_v4(addr, make_uint4((u32)x0, (u32)(x0 >> 32), (u32)x1, (u32)(x1 >> 32)));
		sts_v4(addr + 16, make_uint4((u32)x2, (u32)(x2 >> 32), end, 0u));
	}
	static __device__ __forceinline__ void get(u32 addr, u64 &x0, u64 &x1, u64 &x2, u32 &end)
	{
		const uint4 lo = lds_v4(addr), hi = lds_v4(addr + 16);
		x0 = (u64)lo.x | ((u64)lo.y << 32);
		x1 = (u64)lo.z | ((u64)lo.w << 32);
		x2 = (u64)hi.x | ((u64)hi.y << 32);
		end = hi.z;
	}
};

// ---------------------------------------------------------------------------------------------
// occ: counts of A,C,G,T among the first r (0..64) symbols of this lane's half block, packed one
// byte per base.  w[4..5] = hi plane, w[6..7] = lo plane.  6 POPC, no shifting of data.
__device__ __forceinline__ u32 occ_half(const u32 (&w)[8], int r)
{
	const u32 m0 = __funnelshift_rc(0u, 0xffffffffu, r);                 // top min(r,32) bits
	const u32 m1 = __funnelshift_rc(0u, 0xffffffffu, max(r - 32, 0));
	const u32 h0 = w[4] & m0, h1 = w[5] & m1, l0 = w[6] & m0, l1 = w[7] & m1;
	const u32 nh = __popc(h0) + __popc(h1), nl = __popc(l0) + __popc(l1), nt = __popc(h0 & l0) + __popc(h1 & l1);
	// T = hi&lo, G = hi&~lo, C = ~hi&lo, A = the rest of the r symbols
	return ((u32)r - nh - nl + nt) | ((nl - nt) << 8) | ((nh - nt) << 16) | (nt << 24);
}

struct Ext { u64 a, b, s; };   // a = x[!is_back], b = x[is_back], s = x[2]

// bwt_extend for one chosen base c, executed by a converged warp of lane pairs (half = lane & 1).
// in: (a = x[!is_back], b = x[is_back], s = x[2], c) identical in both lanes of a pair; out likewise.
// w / v receive this lane's sector of the K / L block.  (Keeping them across calls to skip the gather when the block
// number repeats was measured: +14 % at equal occupancy, a tie once its registers are paid for -- removed.)
// NARROW: every base occurs fewer than 2^32 times in the text (checked at upload), so occurrence counts and the sizes of
// the four child intervals are 32-bit quantities: half the shuffles and none of the carries of the general form.
template <bool NARROW>
__device__ __forceinline__ Ext extend_pair(const DevIndex &ix, u64 a, u64 b, u64 s, int c, int half, int lane, u64 hot_min, u64 pol_hot, u64 pol_cold,
                                           u32 (&w)[8], u32 (&v)[8], const uint4 *alt = nullptr)
{
	const u64 k = a - 1, l = a - 1 + s;
	const u64 kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);   // '$' is not stored (bwt.c:194)
	const u64 bk = kk >> 7, bl = ll >> 7;
	const bool same = bk == bl;
	// `alt` (unique-walk phases): this lane's gather goes to another table instead; the pair rides on the interval (1,1,1)
	const uint4 *pk = alt ? alt : ix.blk + bk * 4 + half * 2, *pl = ix.blk + bl * 4 + half * 2;
	if (hot_min) {                       // uniform branch; ONE load site, the policy is a per-lane operand
		const u64 policy = s >= hot_min ? pol_hot : pol_cold;
		ld_sector_hot(w, pk, policy);
		if (!same) ld_sector_hot(v, pl, policy);
	} else {
		ld_sector(w, pk);
		if (!same) ld_sector(v, pl);
	}
	if (same) {
#pragma unroll
		for (int j = 0; j < 8; ++j) v[j] = w[j];
	}
	// symbols 0..kk&127 (inclusive) of the block count; this lane owns symbols 64*half .. 64*half+63
	const int rk = min(max((int)(kk & 127) + 1 - 64 * half, 0), 64), rl = min(max((int)(ll & 127) + 1 - 64 * half, 0), 64);
	u32 ck = occ_half(w, rk), cl = occ_half(v, rl);
	ck += __shfl_xor_sync(FULL_MASK, ck, 1);          // whole-block counts (each byte <= 128)
	cl += __shfl_xor_sync(FULL_MASK, cl, 1);
	// this lane owns bases j0 = 2*half and j0 + 1
	const int j0 = 2 * half;
	if (NARROW) {
		const u32 tk0 = w[0] + ((ck >> (8 * j0)) & 0xffu), tk1 = w[2] + ((ck >> (8 * j0 + 8)) & 0xffu);
		const u32 tl0 = v[0] + ((cl >> (8 * j0)) & 0xffu), tl1 = v[2] + ((cl >> (8 * j0 + 8)) & 0xffu);
		const u32 sz0 = tl0 - tk0, sz1 = tl1 - tk1;
		u64 acc = (u64)(j0 > c ? sz0 : 0u) + (u64)(j0 + 1 > c ? sz1 : 0u);
		acc += __shfl_xor_sync(FULL_MASK, acc, 1);
		const u32 tkc = (c & 1) ? tk1 : tk0, szc = (c & 1) ? sz1 : sz0;
		const int owner = (lane & ~1) | (c >> 1);
		Ext o;
		o.a = ix.L2[c] + 1 + (u64)__shfl_sync(FULL_MASK, tkc, owner);
		o.s = (u64)__shfl_sync(FULL_MASK, szc, owner);
		o.b = b + ((a <= ix.primary && a + s - 1 >= ix.primary) ? 1 : 0) + acc;
		return o;
	}
	const u64 tk0 = ((u64)w[0] | ((u64)w[1] << 32)) + ((ck >> (8 * j0)) & 0xffu), tk1 = ((u64)w[2] | ((u64)w[3] << 32)) + ((ck >> (8 * j0 + 8)) & 0xffu);
	const u64 tl0 = ((u64)v[0] | ((u64)v[1] << 32)) + ((cl >> (8 * j0)) & 0xffu), tl1 = ((u64)v[2] | ((u64)v[3] << 32)) + ((cl >> (8 * j0 + 8)) & 0xffu);
	const u64 sz0 = tl0 - tk0, sz1 = tl1 - tk1;
	// other strand: bases are laid out T,G,C,A after the optional '$' (bwt.c:425-428): sum the sizes of bases > c
	u64 acc = (j0 > c ? sz0 : 0) + (j0 + 1 > c ? sz1 : 0);
	acc += __shfl_xor_sync(FULL_MASK, acc, 1);
	const u64 tkc = (c & 1) ? tk1 : tk0, szc = (c & 1) ? sz1 : sz0;       // valid in the lane that owns base c
	const int owner = (lane & ~1) | (c >> 1);
	Ext o;
	o.a = ix.L2[c] + 1 + __shfl_sync(FULL_MASK, tkc, owner);
	o.s = __shfl_sync(FULL_MASK, szc, owner);
	o.b = b + ((a <= ix.primary && a + s - 1 >= ix.primary) ? 1 : 0) + acc;
	return o;
}

// ---------------------------------------------------------------------------------------------
// One LANE per read (seed_kernel<..., LPR = 1>).  Instruction issue, not the DRAM request rate, bounds the lane-pair kernel, and
// two thirds of its instructions are the per-read state machine that both lanes of a pair execute in duplicate.  When every
// base occurs fewer than 2^32 times the checkpoints fit 32 bits, and 64 symbols with their own four checkpoints fill exactly
// one 32-byte sector:
//   sector m = { uint32 cnt[4]      occurrences of A,C,G,T in symbols [0, 64 m)
//                uint32 hi[2], lo[2] bit planes of symbols 64 m .. 64 m + 63 (as in the 64-byte block) }
// (same 0.5 byte per symbol; built from the 64-byte blocks by sectors_from_blocks_kernel).  A lane then computes bwt_occ4 from
// ONE 32-byte load = one DRAM request -- the request ceiling is per request, not per byte -- with no shuffles, and a warp
// carries 32 reads instead of 16.
__device__ __forceinline__ Ext extend_single(const DevIndex &ix, u64 a, u64 b, u64 s, int c, u32 (&w)[8], u32 (&v)[8],
                                             const uint4 *alt = nullptr, const uint4 *alt2 = nullptr)
{
	const u64 k = a - 1, l = a - 1 + s;
	const u64 kk = k - (k >= ix.primary), ll = l - (l >= ix.primary);   // '$' is not stored (bwt.c:194)
	const u64 mk = kk >> 6, ml = ll >> 6;
	const bool same = mk == ml && !alt2;
	// `alt` / `alt2` (unique-walk phases): the gathers go to another table instead; the lane rides on the interval (1,1,1)
	const uint4 *pk = alt ? alt : ix.sec + mk * 2, *pl = alt2 ? alt2 : ix.sec + ml * 2;
	ld_sector(w, pk);
	if (!same) ld_sector(v, pl);
	if (same) {
#pragma unroll
		for (int j = 0; j < 8; ++j) v[j] = w[j];
	}
	const u32 ck = occ_half(w, (int)(kk & 63) + 1), cl = occ_half(v, (int)(ll & 63) + 1);   // symbols 0 .. kk & 63 of the sector count
	const u32 tk0 = w[0] + (ck & 0xffu), tk1 = w[1] + ((ck >> 8) & 0xffu), tk2 = w[2] + ((ck >> 16) & 0xffu), tk3 = w[3] + (ck >> 24);
	const u32 sz0 = v[0] + (cl & 0xffu) - tk0, sz1 = v[1] + ((cl >> 8) & 0xffu) - tk1, sz2 = v[2] + ((cl >> 16) & 0xffu) - tk2, sz3 = v[3] + (cl >> 24) - tk3;
	// other strand: bases are laid out T,G,C,A after the optional '$' (bwt.c:425-428): sum the sizes of bases > c
	const u64 acc = (u64)(c < 1 ? sz1 : 0u) + (u64)(c < 2 ? sz2 : 0u) + (u64)(c < 3 ? sz3 : 0u);
	const u32 tkc = c == 0 ? tk0 : c == 1 ? tk1 : c == 2 ? tk2 : tk3;
	const u32 szc = c == 0 ? sz0 : c == 1 ? sz1 : c == 2 ? sz2 : sz3;
	Ext o;
	o.a = ix.L2[c] + 1 + (u64)tkc;
	o.s = (u64)szc;
	o.b = b + ((a <= ix.primary && a + s - 1 >= ix.primary) ? 1 : 0) + acc;
	return o;
}
