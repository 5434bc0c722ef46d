// smem_repeat.cuh -- the repeat filter of the re-seeding pass (sm_100a): table, builder kernels, device-side test.
//
// smem_next2 (bwamem.c:244-305) re-seeds from the middle of a long, (nearly) unique SMEM: a second bwt_smem1 at
// x = (start + end) >> 1 with min_intv = x[2] + 1 >= 2 (bwamem.c:272-278), about a third of all bwt_extend calls of a
// 101 bp read.  Its result only enters the merged list through the test of bwamem.c:288,297: length >= max >> 1 (max =
// length of that longest SMEM) and end > ori_start.  Every entry of the second pass is a pattern q[s..e) with s <= x < e
// that occurs >= min_intv >= 2 times in the indexed text T; a pattern of length >= K through x contains a K-mer window
// q[a..a+K), x-K+1 <= a <= x, and a substring occurs at least as often as the pattern.  Hence:
//
//     if max >> 1 >= K and every K-mer window of the read through x occurs at most ONCE in T, the second pass cannot
//     contribute an entry that survives the merge, and smem_next2 returns its pass-1 list unchanged
//
// (itr->start is set from pass 1 only, bwamem.c:262).  The kernel then skips the pass: one test instead of ~150
// extends.  "Occurs at most once" is answered by a bit table over a hash of the K-mer: bit set <=> some K-mer with that
// hash occurs more than once in T (built once per index, below); a collision can only set a bit, i.e. cost a skip, never
// cause a wrong one.  The model of this rule with exact counts is in oracle/smem_oracle.c (pass2_is_void) and is checked
// against the unmodified algorithm on the CPU (tests/test_repeat_filter.py).  TRACE / SMEM1 modes never skip: they
// return raw bwt_smem1 lists.
#pragma once
#include "smem_device.cuh"

#define RF_MULT_BIT  0x9E3779B97F4A7C15ull     // bit index   = (code * RF_MULT_BIT) >> (64 - log2_bits)
#define RF_MULT_SLOT 0xC2B2AE3D27D4EB4Full     // builder hash-table slot
#define RF_MULT_GRP  0xD6E8FEB86659FD93ull     // builder pass (group) of a k-mer
#define RF_EMPTY 0xffffffffffffffffull

__host__ __device__ __forceinline__ u64 rf_bit_index(u64 code, int log2_bits) { return (code * RF_MULT_BIT) >> (64 - log2_bits); }

// T = forward + reverse complement, 32 bases per 64-bit word, first base in the top bits, zero padded -- the builder's view
// of the text.  `pac` is the reference's .pac layout of the forward strand (bntseq.c: four bases per byte, first base in
// the top bits); the reverse complement is appended the way bntseq.c:268-273 does.
__global__ void __launch_bounds__(256) pack_text_kernel(const uint8_t *__restrict__ pac, long long l_pac, u64 *__restrict__ tw, long long n_words)
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

// K-mer code at text position pos (tw: 32 bases per 64-bit word, first base in the top bits, zero padded by >= 1 word)
__device__ __forceinline__ u64 rf_text_code(const u64 *__restrict__ tw, long long pos, int K)
{
	const u64 w0 = tw[pos >> 5], w1 = tw[(pos >> 5) + 1];
	const int o = 2 * (int)(pos & 31);
	const u64 W = o ? (w0 << o) | (w1 >> (64 - o)) : w0;
	return W >> (64 - 2 * K);
}

// One pass of the builder: the K-mers of group g (of G) go into an open-addressing table; a K-mer that is already there
// occurs more than once -> its bit is set.  Distinct K-mers per group <= positions / G, so the table never fills
// (the host sizes G for a load <= 1/2).
__global__ void __launch_bounds__(256) rf_insert_kernel(const u64 *__restrict__ tw, long long n, int K, u32 G, u32 g, u64 *__restrict__ slots,
                                                        int log2_slots, u32 *__restrict__ bits, int log2_bits)
{
	const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (pos + K > n) return;
	const u64 code = rf_text_code(tw, pos, K);
	if ((u32)(((code * RF_MULT_GRP) >> 32) % G) != g) return;
	const u64 mask = (1ull << log2_slots) - 1;
	u64 s = (code * RF_MULT_SLOT) >> (64 - log2_slots);
	for (;;) {
		const u64 old = atomicCAS(&slots[s], RF_EMPTY, code);
		if (old == RF_EMPTY) return;
		if (old == code) { const u64 b = rf_bit_index(code, log2_bits); atomicOr(&bits[b >> 5], 1u << (b & 31)); return; }
		s = (s + 1) & mask;
	}
}

// The bit index is the TOP log2_bits bits of a product, so the table for log2_bits - 1 is the OR of adjacent bit pairs of
// the table for log2_bits: the builder fills a large table once and folds it down until the share of set bits would pass
// 1/256 -- small enough for the L2 on a repeat-poor text, large on a repeat-rich one.  `pop` receives the popcount of dst.
__global__ void __launch_bounds__(256) rf_fold_kernel(const u32 *__restrict__ src, u32 *__restrict__ dst, long long n_dst_words, unsigned long long *pop)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	u32 out = 0;
	if (i < n_dst_words) {
		u64 v = (u64)src[2 * i] | ((u64)src[2 * i + 1] << 32);
		v = (v | (v >> 1)) & 0x5555555555555555ull;                 // pair ORs on the even bits, then pack the even bits
		v = (v | (v >> 1)) & 0x3333333333333333ull;
		v = (v | (v >> 2)) & 0x0f0f0f0f0f0f0f0full;
		v = (v | (v >> 4)) & 0x00ff00ff00ff00ffull;
		v = (v | (v >> 8)) & 0x0000ffff0000ffffull;
		v = (v | (v >> 16)) & 0x00000000ffffffffull;
		out = (u32)v;
		dst[i] = out;
	}
	u32 c = __popc(out);
#pragma unroll
	for (int m = 16; m >= 1; m >>= 1) c += __shfl_xor_sync(0xffffffffu, c, m);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(pop, (unsigned long long)c);
}

// Per-read window flags, computed by pack_reads_kernel right before the seed kernel: bit a of a read's flag words is SET
// when the K-mer window starting at base a cannot be vouched for -- it runs past the end of the read, holds an ambiguous
// base, or its bit in the table is set.  The seed kernel's test is then two word loads and a mask (rf_range_is_clear):
// no table access, no extra registers inside the seeding kernel.  (A first version looked the windows up from inside
// the seed kernel's cold section: the out-of-line call cost spills in the main loop and every test stalled the warp for
// two DRAM round trips.)
__device__ __forceinline__ bool rf_range_is_clear(const u32 *__restrict__ flags, int K, int len, int x)
{
	const int lo = max(x - K + 1, 0), hi = min(x, len - K);      // window starts through x; hi - lo < K <= 32
	if (hi < lo) return false;
	const u32 w0 = flags[lo >> 5], w1 = flags[hi >> 5];
	const u32 m0 = 0xffffffffu << (lo & 31), m1 = 0xffffffffu >> (31 - (hi & 31));
	const u32 bad = (lo >> 5) == (hi >> 5) ? (w0 & m0 & m1) : ((w0 & m0) | (w1 & m1));
	return bad == 0;
}
