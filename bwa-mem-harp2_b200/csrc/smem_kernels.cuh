// smem_kernels.cuh -- the seeding kernel and its small companions (sm_100a).
#pragma once
#include "smem_device.cuh"
#include "smem_repeat.cuh"

enum { MODE_COLLECT = 0, MODE_SMEM1 = 1, MODE_TRACE = 2 };   // TRACE: COLLECT's walk, but every bwt_smem1 call's raw list is kept
// hot phases first: the main loop only ever extends in PH_FWD / PH_BWD / PH_SPEC
enum { PH_FWD = 0, PH_BWD = 1, PH_SPEC = 2 /* backward walk of the longest candidate alone */,
       PH_UW_SA = 3, PH_UW_TEXT = 4, PH_UW_ISA = 5, PH_UW_LF = 6 /* unique forward walk by text comparison */, PH_IDLE = 7, PH_NEED_READ = 8, PH_NEXT_STEP, PH_INIT_CALL, PH_FWD_END, PH_FWD_DONE, PH_BWD_LAST, PH_CALL_DONE };

#define STEP_SHIFT 48        // inside the slots the smem_next2 step index (TRACE: step*2 + pass) rides in info bits 48..63
#define AUX_SHIFT 16         // ... and, in TRACE mode, bwt_smem1's return value in bits 16..31 (query positions are < 2^16)
#define INFO_MASK 0x0000ffff0000ffffull
#ifndef SEED_BLOCK
#define SEED_BLOCK 128       // threads per CTA of the lane-pair seeding kernel (= 64 lane pairs = 64 reads in flight)
#endif
#ifndef SEED_BLOCK1
#define SEED_BLOCK1 64       // threads per CTA of the one-lane-per-read kernel (= 64 reads in flight; nine CTAs per SM by shared memory)
#endif

// per-pair shared memory: [B entries | cold state | query, two bases per byte]
// cold state: 32-bit rk, rid, n_out, max_s (2 words); everything else is a query position or a small count (16 bit)
enum { CS_RK = 0, CS_RID = 4, CS_NOUT = 8, CS_MAXS_LO = 12, CS_MAXS_HI = 16,
       CS_START = 20, CS_STEP = 22, CS_ORI = 24, CS_SPLIT = 26, CS_PASS = 28, CS_X = 30, CS_NMEM = 32, CS_LMS = 34, CS_NM1 = 36,
       CS_KEEP = 38, CS_MAXLEN = 40, CS_MAXSTART = 42, CS_MAXEND = 44, CS_RET = 46, COLD_BYTES = 48 };

// Shared-memory address of byte `off` of a read's private area.  Lane pairs (LPR 2): the area is contiguous at sp.  One lane per
// read (LPR 1): the 32 areas of a warp are interleaved in 16-byte units (unit u of lane t at warp base + 512 u + 16 t, sp = warp
// base + 16 t), so that the same unit of 32 reads covers all banks: data-dependent indices differ per lane, the bank does not.
template <int LPR> __device__ __forceinline__ u32 sh_at(u32 sp, u32 off) { return LPR == 2 ? sp + off : sp + ((off >> 4) << 9) + (off & 15u); }
__device__ __noinline__ void bx_put(Intv *p, u64 x0, u64 x1, u64 x2, u32 end) { st_intv(p, x0, x1, x2, (u64)end); }
__device__ __noinline__ Intv bx_get(const Intv *p) { return ld_intv(p); }

// One persistent LANE PAIR = one read at a time (see smem_device.cuh for why a pair).  Both lanes run
// the same state machine on identical state; they differ only in which sector of an occ block they
// load and which two bases they account for.  Every trip of the main loop performs exactly one
// bwt_extend for every pair of the warp, whatever read / pass / direction each pair is in; rare
// transitions (next read, next smem_next2 step, call set-up, merge) run in a divergent cold section
// whose state lives in shared memory so that the hot loop stays small.
//
// Per-read storage:
//   shared : cold state, q[len] (the staged query), B[b_cap] = prev/curr of bwt_smem1 compacted IN
//            PLACE (curr[n] is written at or behind the prev[j] just consumed, so one array serves both,
//            bwt.c:810-828) and addressed from its top so that "reverse curr" (bwt.c:807) costs nothing;
//   global : M1, M2 = `matches` / `sub` of smem_next2 in emission order (descending start), each of
//            scratch_cap = max_read_len + 2 entries, which bounds every list of bwt.c:776-835.
//            BX = the entries of B beyond b_cap (only the deepest forward passes reach them).
//
// LPR = lanes per read: 2 = the lane pair described above (any index); 1 = one lane per read on the 32-byte sector form of the
// index (extend_single, smem_device.cuh; needs 32-bit occurrence counts, i.e. !WIDE) -- the same state machine, but a warp
// carries 32 reads, nothing is executed in duplicate and nothing is shuffled.
// SEC (LPR 2 only): both lanes of the pair run extend_single on the same 32-byte sector -- identical addresses, still one request --
// instead of sharing a 64-byte block: no shuffles, a third fewer instructions in the extend.
// SPLIT (with SEC): in the backward sweep the two lanes of a pair extend TWO DIFFERENT entries of prev[] in the same trip (lane 0
// prev[j], lane 1 prev[j + 1]; each lane can compute a whole bwt_extend by itself on the sector form).  The entries of one round are
// independent of each other (bwt.c:812-825 only looks at the previous entry's SIZE to de-duplicate), so the pair exchanges the two
// sizes, both lanes replay the two push / emit decisions in order on identical state, and each lane stores its own entry.  Half the
// trips in the triangular sweeps at error positions, which are 40 % of all extends.
template <int MODE, int MIN_BLOCKS, bool WIDE, int LPR, bool SEC = false, bool SPLIT = false>
__global__ void __launch_bounds__(LPR == 2 ? SEED_BLOCK : SEED_BLOCK1, MIN_BLOCKS) seed_kernel(const SeedParams p)
{
	static_assert((LPR == 2 && !SEC) || !WIDE, "the 32-bit sector form of the index needs 32-bit occurrence counts");
	static_assert(!SPLIT || (SEC && LPR == 2), "splitting the backward sweep needs lane pairs on the sector form");
	const bool SPEC = MODE != MODE_SMEM1 && p.spec_walk;     // speculative longest-only backward walk (PH_SPEC)
	typedef BEntry<WIDE> BE;
	extern __shared__ uint4 smem_raw[];
	const int lane = threadIdx.x & 31, half = LPR == 2 ? lane & 1 : 0;
	const int pair = LPR == 2 ? threadIdx.x >> 1 : threadIdx.x;      // the read slot of this CTA the thread works for
	const u32 pm2 = LPR == 2 ? 3u << (lane & ~1) : 1u << lane;       // the lanes that carry this read
	u32 sp = (u32)__cvta_generic_to_shared(smem_raw) +
	         (LPR == 2 ? (u32)pair * (u32)p.pair_stride : (u32)(threadIdx.x >> 5) * 32u * (u32)p.pair_stride + (u32)lane * 16u);   // this read's shared memory (sh_at)
#ifdef SEED_KEEP_SP
	asm volatile("" : "+r"(sp));                             // opaque: kept in its register instead of being re-derived from the thread id at every use
#endif
	// cold state first, then B, then the query: the cold-state and B addresses are sp + a compile-time constant
	auto SC = [&](u32 off) -> u32 { return sh_at<LPR>(sp, off); };                             // cold state
	auto SB = [&](int idx) -> u32 { return LPR == 2 ? sp + COLD_BYTES + (u32)idx * BE::BYTES : sp + ((u32)(COLD_BYTES / 16 + idx) << 9); };   // B entries (16-byte aligned)
	const u32 sq = sh_at<LPR>(sp, COLD_BYTES + (u32)p.b_cap * BE::BYTES);                      // query, two bases per byte (a multiple of 16 bytes from sp)
	auto QBYTE = [&](u32 byte) -> u32 { return LPR == 2 ? sq + byte : sq + ((byte >> 4) << 9) + (byte & 15u); };
	auto qbase = [&](int i) -> u32 { return (lds_u8(QBYTE((u32)i >> 1)) >> ((i & 1) * 4)) & 15u; };
#ifdef SEED_KEEP_GP
	u32 gp3 = (u32)(blockIdx.x * ((LPR == 2 ? SEED_BLOCK : SEED_BLOCK1) / LPR) + pair) * 3u;
	asm volatile("" : "+r"(gp3));                            // opaque, one register: the scratch base is two multiply-adds away instead of a dozen instructions
	auto scratch_base = [&]() -> Intv * { return p.scratch + (size_t)gp3 * (size_t)p.scratch_cap; };
#else
	Intv *const M1s = p.scratch + (size_t)(blockIdx.x * ((LPR == 2 ? SEED_BLOCK : SEED_BLOCK1) / LPR) + pair) * 3 * p.scratch_cap;
	auto scratch_base = [&]() -> Intv * { return M1s; };   // (forcing this out of the main loop with volatile reads cost spills: slower)
#endif

	u64 pol_hot = 0, pol_cold = 0;
	if (p.hot_min_intv) {
		if (p.l2_mode == 1) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_hot));
		else asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_hot));
		if (p.l2_mode == 0) asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol_cold));
		else asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_cold));
	}

	// B beyond b_cap spills to global memory through an out-of-line slow path: one compare + a branch that is
	// (almost) never taken on the hot path, and correctness never depends on b_cap
	auto b_put = [&](int idx, u64 x0, u64 x1, u64 x2, u32 end) {
		if (__builtin_expect(idx < p.b_cap, 1)) BE::put(SB(idx), x0, x1, x2, end);
		else bx_put(scratch_base() + 2 * p.scratch_cap + idx, x0, x1, x2, end);
	};
	auto b_get = [&](int idx, u64 &x0, u64 &x1, u64 &x2, u32 &end) {
		if (__builtin_expect(idx < p.b_cap, 1)) BE::get(SB(idx), x0, x1, x2, end);
		else { const Intv t = bx_get(scratch_base() + 2 * p.scratch_cap + idx); x0 = t.x0; x1 = t.x1; x2 = t.x2; end = (u32)t.info; }
	};
	// bwt.c:815-820: a hit that cannot be extended is recorded unless a longer match already covers it
	auto emit = [&](u64 x0, u64 x1, u64 x2, u32 end, int st) {
		const int n_mem = lds_u16(SC(CS_NMEM));
		if (n_mem == 0 || st < lds_u16(SC(CS_LMS))) {
			Intv *M = scratch_base() + (size_t)lds_u16(SC(CS_PASS)) * p.scratch_cap;
			st_intv(&M[n_mem], x0, x1, x2, (u64)end | ((u64)st << 32));
			sts_u16(SC(CS_NMEM), n_mem + 1);
			sts_u16(SC(CS_LMS), st);
			const int l = (int)end - st;
			if (l >= lds_u16(SC(CS_MAXLEN))) {          // ">=": the first maximum in ascending start order wins (bwamem.c:266-270)
				sts_u16(SC(CS_MAXLEN), l); sts_u16(SC(CS_MAXSTART), st); sts_u16(SC(CS_MAXEND), (int)end);
				sts_i32(SC(CS_MAXS_LO), (int)(u32)x2); sts_i32(SC(CS_MAXS_HI), (int)(u32)(x2 >> 32));
			}
		}
	};

	// ---- hot state (registers)
	int phase = PH_NEED_READ;
	u64 a = 1, b = 1, s = 1;             // FWD: ik as (a = x[1], b = x[0]); BWD: prev[j] as (a = x[0], b = x[1])
	u32 end = 0;                         // info of that interval (query end)
	int c = 0;                           // base of the pending extend
	int i = 0, j = 0, n0 = 0, n_prev = 0, n_curr = 0, len = 0, guard = 0;
	u32 min_intv = 1;                    // <= split_width + 1, an int (bwamem.c:272-278), or the caller's int32 (bwt.c:784)
	u64 last_s = 0;
	u32 blk_k[8], blk_l[8];              // this lane's sectors of the K / L occ blocks (see extend_pair)
#pragma unroll
	for (int t = 0; t < 8; ++t) blk_k[t] = blk_l[t] = 0;
	// One lane per read: a warp waits for every global round trip that any of its 32 reads makes in the cold section, so those are
	// taken off the critical path -- the next read's work index is drawn while the current read is being seeded (rk_pref), a new
	// read's length, bases and repeat-filter flags arrive in ONE round trip (flags kept in registers), and result lists are copied
	// four entries per round trip.
	int rk_pref = 0;
	u32 qf0 = 0, qf1 = 0, qf2 = 0, qf3 = 0;          // the read's window flags (repeat filter) when they fit four words
	// (drawing the work index ahead was also tried for the lane pairs: one more live register in the 56-register kernel, 18.8 -> 20.1 ms)
	constexpr bool PREF = LPR == 1;
#ifdef SEED_NO_SPLIT_COPY
	constexpr bool PCOPY = false;
#else
	constexpr bool PCOPY = SPLIT;
#endif
	if (PREF && !half) rk_pref = atomicAdd(&p.status[0], 1);
	// the entry of prev[] this lane extends next: prev[j] -- or, SPLIT in the backward sweep proper, prev[j + half] (a lane without an
	// entry rides on the interval (1,1,1)); every other phase needs prev[0] in both lanes
	auto load_prev = [&]() {
		if (SPLIT && phase == PH_BWD) {
			const int jj = j + half;
			if (jj < n_prev) b_get(n0 - 1 - jj, a, b, s, end); else { a = 1; b = 1; s = 1; }
		} else b_get(n0 - 1 - j, a, b, s, end);
	};

	for (;;) {
		__syncwarp();
		// ============================================================== cold section (divergent between pairs)
		// (the cases are laid out in the order a call runs through them and fall through where the next phase is known: one trip through the
		//  switch per transition instead of one per phase -- 17.95 -> 17.5 ms)
		while (phase > PH_IDLE) {
			switch (phase) {
			case PH_NEED_READ: {
				int rk = 0;
				if (LPR == 1) { rk = rk_pref; sts_i32(SC(CS_RK), rk); }
				else if (PREF) { rk = __shfl_sync(pm2, rk_pref, lane & ~1); if (!half) sts_i32(SC(CS_RK), rk); __syncwarp(pm2); }
				else {
					if (!half) { rk = atomicAdd(&p.status[0], 1); sts_i32(SC(CS_RK), rk); }
					__syncwarp(pm2); rk = lds_i32(SC(CS_RK));
				}
				if ((long long)rk >= p.n) { phase = PH_IDLE; break; }
				if (PREF && !half) rk_pref = atomicAdd(&p.status[0], 1);   // (needed at the next PH_NEED_READ, a few hundred extends from here)
				const int rid = p.list ? p.list[rk] : rk;
				if (LPR == 1) {
					// length, the first 128 bases and the window flags: issued together, one round trip
					const int cpr = p.q_stride >> 4;
					const uint4 *src = p.qpack + (size_t)rid * (size_t)cpr;
					const int len_ = p.rlen[rid];
					uint4 ch0 = __ldg(src), ch1 = ch0, ch2 = ch0, ch3 = ch0;
					if (cpr > 1) ch1 = __ldg(src + 1);
					if (cpr > 2) ch2 = __ldg(src + 2);
					if (cpr > 3) ch3 = __ldg(src + 3);
					if (p.qflags && cpr <= 4) {
						const u32 *fl = p.qflags + (size_t)rid * (size_t)cpr;
						qf0 = __ldg(fl);
						if (cpr > 1) qf1 = __ldg(fl + 1);
						if (cpr > 2) qf2 = __ldg(fl + 2);
						if (cpr > 3) qf3 = __ldg(fl + 3);
					}
					len = len_;
					sts_v4(sq, ch0);
					if (len > 32) sts_v4(sq + (1u << 9), ch1);
					if (len > 64) sts_v4(sq + (2u << 9), ch2);
					if (len > 96) sts_v4(sq + (3u << 9), ch3);
					for (int t0 = 4; 32 * t0 < len; t0 += 4) {               // longer reads: four chunks per round trip
						ch0 = __ldg(src + t0);
						if (32 * (t0 + 1) < len) ch1 = __ldg(src + t0 + 1);
						if (32 * (t0 + 2) < len) ch2 = __ldg(src + t0 + 2);
						if (32 * (t0 + 3) < len) ch3 = __ldg(src + t0 + 3);
						sts_v4(sq + ((u32)t0 << 9), ch0);
						if (32 * (t0 + 1) < len) sts_v4(sq + ((u32)(t0 + 1) << 9), ch1);
						if (32 * (t0 + 2) < len) sts_v4(sq + ((u32)(t0 + 2) << 9), ch2);
						if (32 * (t0 + 3) < len) sts_v4(sq + ((u32)(t0 + 3) << 9), ch3);
					}
				} else {
					{
						len = p.rlen[rid];
						// staged by pack_reads_kernel: 16 bytes (32 bases) per copy
						const uint4 *src = p.qpack + (size_t)rid * (size_t)(p.q_stride >> 4);
						for (int t = half; 32 * t < len; t += 2) sts_v4(sq + 16 * t, __ldg(src + t));
					}
					__syncwarp(pm2);
				}
				sts_i32(SC(CS_RID), rid); sts_i32(SC(CS_NOUT), 0); sts_u16(SC(CS_START), 0); sts_u16(SC(CS_STEP), 0);
				if (MODE != MODE_SMEM1) {
					sts_u16(SC(CS_SPLIT), p.split_len_init < len ? p.split_len_init : len);     // bwamem.c:458
					phase = PH_NEXT_STEP;
				} else {
					const int x = p.xs[rid], mi = p.min_intvs[rid];
					min_intv = mi < 1 ? 1u : (u32)mi;                                            // bwt.c:784
					sts_u16(SC(CS_PASS), 0); sts_u16(SC(CS_X), x);
					if (x < 0 || x >= len || qbase(x) > 3) {                                // bwt.c:783
						p.ret[rid] = x + 1; p.counts[rk] = 0; phase = PH_NEED_READ;
					} else phase = PH_INIT_CALL;
				}
			} break;
			case PH_FWD_END: {       // bwt.c:800-803 (ambiguous base) / :806 (end of read): push the last interval
				b_put(n_curr++, b, a, s, end);
				phase = PH_FWD_DONE;
			}
			// fall through: the next phase follows
			case PH_FWD_DONE: {      // bwt.c:807-809: "reverse curr" == address B from its top (n0 - 1 - j)
				n0 = n_curr; n_prev = n_curr; n_curr = 0;
				i = lds_u16(SC(CS_X)) - 1; j = 0;
				c = i < 0 ? -1 : (int)qbase(i);
				if (c > 3) c = -1;
				b_get(n0 - 1, a, b, s, end);             // prev[0] = the last push; its info is bwt_smem1's return value
				sts_u16(SC(CS_RET), (int)end);
				if (MODE != MODE_SMEM1 && lds_u16(SC(CS_PASS)) == 0) sts_u16(SC(CS_START), (int)end);   // bwamem.c:262
				phase = c < 0 ? PH_BWD_LAST : PH_BWD;
				// Every candidate of this sweep whose start reaches lim = CS_KEEP - 1 contains a pattern known to be too rare, so the
				// sweep ends there with everything dead; and while the LONGEST candidate (a, b, s) lives nothing is emitted
				// (bwt.c:815: curr->n != 0).  So walk it back alone: if it lives down to start lim + 1 (or to an ambiguous base /
				// the read start, where everything dies too) it is the call's only result and the shorter candidates never needed
				// extending; if it dies earlier the full sweep starts over (PH_SPEC in the main loop).  Model: oracle/smem_oracle.c.
				if (SPEC && phase == PH_BWD && lds_u16(SC(CS_PASS)) == 0 && lds_u16(SC(CS_KEEP)) > 0) {
					if (i < lds_u16(SC(CS_KEEP))) phase = PH_BWD_LAST;      // no round left before the limit: (i + 1, end) is the result
					else phase = PH_SPEC;
				}
				if (SPLIT && phase == PH_BWD && half) load_prev();          // lane 1 starts the sweep with prev[1]
			}
			if (phase != PH_BWD_LAST) break;      // else fall through: the next phase follows
			case PH_BWD_LAST: {
				// bwt.c:815-821 with c == -1 (read start or ambiguous base): nothing can enter curr and only
				// prev[0] can pass the containment test, so the round collapses to one emission test.
				emit(a, b, s, end, i + 1);
				phase = PH_CALL_DONE;
			}
			// fall through: the next phase follows
			case PH_CALL_DONE: {
				const int rk = lds_i32(SC(CS_RK)), n_mem = lds_u16(SC(CS_NMEM));
				Intv *const slot = p.slots + (size_t)rk * p.slot_cap;
				const Intv *const M1 = scratch_base();
				const Intv *const M2 = M1 + p.scratch_cap;
				if (MODE == MODE_SMEM1) {
					const int rid = lds_i32(SC(CS_RID));
					for (int e = n_mem - 1, o = 0; e >= 0; --e, ++o)             // bwt.c:829: ascending start
						if (o < p.slot_cap) { const Intv t = ld_intv(&M1[e]); st_intv(&slot[o], t.x0, t.x1, t.x2, t.info); }
					p.counts[rk] = n_mem; p.ret[rid] = lds_u16(SC(CS_RET));
					if (n_mem > p.slot_cap && !half) { p.overflow_list[atomicAdd(&p.status[1], 1)] = rid; atomicMax(&p.status[3], n_mem); }
					phase = PH_NEED_READ;
					break;
				}
				const int step = lds_u16(SC(CS_STEP));
				int n_out = lds_i32(SC(CS_NOUT));
				if (MODE == MODE_TRACE) {
					// what two successive DO calls of bwt_smem1_batched return (bwt.c:719-749): the raw list of this
					// bwt_smem1 (ascending start) tagged step*2 + pass, with its return value alongside
					const int pass = lds_u16(SC(CS_PASS));
					const u64 ttag = ((u64)(step * 2 + pass) << STEP_SHIFT) | ((u64)(u32)lds_u16(SC(CS_RET)) << AUX_SHIFT);
					const Intv *const Mp = M1 + (size_t)pass * p.scratch_cap;
					for (int e = n_mem - 1; e >= 0; --e) {
						if (n_out < p.slot_cap) { const Intv t = ld_intv(&Mp[e]); st_intv(&slot[n_out], t.x0, t.x1, t.x2, t.info | ttag); }
						++n_out;
					}
					sts_i32(SC(CS_NOUT), n_out);
					if (pass == 0) {
						const int max_len = lds_u16(SC(CS_MAXLEN)), split_len = lds_u16(SC(CS_SPLIT));
						const u64 max_s = (u64)(u32)lds_i32(SC(CS_MAXS_LO)) | ((u64)(u32)lds_i32(SC(CS_MAXS_HI)) << 32);
						if (n_mem > 0 && split_len > 0 && max_len >= split_len && max_s <= (u64)p.split_width) {   // bwamem.c:272
							sts_u16(SC(CS_PASS), 1);
							sts_u16(SC(CS_X), (lds_u16(SC(CS_MAXEND)) + lds_u16(SC(CS_MAXSTART))) >> 1);
							min_intv = (u32)max_s + 1u;
							phase = PH_INIT_CALL;
							break;
						}
					}
					sts_u16(SC(CS_STEP), step + 1);
					phase = PH_NEXT_STEP;
					break;
				}
				const u64 tag = (u64)step << STEP_SHIFT;
				if (lds_u16(SC(CS_PASS)) == 0) {
					const int max_len = lds_u16(SC(CS_MAXLEN)), split_len = lds_u16(SC(CS_SPLIT));
					const u64 max_s = (u64)(u32)lds_i32(SC(CS_MAXS_LO)) | ((u64)(u32)lds_i32(SC(CS_MAXS_HI)) << 32);
					// bwamem.c:272: re-seed from the middle of the longest SMEM if it is long and (nearly) unique -- unless the
					// repeat filter proves that the second pass cannot contribute (smem_repeat.cuh)
					bool reseed = n_mem > 0 && split_len > 0 && max_len >= split_len && max_s <= (u64)p.split_width;
					bool rf_clear = false;
					if (reseed && p.qflags && (max_len >> 1) >= p.rf_k) {
						const int xm = (lds_u16(SC(CS_MAXEND)) + lds_u16(SC(CS_MAXSTART))) >> 1;
						if (LPR == 1 && p.q_stride <= 64) {                  // the flags came with the read (rf_range_is_clear on registers)
							const int lo = max(xm - p.rf_k + 1, 0), hi = min(xm, len - p.rf_k);
							if (hi >= lo) {
								const int wl = lo >> 5, wh = hi >> 5;
								const u32 w0 = wl == 0 ? qf0 : wl == 1 ? qf1 : wl == 2 ? qf2 : qf3, w1 = wh == 0 ? qf0 : wh == 1 ? qf1 : wh == 2 ? qf2 : qf3;
								const u32 m0 = 0xffffffffu << (lo & 31), m1 = 0xffffffffu >> (31 - (hi & 31));
								rf_clear = (wl == wh ? (w0 & m0 & m1) : ((w0 & m0) | (w1 & m1))) == 0;
							}
						} else
							rf_clear = rf_range_is_clear(p.qflags + (size_t)lds_i32(SC(CS_RID)) * (size_t)(p.q_stride >> 4), p.rf_k, len, xm);
					}
					if (rf_clear) {
						reseed = false;
						if (p.count_skips && !half) atomicAdd(&p.status[6], 1);
					}
					if (reseed) {
						sts_u16(SC(CS_NM1), n_mem); sts_u16(SC(CS_KEEP), max_len); sts_u16(SC(CS_PASS), 1);
						sts_u16(SC(CS_X), (lds_u16(SC(CS_MAXEND)) + lds_u16(SC(CS_MAXSTART))) >> 1);
						min_intv = (u32)max_s + 1u;
						phase = PH_INIT_CALL;
						break;
					}
					if (LPR == 1) {
						for (int e = n_mem - 1; e >= 0; e -= 4) {        // four loads in flight, then the stores
							Intv t0 = ld_intv(&M1[e]), t1 = t0, t2 = t0, t3 = t0;
							if (e >= 1) t1 = ld_intv(&M1[e - 1]);
							if (e >= 2) t2 = ld_intv(&M1[e - 2]);
							if (e >= 3) t3 = ld_intv(&M1[e - 3]);
							if (n_out < p.slot_cap) st_intv(&slot[n_out], t0.x0, t0.x1, t0.x2, t0.info | tag);
							if (e >= 1 && n_out + 1 < p.slot_cap) st_intv(&slot[n_out + 1], t1.x0, t1.x1, t1.x2, t1.info | tag);
							if (e >= 2 && n_out + 2 < p.slot_cap) st_intv(&slot[n_out + 2], t2.x0, t2.x1, t2.x2, t2.info | tag);
							if (e >= 3 && n_out + 3 < p.slot_cap) st_intv(&slot[n_out + 3], t3.x0, t3.x1, t3.x2, t3.info | tag);
							n_out += min(e + 1, 4);
						}
					} else if (PCOPY) {                                  // the two lanes of the pair copy alternate entries: two round trips in flight
						for (int e = n_mem - 1 - half; e >= 0; e -= 2) {
							const int o = n_out + (n_mem - 1 - e);
							if (o < p.slot_cap) { const Intv t = ld_intv(&M1[e]); st_intv(&slot[o], t.x0, t.x1, t.x2, t.info | tag); }
						}
						n_out += n_mem;
					} else
					for (int e = n_mem - 1; e >= 0; --e) {
						if (n_out < p.slot_cap) { const Intv t = ld_intv(&M1[e]); st_intv(&slot[n_out], t.x0, t.x1, t.x2, t.info | tag); }
						++n_out;
					}
				} else {
					// ordered merge, bwamem.c:281-301; both lists are walked in ascending start = reverse emission
					int ia = lds_u16(SC(CS_NM1)) - 1, ib = n_mem - 1;
					const int half_len = lds_u16(SC(CS_KEEP)) >> 1, ori_start = lds_u16(SC(CS_ORI));
					Intv va, vb;
					va.x0 = va.x1 = va.x2 = va.info = 0; vb = va;
					bool have_a = false, have_b = false;
					while (ia >= 0 || ib >= 0) {
						if (ia >= 0 && !have_a) { va = ld_intv(&M1[ia]); have_a = true; }
						if (ib >= 0 && !have_b) { vb = ld_intv(&M2[ib]); have_b = true; }
						bool take_a;
						if (ia >= 0 && ib >= 0) {
							const long long xa = (long long)((va.info >> 32 << 32) | (u32)(len - (int)(u32)va.info));
							const long long xb = (long long)((vb.info >> 32 << 32) | (u32)(len - (int)(u32)vb.info));
							take_a = xa < xb;
						} else take_a = ia >= 0;
						if (take_a) {
							if (n_out < p.slot_cap) st_intv(&slot[n_out], va.x0, va.x1, va.x2, va.info | tag);
							++n_out; --ia; have_a = false;
						} else {
							const int sl = (int)(u32)vb.info - (int)(vb.info >> 32);
							if (sl >= half_len && (int)(u32)vb.info > ori_start) {
								if (n_out < p.slot_cap) st_intv(&slot[n_out], vb.x0, vb.x1, vb.x2, vb.info | tag);
								++n_out;
							}
							--ib; have_b = false;
						}
					}
				}
				sts_i32(SC(CS_NOUT), n_out); sts_u16(SC(CS_STEP), step + 1);
				phase = PH_NEXT_STEP;
			}
			if (phase != PH_NEXT_STEP) break;      // else fall through: the next phase follows
			case PH_NEXT_STEP: {     // head of smem_next2, bwamem.c:249-258
				int start = lds_u16(SC(CS_START));
				// Speculative walk (see PH_FWD_DONE): the previous pass-1 call at CS_ORI ended its forward sweep at `start` because
				// q[CS_ORI .. start] occurs fewer than start_width times; that bounds the coming backward sweep at CS_ORI.
				const int lim1 = (SPEC && lds_u16(SC(CS_STEP)) > 0 && start < len && qbase(start) <= 3) ? lds_u16(SC(CS_ORI)) + 1 : 0;
				while (start < len && qbase(start) > 3) ++start;
				if (start >= len) {
					const int rk = lds_i32(SC(CS_RK)), n_out = lds_i32(SC(CS_NOUT));
					p.counts[rk] = n_out;
					if (n_out > p.slot_cap && !half) { p.overflow_list[atomicAdd(&p.status[1], 1)] = lds_i32(SC(CS_RID)); atomicMax(&p.status[3], n_out); }
					phase = PH_NEED_READ;
					break;
				}
				sts_u16(SC(CS_START), start); sts_u16(SC(CS_ORI), start); sts_u16(SC(CS_X), start); sts_u16(SC(CS_PASS), 0);
				if (SPEC) sts_u16(SC(CS_KEEP), lim1);       // (CS_KEEP is otherwise only used by the merge after a re-seeding pass)
				min_intv = p.start_width < 1 ? 1u : (u32)p.start_width;
				phase = PH_INIT_CALL;
			}
			if (phase != PH_INIT_CALL) break;      // else fall through: the next phase follows
			case PH_INIT_CALL: {     // bwt_set_intv, bwt.c:788-789, then the head of the forward loop
				const int x = lds_u16(SC(CS_X));
				const int c0 = (int)qbase(x);
				a = p.ix.L2[3 - c0] + 1;             // is_back = 0 walks x[1]
				b = p.ix.L2[c0] + 1;
				s = p.ix.L2[c0 + 1] - p.ix.L2[c0];
				end = (u32)(x + 1);
				i = x + 1; n_curr = 0; j = 0;
				sts_u16(SC(CS_NMEM), 0); sts_u16(SC(CS_MAXLEN), 0); sts_u16(SC(CS_MAXSTART), 0); sts_u16(SC(CS_MAXEND), 0);
				sts_i32(SC(CS_MAXS_LO), 0); sts_i32(SC(CS_MAXS_HI), 0);
				guard = (int)min(2ll * (len + 2) * (len + 2) + 64, 0x7fffffffll);   // > every extend one bwt_smem1 can issue (saturates for reads beyond 32 k bases)
				const u32 qv = i < len ? qbase(i) : 4u;
				if (qv > 3) phase = PH_FWD_END;
				else { c = 3 - (int)qv; phase = PH_FWD; }     // bwt.c:793: forward extension uses the complement
			} break;
			default: break;
			}
		}
		__syncwarp();
		if (__all_sync(FULL_MASK, phase == PH_IDLE)) break;

		// ============================================================== one bwt_extend per pair (warp converged)
		// idle pairs ride along on the interval (1,1,1): its block is hot in L2 and the result is dropped
		// Unique forward walk (PH_UW_*): once the forward sweep of a pass-1 call has extended an interval of size 1 three times,
		// the rest of the walk is a comparison of the read with the text at the pattern's only occurrence: text position
		// t = SA[x0] (one gather), the text itself (one gather of 2 x 64 bases, a nibble each like the staged read), and the
		// reverse-strand row of the longer pattern = ISA[n - t - length] (T = forward + reverse complement, so rc(P) sits
		// mirrored); x0 and the size stay.  SA and ISA entries are 33 bits, seven to a 32-byte sector (uw_get33); the ISA is
		// sampled every 2^uw_isa_shift positions counted from the END of the text, so the row comes from the sample at or above
		// the position (one gather) plus (t + length) mod 2^shift ordinary backward extends of that single row by the
		// complemented read bases it is known to be preceded by (PH_UW_LF).  The gathers go through extend_pair's own load (same
		// registers); the walk ends where the reference's last bwt_extend fails.  t rides in last_s, which the forward sweep
		// does not use.
		const uint4 *alt = nullptr, *alt2 = nullptr;
		if (phase >= PH_UW_SA && phase <= PH_UW_ISA) {              // (only entered with the tables present)
			const u64 tq = last_s + (u64)(end - (u32)lds_u16(SC(CS_X)));    // t + pattern length (PH_UW_SA: unused)
			u64 byte;
			if (phase == PH_UW_SA) byte = (u64)p.uw_fsa + 32 * (b / 7);
			else if (phase == PH_UW_ISA) byte = (u64)p.uw_isa + 32 * ((tq >> p.uw_isa_shift) / 7);
			else byte = (u64)p.uw_text + 32 * ((tq >> 6) + (u64)half);       // (SEC: each lane still compares its own 64-base sector)
			alt = reinterpret_cast<const uint4 *>(byte);
			if (LPR == 1 && phase == PH_UW_TEXT) alt2 = alt + 2;     // one lane fetches both 64-base sectors of the window
		}
		Ext ok;
		if (LPR == 1 || SEC) ok = extend_single(p.ix, a, b, s, c & 3, blk_k, blk_l, alt, alt2);
		else {
#ifdef SEED_NARROW
			ok = extend_pair<!WIDE>(p.ix, a, b, s, c & 3, half, lane, p.hot_min_intv, pol_hot, pol_cold, blk_k, blk_l, alt);
#else
			ok = extend_pair<false>(p.ix, a, b, s, c & 3, half, lane, p.hot_min_intv, pol_hot, pol_cold, blk_k, blk_l, alt);
#endif
		}
		if (phase == PH_IDLE) continue;
		if (--guard < 0) { if (!half) atomicAdd(&p.status[2], 1); p.counts[lds_i32(SC(CS_RK))] = 0; phase = PH_NEED_READ; continue; }

		// ============================================================== consume the result, set up the next extend
		// (forward: bwt.c:794-799, backward: bwt.c:813-824; written once for both so that the warp does not
		//  run two divergent copies of the push / advance code)
		if (alt) {
			if (phase == PH_UW_TEXT) {
				// text (pack_text_nib_kernel) and read carry one base per nibble, first base lowest: eight bases per XOR.  This lane
				// compares the part of [tp, lim) that lies in its 64-base sector (LPR 1: in both sectors of the window); the read word is
				// funnel-shifted into place.
				constexpr int SPAN = LPR == 2 ? 64 : 128;
				const u64 tp = last_s + (u64)(end - (u32)lds_u16(SC(CS_X)));   // text position that faces read base i
				const u64 sec0 = tp >> 6, mybase = (sec0 + (u64)half) << 6;
				const u64 lim = min(min(tp + (u64)(len - i), p.ix.seq_len), (sec0 + 2) << 6);   // compare [tp, lim)
				const int r_lo = tp > mybase ? (int)(tp - mybase) : 0, r_hi = lim > mybase ? (int)min(lim - mybase, (u64)SPAN) : 0;
				int first = SPAN;
				if (r_hi > r_lo) {
					const int base_q = i + (int)((long long)mybase - (long long)tp);   // read index facing the sector's first base (< 0: before i)
					const int nw1 = (p.q_stride >> 2) - 1, qw = base_q >> 3;
					const u32 sh = ((u32)base_q & 7u) * 4u;
					u32 lo = (u32)lds_i32(QBYTE(4u * (u32)min(max(qw, 0), nw1)));   // (clamped words only face masked positions)
					// words w_lo .. w_hi take part; only the first and the last of them partially
					const int w_lo = r_lo >> 3, w_hi = (r_hi - 1) >> 3, rh = r_hi - 8 * w_hi;
					const u32 m_lo = 0xffffffffu << (4 * (r_lo & 7)), m_hi = rh >= 8 ? 0xffffffffu : ~(0xffffffffu << (4 * rh));
#pragma unroll
					for (int wq = 0; wq < SPAN / 8; ++wq) {
						const u32 hi = (u32)lds_i32(QBYTE(4u * (u32)min(max(qw + 1 + wq, 0), nw1)));
						u32 d = (wq < 8 ? blk_k[wq & 7] : blk_l[wq & 7]) ^ __funnelshift_r(lo, hi, sh);
						lo = hi;
						u32 vm = (wq >= w_lo && wq <= w_hi) ? 0xffffffffu : 0u;
						if (wq == w_lo) vm &= m_lo;
						if (wq == w_hi) vm &= m_hi;
						d &= vm;
						const u32 m = (d | (d >> 1) | (d >> 2)) & 0x11111111u;        // a base differs (an ambiguous read base, 4, always does)
						if (m && first == SPAN) first = 8 * wq + ((__ffs((int)m) - 1) >> 2);
					}
				}
				const bool full = first == SPAN;
				const u32 cnt = r_hi > r_lo ? (u32)((full ? r_hi : first) - r_lo) : 0u;
				u32 c0 = cnt, c1 = 0;
				bool f0 = full, f1 = true;
				if (LPR == 2) {
					c0 = __shfl_sync(pm2, cnt, lane & ~1); c1 = __shfl_sync(pm2, cnt, lane | 1);
					f0 = __shfl_sync(pm2, (int)full, lane & ~1) != 0; f1 = __shfl_sync(pm2, (int)full, lane | 1) != 0;
				}
				const u32 m = f0 ? c0 + c1 : c0;
				i += (int)m; end = (u32)i;
				// ran through the whole 128-base window without a verdict: another round of text, else the inverse SA
				if (!(f0 && f1 && tp + m == ((sec0 + 2) << 6) && i < len && tp + m < p.ix.seq_len)) phase = PH_UW_ISA;
			} else {
				const u64 tq = last_s + (u64)(end - (u32)lds_u16(SC(CS_X)));
				const u32 slot = (u32)((phase == PH_UW_SA ? b : tq >> p.uw_isa_shift) % 7);
				u32 lo = blk_k[0];
#pragma unroll
				for (int t = 1; t < 7; ++t) if (slot == (u32)t) lo = blk_k[t];
				const u64 val = (u64)lo | ((u64)((blk_k[7] >> slot) & 1u) << 32);
				if (phase == PH_UW_SA) { last_s = val; phase = PH_UW_TEXT; }
				else {
					// The comparison stopped at the end of the read, at an ambiguous base (bwt.c:800-806), at the end of the text or at
					// a base that differs from the text at the only occurrence: the reference's next bwt_extend finds nothing
					// (bwt.c:796-798).  Either way the interval is pushed and the forward sweep is over -- once x[1], the row of
					// rc(pattern), is known: `val` is the row of the sampled suffix j positions above it; each backward extend of
					// that single row by the base in front of it (the complement of read base i - j, then i - j + 1, ...) moves
					// one position down.  x[0] (b) and the size stay.
					a = val;
					j = (int)(tq & ((1ull << p.uw_isa_shift) - 1));
					if (j == 0) phase = PH_FWD_END;
					else { phase = PH_UW_LF; c = 3 - (int)qbase(i - j); }
				}
			}
			continue;
		}
		if (phase == PH_UW_LF) {                                 // one position down the sampled inverse suffix array
			a = ok.a;
			if (--j == 0) phase = PH_FWD_END;
			else c = 3 - (int)qbase(i - j);
			continue;
		}
		if (SPEC && phase == PH_SPEC) {
			if (ok.s < min_intv) {                               // died before the limit: the shorter candidates matter after all
				i = lds_u16(SC(CS_X)) - 1; j = 0; n_prev = n0; n_curr = 0;
				c = (int)qbase(i);                           // (valid: the walk started with it)
				phase = PH_BWD;
				load_prev();
				continue;
			}
			a = ok.a; b = ok.b; s = ok.s;
			--i;
			c = i < 0 ? -1 : (int)qbase(i);
			if (c > 3) c = -1;
			if (c < 0 || i < lds_u16(SC(CS_KEEP))) phase = PH_BWD_LAST;   // everything dies at the next round: emit (i + 1, end)
			continue;
		}
		if (SPLIT && phase == PH_BWD) {
			// two entries of the round at once: A = prev[j] in lane 0, B = prev[j + 1] in lane 1 (if there is one)
			const bool two = j + 1 < n_prev;
			const u32 sA = __shfl_sync(pm2, (u32)ok.s, lane & ~1), sB = __shfl_sync(pm2, (u32)ok.s, lane | 1);
			const bool smallA = sA < min_intv, smallB = sB < min_intv;
			const bool emitA = smallA && n_curr == 0;
			const bool pushA = !smallA && (n_curr == 0 || (u64)sA != last_s);
			const int idxA = n0 - 1 - n_curr;
			if (pushA) { ++n_curr; last_s = sA; }
			const bool emitB = two && !emitA && smallB && n_curr == 0;        // (a second emission of the same round never passes bwt.c:815's test)
			const bool pushB = two && !smallB && (n_curr == 0 || (u64)sB != last_s);
			const int idxB = n0 - 1 - n_curr;
			if (pushB) { ++n_curr; last_s = sB; }
			// each lane stores its own entry; the emission updates the pair's cold state, which the other lane reads after the sync
			if (half ? emitB : emitA) emit(a, b, s, end, i + 1);
			if (half ? pushB : pushA) b_put(half ? idxB : idxA, ok.a, ok.b, ok.s, end);
			__syncwarp(pm2);
			j += two ? 2 : 1;
			if (j >= n_prev) {                               // bwt.c:826-827
				if (n_curr == 0) { phase = PH_CALL_DONE; continue; }
				n_prev = n_curr; n_curr = 0; j = 0; --i;
				c = i < 0 ? -1 : (int)qbase(i);
				if (c > 3) c = -1;
				if (c < 0) phase = PH_BWD_LAST;
			}
			load_prev();
			continue;
		}
		const bool fwd = SPLIT ? true : phase == PH_FWD;     // (SPLIT: the backward sweep never gets here; 18.3 -> 17.9 ms for the dead code alone)
		const bool small = ok.s < min_intv;
		const bool diff = ok.s != (fwd ? s : last_s);
		const bool push = fwd ? diff : (!small && (n_curr == 0 || diff));
		if (!fwd && small && n_curr == 0) emit(a, b, s, end, i + 1);
		if (push) {
			// forward pushes the interval it leaves (ik), backward the one it arrives at (ok[c], info inherited)
			const int idx = fwd ? n_curr : n0 - 1 - n_curr;      // backward: n_curr <= j, lands at or behind the slot just read
			b_put(idx, fwd ? b : ok.a, fwd ? a : ok.b, fwd ? s : ok.s, end);
			++n_curr;
			last_s = ok.s;
		}
		if (fwd) {
			if (diff && small) { phase = PH_FWD_DONE; continue; }
			a = ok.a; b = ok.b; s = ok.s; end = (u32)(i + 1);
			++i;
			const u32 qv = i < len ? qbase(i) : 4u;
			if (qv > 3) phase = PH_FWD_END;
			else c = 3 - (int)qv;
			if (MODE != MODE_SMEM1) {                            // (uw_min_run is out of reach without the tables)
				j = (s == 1 && min_intv == 1) ? j + 1 : 0;           // consecutive extends of a unique interval
				if (j >= p.uw_min_run && phase == PH_FWD && len - i >= p.uw_min_left) {                  // (a is rebuilt from the inverse SA)
					phase = PH_UW_SA; a = 1;
					if (p.count_skips && !half) atomicAdd(&p.status[7], 1);
				}
			}
		} else {
			if (++j == n_prev) {                             // bwt.c:826-827
				if (n_curr == 0) { phase = PH_CALL_DONE; continue; }
				n_prev = n_curr; n_curr = 0; j = 0; --i;
				c = i < 0 ? -1 : (int)qbase(i);
				if (c > 3) c = -1;
				if (c < 0) phase = PH_BWD_LAST;
			}
			load_prev();
		}
	}
}

// ---------------------------------------------------------------------------------------------
// Pack pre-pass: whatever form the caller's reads arrive in, the seed kernel stages a read from `qpack` -- one base per
// nibble (0..3, > 3 ambiguous, padded with 4), eight bases per 32-bit word, first base lowest, q_stride bytes per read --
// with a few 16-byte copies, and takes its length from rlen[].  One thread per 16-byte chunk (32 bases).  Reads whose
// length does not fit the handle set status[5] (checked on the host after the run: no per-read loop on the CPU).

// 32-bit word at an arbitrary byte address: two aligned loads and a funnel shift (reads stay inside the staging buffer:
// it is allocated with 64 bytes of slack and starts 256-byte aligned)
__device__ __forceinline__ u32 ld_u32_unaligned(const uint8_t *p)
{
	const u32 *q = reinterpret_cast<const u32 *>(reinterpret_cast<uintptr_t>(p) & ~(uintptr_t)3);
	const u32 sh = ((u32)reinterpret_cast<uintptr_t>(p) & 3u) * 8u;
	return __funnelshift_r(__ldg(q), __ldg(q + 1), sh);
}
// bases [8k, 8k+8) of a chunk that holds n_valid real bases: everything behind the read's end becomes 4
__device__ __forceinline__ u32 pad_nibbles(u32 w, int n_valid, int k)
{
	const int v = min(max(n_valid - 8 * k, 0), 8);
	const u32 m = v >= 8 ? 0xffffffffu : ((1u << (4 * v)) - 1u);
	return (w & m) | (0x44444444u & ~m);
}

// (a) one base per byte (0..3, > 3 ambiguous; bwamem.c:1403-1406) at CSR offsets -- the layout of smem_gpu_collect
__global__ void __launch_bounds__(256) pack_bytes_kernel(const uint8_t *__restrict__ seq, const long long *__restrict__ offs, long long n,
                                                         int chunks_per_read, int max_len, uint4 *__restrict__ qpack, int *__restrict__ rlen,
                                                         int *__restrict__ status)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t / chunks_per_read;
	if (r >= n) return;
	const int ch = (int)(t % chunks_per_read);
	const long long o0 = offs[r], l64 = offs[r + 1] - o0;
	const bool bad = l64 < 0 || l64 > (long long)max_len;
	const int len = bad ? 0 : (int)l64, base = 32 * ch;
	if (ch == 0) { rlen[r] = len; if (bad) atomicOr(&status[5], 1); }
	const int n_valid = min(max(len - base, 0), 32);
	u32 w[4];
	const uint8_t *src = seq + o0 + base;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		u32 v = 0x44444444u;
		if (8 * k < n_valid) {
			u32 lo = __vminu4(ld_u32_unaligned(src + 8 * k), 0x04040404u), hi = __vminu4(ld_u32_unaligned(src + 8 * k + 4), 0x04040404u);
			lo = (lo | (lo >> 4)) & 0x00ff00ffu; lo = (lo | (lo >> 8)) & 0x0000ffffu;      // four bytes -> four nibbles, first base lowest
			hi = (hi | (hi >> 4)) & 0x00ff00ffu; hi = (hi | (hi >> 8)) & 0x0000ffffu;
			v = pad_nibbles(lo | (hi << 16), n_valid, k);
		}
		w[k] = v;
	}
	qpack[t] = make_uint4(w[0], w[1], w[2], w[3]);
}

// (b) the compact wire format (smem_reads2_t): fixed-stride records of two bits per base, four bases per byte, first base
// in the top bits (the reference's .pac convention, bntseq.c:_set_pac); lengths from lens[] or one length for all
__global__ void __launch_bounds__(256) unpack2_kernel(const uint8_t *__restrict__ seq2, int stride, const unsigned short *__restrict__ lens, int read_len,
                                                      long long n, int chunks_per_read, int max_len, uint4 *__restrict__ qpack, int *__restrict__ rlen,
                                                      int *__restrict__ status)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t / chunks_per_read;
	if (r >= n) return;
	const int ch = (int)(t % chunks_per_read);
	const int l0 = lens ? (int)lens[r] : read_len;
	const bool bad = l0 < 0 || l0 > max_len || l0 > 4 * stride;
	const int len = bad ? 0 : l0, base = 32 * ch;
	if (ch == 0) { rlen[r] = len; if (bad) atomicOr(&status[5], 1); }
	const int n_valid = min(max(len - base, 0), 32);
	u32 w[4];
	const uint8_t *src = seq2 + (size_t)r * (size_t)stride + 8 * ch;
#pragma unroll
	for (int k = 0; k < 4; ++k) {
		u32 v = 0x44444444u;
		if (8 * k < n_valid) {
			const u32 two = ld_u32_unaligned(src + 2 * k) & 0xffffu;            // bytes 2k, 2k+1 of the chunk = bases 8k .. 8k+7
			const u32 b0 = two & 255u, b1 = two >> 8;
			const u32 n0 = (b0 >> 6) | (b0 & 0x30u) | ((b0 & 0x0cu) << 6) | ((b0 & 3u) << 12);
			const u32 n1 = (b1 >> 6) | (b1 & 0x30u) | ((b1 & 0x0cu) << 6) | ((b1 & 3u) << 12);
			v = pad_nibbles(n0 | (n1 << 16), n_valid, k);
		}
		w[k] = v;
	}
	qpack[t] = make_uint4(w[0], w[1], w[2], w[3]);
}

// ... whose ambiguous bases travel as an exception list {read, pos} (a two-bit code cannot say N): bit 2 of the nibble is set
struct AmbEntry { u32 read; unsigned short pos, reserved; };
__global__ void amb_patch_kernel(const AmbEntry *__restrict__ amb, long long n_amb, long long read_lo, long long n, int q_stride,
                                 u32 *__restrict__ qpack_words, const int *__restrict__ rlen, int *__restrict__ status)
{
	const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (e >= n_amb) return;
	const long long r = (long long)amb[e].read - read_lo;
	const int pos = amb[e].pos;
	if (r < 0 || r >= n) { atomicOr(&status[5], 2); return; }
	if (pos >= rlen[r]) return;
	atomicOr(&qpack_words[(size_t)r * (size_t)(q_stride >> 2) + (pos >> 3)], 4u << (4 * (pos & 7)));
}

// Repeat-filter flags of the 32 windows that start in a chunk (smem_repeat.cuh), from the packed reads: rolling K-mer code,
// `amb` = bases until the most recent ambiguous one leaves the window; eight independent table loads per round.
__global__ void __launch_bounds__(256) window_flags_kernel(const uint4 *__restrict__ qpack, const int *__restrict__ rlen, long long n, int chunks_per_read,
                                                           const u32 *__restrict__ rf_bits, int K, int log2_bits, u32 *__restrict__ qflags)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t / chunks_per_read;
	if (r >= n) return;
	const int ch = (int)(t % chunks_per_read);
	const int len = rlen[r], base = 32 * ch;
	if (base + K > len) { qflags[t] = 0xffffffffu; return; }            // every window of this chunk runs off the read
	u32 q[8];
	{
		const uint4 c0 = __ldg(qpack + t);
		uint4 c1 = make_uint4(0x44444444u, 0x44444444u, 0x44444444u, 0x44444444u);
		if (ch + 1 < chunks_per_read) c1 = __ldg(qpack + t + 1);
		q[0] = c0.x; q[1] = c0.y; q[2] = c0.z; q[3] = c0.w; q[4] = c1.x; q[5] = c1.y; q[6] = c1.z; q[7] = c1.w;
	}
	// bases base .. base + 63 in order (static register indices): base i is the last base of the window starting at
	// a = i - (K - 1); windows 0 .. 31 belong to this chunk
	u32 flags = 0;
	u64 code = 0;
	int amb = 0;
	const u64 kmask = K == 32 ? ~0ull : (1ull << (2 * K)) - 1;
#pragma unroll
	for (int w = 0; w < 8; ++w) {
		u32 wd[8], sh[8], bad[8];
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const u32 c = (q[w] >> (4 * k)) & 15u;
			code = ((code << 2) | (c & 3u)) & kmask;
			amb = c > 3u ? K : max(amb - 1, 0);
			const int a0 = 8 * w + k - (K - 1);
			bad[k] = amb > 0;
			const u64 b = rf_bit_index(code, log2_bits);
			wd[k] = (bad[k] || a0 < 0 || a0 > 31) ? 0u : __ldg(rf_bits + (b >> 5));
			sh[k] = (u32)(b & 31);
		}
#pragma unroll
		for (int k = 0; k < 8; ++k) {
			const int a0 = 8 * w + k - (K - 1);
			if (a0 >= 0 && a0 <= 31) flags |= (bad[k] | ((wd[k] >> sh[k]) & 1u)) << a0;
		}
	}
	qflags[t] = flags;
}

// Upload-time re-pack of bwt_t::bwt (bwtindex.c:128-150: per 128 symbols 4 x uint64 checkpoints + 8 words of
// 16 two-bit symbols, first symbol in the top bits; last block truncated) into the split bit-plane block
// described in smem_device.cuh.  One thread per block.
__global__ void repack_kernel(const u32 *__restrict__ src, u64 n_blocks, u64 seq_len, uint4 *__restrict__ dst)
{
	const u64 bidx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (bidx >= n_blocks) return;
	const u32 *s = src + bidx * 16;
	const u64 rem = seq_len - bidx * 128;                         // symbols in this block
	const int nw = rem >= 128 ? 8 : (int)((rem + 15) >> 4);        // symbol words actually present
	u32 cnt[8], hi[4] = {0, 0, 0, 0}, lo[4] = {0, 0, 0, 0};
#pragma unroll
	for (int k = 0; k < 8; ++k) cnt[k] = s[k];
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const u32 w = k < nw ? s[8 + k] : 0u;
		u32 h = 0, l = 0;                                          // compress the odd / even bits of w into 16 bits
#pragma unroll
		for (int t = 0; t < 16; ++t) { h |= ((w >> (2 * t + 1)) & 1u) << t; l |= ((w >> (2 * t)) & 1u) << t; }
		hi[k >> 1] |= h << (16 * (1 - (k & 1)));
		lo[k >> 1] |= l << (16 * (1 - (k & 1)));
	}
	uint4 *d = dst + bidx * 4;
	d[0] = make_uint4(cnt[0], cnt[1], cnt[2], cnt[3]);             // sector 0: checkpoints A, C
	d[1] = make_uint4(hi[0], hi[1], lo[0], lo[1]);                 //           planes of symbols 0..63
	d[2] = make_uint4(cnt[4], cnt[5], cnt[6], cnt[7]);             // sector 1: checkpoints G, T
	d[3] = make_uint4(hi[2], hi[3], lo[2], lo[3]);                 //           planes of symbols 64..127
}

// 64-byte blocks -> 32-byte sectors with their own 32-bit checkpoints (smem_device.cuh, extend_single).  One thread per block;
// the second sector's checkpoints are the block's plus the symbol counts of its first half.
__global__ void sectors_from_blocks_kernel(const uint4 *__restrict__ blk, u64 n_blocks, uint4 *__restrict__ sec)
{
	const u64 bidx = (u64)blockIdx.x * blockDim.x + threadIdx.x;
	if (bidx >= n_blocks) return;
	const uint4 c01 = blk[bidx * 4], p0 = blk[bidx * 4 + 1], c23 = blk[bidx * 4 + 2], p1 = blk[bidx * 4 + 3];
	const u32 nh = __popc(p0.x) + __popc(p0.y), nl = __popc(p0.z) + __popc(p0.w), nt = __popc(p0.x & p0.z) + __popc(p0.y & p0.w);
	sec[bidx * 4] = make_uint4(c01.x, c01.z, c23.x, c23.z);                  // low words of the A, C, G, T checkpoints
	sec[bidx * 4 + 1] = p0;
	sec[bidx * 4 + 2] = make_uint4(c01.x + (64u - nh - nl + nt), c01.z + (nl - nt), c23.x + (nh - nt), c23.z + nt);
	sec[bidx * 4 + 3] = p1;
}

// Latency path of small batches (n <= TINY_MAX reads): ONE CTA does what scan_*_kernel + compact_kernel + publish_status_kernel do for
// large ones -- counts -> offsets, slots -> dense output -- and also stores the results into a pinned host arena, so that the host
// needs a single wait and no further copy: arena = [status 16 x int | off (n + 1) x int64 | ret n x int | intv | step | aux] with the
// last three sized arena_cap entries.  Reads that outgrew their slots, or a total beyond the capacities, are only reported
// (status[1] / the total): the host then takes the ordinary path.  Leaves the status words zeroed for the next run.
#define TINY_MAX 2048
#define TINY_TPB 256
__global__ void __launch_bounds__(TINY_TPB) tiny_tail_kernel(const Intv *__restrict__ slots, int slot_cap, const int *__restrict__ counts, int n, int *__restrict__ status,
                                                             const int *__restrict__ ret, long long *__restrict__ off, Intv *__restrict__ out,
                                                             unsigned short *__restrict__ step_out, unsigned short *__restrict__ aux_out, long long out_cap,
                                                             int *__restrict__ h_status, long long *__restrict__ h_off, int *__restrict__ h_ret,
                                                             Intv *__restrict__ h_out, unsigned short *__restrict__ h_step, unsigned short *__restrict__ h_aux,
                                                             long long arena_cap)
{
	__shared__ long long wsum[TINY_TPB / 32];
	__shared__ long long s_off[TINY_MAX + 1];
	const int t = threadIdx.x, per = (n + TINY_TPB - 1) / TINY_TPB;
	const int r0 = t * per, r1 = min(r0 + per, n);
	long long sum = 0;
	for (int r = r0; r < r1; ++r) sum += counts[r];
	long long incl = sum;
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const long long o = __shfl_up_sync(FULL_MASK, incl, d); if ((t & 31) >= d) incl += o; }
	if ((t & 31) == 31) wsum[t >> 5] = incl;
	__syncthreads();
	long long base = incl - sum, total = 0;
#pragma unroll
	for (int w = 0; w < TINY_TPB / 32; ++w) { if (w < (t >> 5)) base += wsum[w]; total += wsum[w]; }
	for (int r = r0; r < r1; ++r) { s_off[r] = base; off[r] = base; h_off[r] = base; if (h_ret) h_ret[r] = ret[r]; base += counts[r]; }
	if (t == 0) { s_off[n] = total; off[n] = total; h_off[n] = total; }
	__syncthreads();
	const bool ok = status[1] == 0 && total <= out_cap;          // (else: the ordinary path's overflow re-run / buffer growth)
	const bool to_host = ok && total <= arena_cap;
	if (ok) {
		for (int r = t >> 5; r < n; r += TINY_TPB / 32) {
			const long long o0 = s_off[r];
			const int c = (int)(s_off[r + 1] - o0);
			for (int e = t & 31; e < c; e += 32) {
				const Intv v = ld_intv(&slots[(size_t)r * slot_cap + e]);
				const unsigned short st = (unsigned short)(v.info >> STEP_SHIFT), ax = (unsigned short)(v.info >> AUX_SHIFT);
				st_intv(&out[o0 + e], v.x0, v.x1, v.x2, v.info & INFO_MASK);
				step_out[o0 + e] = st;
				if (aux_out) aux_out[o0 + e] = ax;
				if (to_host) {
					st_intv(&h_out[o0 + e], v.x0, v.x1, v.x2, v.info & INFO_MASK);
					h_step[o0 + e] = st;
					if (h_aux) h_aux[o0 + e] = ax;
				}
			}
		}
	}
	__syncthreads();
	if (t < 8) { h_status[t] = status[t]; }
	if (t == 8) { h_status[8] = (int)(u32)total; h_status[9] = (int)(u32)((unsigned long long)total >> 32); h_status[10] = to_host ? 1 : 0; h_status[11] = ok ? 1 : 0; }
	__threadfence_system();
	__syncthreads();
	if (t < 8) status[t] = 0;
}

// ---------------------------------------------------------------------------------------------
// counts -> CSR offsets: a small three-kernel exclusive scan (int counts -> int64 offsets).  Hand-written
// rather than a library scan so that its CTAs (128 threads, few registers, 32 B of shared memory) fit into
// the slots a running persistent seed kernel leaves free: with several pipeline lanes on one GPU the scan and
// compaction of a finished lane run next to the following lane's seed kernel.
#define SCAN_TPB 128
#define SCAN_PER_THREAD 8
#define SCAN_PER_BLOCK (SCAN_TPB * SCAN_PER_THREAD)

__global__ void __launch_bounds__(SCAN_TPB) scan_local_kernel(const int *__restrict__ counts, long long n, long long *__restrict__ off,
                                                              long long *__restrict__ bsum)
{
	__shared__ long long wsum[SCAN_TPB / 32];
	const long long base = (long long)blockIdx.x * SCAN_PER_BLOCK + (long long)threadIdx.x * SCAN_PER_THREAD;
	int v[SCAN_PER_THREAD];
	long long t = 0;
#pragma unroll
	for (int k = 0; k < SCAN_PER_THREAD; ++k) { v[k] = base + k < n ? counts[base + k] : 0; t += v[k]; }
	long long incl = t;                                  // inclusive scan of the thread totals inside the warp
#pragma unroll
	for (int d = 1; d < 32; d <<= 1) { const long long o = __shfl_up_sync(FULL_MASK, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
	if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
	__syncthreads();
	long long woff = 0, total = 0;
#pragma unroll
	for (int w = 0; w < SCAN_TPB / 32; ++w) { if (w < (int)(threadIdx.x >> 5)) woff += wsum[w]; total += wsum[w]; }
	long long run = woff + incl - t;
#pragma unroll
	for (int k = 0; k < SCAN_PER_THREAD; ++k) { if (base + k < n) off[base + k] = run; run += v[k]; }
	if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

// exclusive scan of the block sums in place (one CTA; nb <= a few thousand)
__global__ void __launch_bounds__(SCAN_TPB) scan_bsum_kernel(long long *__restrict__ bsum, int nb)
{
	__shared__ long long wsum[SCAN_TPB / 32];
	__shared__ long long carry_s;
	if (threadIdx.x == 0) carry_s = 0;
	__syncthreads();
	for (int b0 = 0; b0 < nb; b0 += SCAN_TPB) {
		const int i = b0 + threadIdx.x;
		const long long t = i < nb ? bsum[i] : 0;
		long long incl = t;
#pragma unroll
		for (int d = 1; d < 32; d <<= 1) { const long long o = __shfl_up_sync(FULL_MASK, incl, d); if ((threadIdx.x & 31) >= d) incl += o; }
		if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
		__syncthreads();
		long long woff = 0, total = 0;
#pragma unroll
		for (int w = 0; w < SCAN_TPB / 32; ++w) { if (w < (int)(threadIdx.x >> 5)) woff += wsum[w]; total += wsum[w]; }
		const long long carry = carry_s;
		if (i < nb) bsum[i] = carry + woff + incl - t;
		__syncthreads();
		if (threadIdx.x == 0) carry_s = carry + total;
		__syncthreads();
	}
}

__global__ void __launch_bounds__(SCAN_TPB) scan_add_kernel(long long *__restrict__ off, long long n, const long long *__restrict__ bsum)
{
	const long long i = (long long)blockIdx.x * SCAN_TPB + threadIdx.x;
	if (i < n) off[i] += bsum[i / SCAN_PER_BLOCK];
}

// Compaction: slots[n][slot_cap] -> dense intv[total] (+ optional step array), info restored.
// Eight lanes per read, each copying entries lane, lane+8, ... (most reads have <= 16 intervals).
__global__ void __launch_bounds__(128) compact_kernel(const Intv *__restrict__ slots, int slot_cap, const int *__restrict__ counts,
                                                      const long long *__restrict__ off, long long n, Intv *__restrict__ out,
                                                      unsigned short *__restrict__ step_out, unsigned short *__restrict__ aux_out, long long out_cap)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t >> 3;
	if (r >= n) return;
	const long long o0 = off[r];
	// entries beyond slot_cap are filled in by the overflow re-run; entries beyond the buffer wait for the host to re-size it
	const int c = (int)min((long long)min(counts[r], slot_cap), max(out_cap - o0, 0ll));
	for (int e = (int)(t & 7); e < c; e += 8) {
		const Intv v = ld_intv(&slots[(size_t)r * slot_cap + e]);
		st_intv(&out[o0 + e], v.x0, v.x1, v.x2, v.info & INFO_MASK);
		if (step_out) step_out[o0 + e] = (unsigned short)(v.info >> STEP_SHIFT);
		if (aux_out) aux_out[o0 + e] = (unsigned short)(v.info >> AUX_SHIFT);
	}
}

// The compact result record (smem_intv16_t, include/smem_gpu.h): x0 | x1 << 33 | x2 << 66 | qbeg << 99 | qend << 113 --
// 127 bits, valid while seq_len < 2^33 and reads are shorter than 2^14 (the host checks both before choosing this form).
__device__ __forceinline__ uint4 pack_intv16(const Intv &v)
{
	const u64 qb = (v.info >> 32) & 0x3fffull, qe = v.info & 0x3fffull;
	const u64 lo = v.x0 | (v.x1 << 33);
	const u64 hi = (v.x1 >> 31) | (v.x2 << 2) | (qb << 35) | (qe << 49);
	return make_uint4((u32)lo, (u32)(lo >> 32), (u32)hi, (u32)(hi >> 32));
}

// Compaction into the compact wire format: slots -> dense 16-byte records + 32-bit CSR offsets (off32[n] = total).
__global__ void __launch_bounds__(128) compact_packed_kernel(const Intv *__restrict__ slots, int slot_cap, const int *__restrict__ counts,
                                                             const long long *__restrict__ off, long long n, uint4 *__restrict__ out,
                                                             u32 *__restrict__ off32, long long out_cap)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t >> 3;
	if (r >= n) return;
	const long long o0 = off[r];
	const int c = (int)min((long long)min(counts[r], slot_cap), max(out_cap - o0, 0ll));      // (see compact_kernel)
	if ((t & 7) == 0) { off32[r] = (u32)o0; if (r == n - 1) off32[n] = (u32)off[n]; }
	for (int e = (int)(t & 7); e < c; e += 8) {
		Intv v = ld_intv(&slots[(size_t)r * slot_cap + e]);
		v.info &= INFO_MASK;
		out[o0 + e] = pack_intv16(v);
	}
}

// The 12-byte result record (smem_intv12_t, include/smem_gpu.h): w[0] = low word of x0, w[1] = low word of x1, w[2] = bit 32 of x0 |
// bit 32 of x1 << 1 | qbeg << 2 | (qend - 1) << (2 + P) | field << (2 + 2P), P = pos_bits; field = x2 - 1 while that is below
// the all-ones value of its 30 - 2P bits, else all ones and the interval's x2 travels in the exception list {index, x2}
// (unordered; status[4] counts them, entries beyond exc_cap are dropped and the host re-runs the compaction with room).
struct Exc12 { u32 index, x2_lo, x2_hi; };
// rec_bytes = 12: three 32-bit words.  rec_bytes = 11 (smem_intv11_t): the third word shrinks to 24 bits -- the size field keeps
// 22 - 2P bits (8 for reads up to 128 bases) -- and the records are byte-packed.  rec_bytes = -11: the 11-byte record's content in the
// 12-byte layout (aligned word stores); squeeze11_kernel then drops every twelfth byte with coalesced word traffic.
__device__ __forceinline__ void put_intv12(u32 *__restrict__ out, long long o, const Intv &v, int pos_bits, Exc12 *__restrict__ exc, long long exc_cap,
                                           int *__restrict__ n_exc, int rec_bytes = 12)
{
	const u32 qb = (u32)(v.info >> 32) & 0xffffu, qe = (u32)v.info & 0xffffu;
	const int fbits = (rec_bytes == 12 ? 30 : 22) - 2 * pos_bits;
	const u32 esc = (1u << fbits) - 1u;
	const u64 x2m = v.x2 - 1;
	const bool is_exc = x2m >= (u64)esc;
	const u32 w2 = (u32)(v.x0 >> 32 & 1) | ((u32)(v.x1 >> 32 & 1) << 1) | (qb << 2) | ((qe - 1u) << (2 + pos_bits)) | ((is_exc ? esc : (u32)x2m) << (2 + 2 * pos_bits));
	if (rec_bytes != 11) { out[3 * o] = (u32)v.x0; out[3 * o + 1] = (u32)v.x1; out[3 * o + 2] = w2; }
	else {
		uint8_t *b = reinterpret_cast<uint8_t *>(out) + 11 * o;
		const u32 w0 = (u32)v.x0, w1 = (u32)v.x1;
		b[0] = (uint8_t)w0; b[1] = (uint8_t)(w0 >> 8); b[2] = (uint8_t)(w0 >> 16); b[3] = (uint8_t)(w0 >> 24);
		b[4] = (uint8_t)w1; b[5] = (uint8_t)(w1 >> 8); b[6] = (uint8_t)(w1 >> 16); b[7] = (uint8_t)(w1 >> 24);
		b[8] = (uint8_t)w2; b[9] = (uint8_t)(w2 >> 8); b[10] = (uint8_t)(w2 >> 16);
	}
	if (is_exc) {
		const long long k = atomicAdd(n_exc, 1);
		if (k < exc_cap) { exc[k].index = (u32)o; exc[k].x2_lo = (u32)v.x2; exc[k].x2_hi = (u32)(v.x2 >> 32); }
	}
}

__global__ void __launch_bounds__(128) compact_packed12_kernel(const Intv *__restrict__ slots, int slot_cap, const int *__restrict__ counts,
                                                               const long long *__restrict__ off, long long n, u32 *__restrict__ out,
                                                               u32 *__restrict__ off32, long long out_cap, int pos_bits, Exc12 *__restrict__ exc,
                                                               long long exc_cap, int *__restrict__ n_exc, int rec_bytes)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t >> 3;
	if (r >= n) return;
	const long long o0 = off[r];
	const int c = (int)min((long long)min(counts[r], slot_cap), max(out_cap - o0, 0ll));      // (see compact_kernel)
	if ((t & 7) == 0) { off32[r] = (u32)o0; if (r == n - 1) off32[n] = (u32)off[n]; }
	for (int e = (int)(t & 7); e < c; e += 8) {
		Intv v = ld_intv(&slots[(size_t)r * slot_cap + e]);
		v.info &= INFO_MASK;
		put_intv12(out, o0 + e, v, pos_bits, exc, exc_cap, n_exc, rec_bytes);
	}
}

__global__ void compact_list_packed12_kernel(const Intv *__restrict__ big_slots, int big_cap, const int *__restrict__ list,
                                             const int *__restrict__ counts_k, int n_list, const long long *__restrict__ off, u32 *__restrict__ out,
                                             int pos_bits, Exc12 *__restrict__ exc, long long exc_cap, int *__restrict__ n_exc, int rec_bytes)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int k = (int)(t / big_cap), e = (int)(t % big_cap);
	if (k >= n_list || e >= counts_k[k]) return;
	Intv v = ld_intv(&big_slots[(size_t)k * big_cap + e]);
	v.info &= INFO_MASK;
	put_intv12(out, off[list[k]] + e, v, pos_bits, exc, exc_cap, n_exc, rec_bytes);
}

// 12-byte layout with 24-bit third words -> byte-packed 11-byte records: four records (twelve words) in, eleven words out per thread
__global__ void __launch_bounds__(256) squeeze11_kernel(const u32 *__restrict__ in, const long long *__restrict__ total_ptr, u32 *__restrict__ out)
{
	const long long groups = (*total_ptr + 3) >> 2;
	for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (long long)gridDim.x * blockDim.x) {
		const uint4 a = reinterpret_cast<const uint4 *>(in)[3 * g], b = reinterpret_cast<const uint4 *>(in)[3 * g + 1], c = reinterpret_cast<const uint4 *>(in)[3 * g + 2];
		const u32 w0 = a.x, w1 = a.y, w2 = a.z & 0xffffffu, w3 = a.w, w4 = b.x, w5 = b.y & 0xffffffu, w6 = b.z, w7 = b.w, w8 = c.x & 0xffffffu, w9 = c.y, w10 = c.z,
		          w11 = c.w & 0xffffffu;
		u32 *o = out + 11 * g;
		o[0] = w0; o[1] = w1;
		o[2] = w2 | (w3 << 24);
		o[3] = (w3 >> 8) | (w4 << 24);
		o[4] = (w4 >> 8) | (w5 << 24);
		o[5] = (w5 >> 8) | (w6 << 16);
		o[6] = (w6 >> 16) | (w7 << 16);
		o[7] = (w7 >> 16) | (w8 << 16);
		o[8] = (w8 >> 16) | (w9 << 8);
		o[9] = (w9 >> 24) | (w10 << 8);
		o[10] = (w10 >> 24) | (w11 << 8);
	}
}

// dense 32-byte results of an earlier run -> 12-byte records + 32-bit offsets (smem_gpu_fetch_packed12 after smem_gpu_run_collect)
__global__ void intv12_from_dense_kernel(const Intv *__restrict__ in, long long total, const long long *__restrict__ off, long long n,
                                         u32 *__restrict__ out, u32 *__restrict__ off32, int pos_bits, Exc12 *__restrict__ exc, long long exc_cap,
                                         int *__restrict__ n_exc, int rec_bytes)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < total) put_intv12(out, t, ld_intv(&in[t]), pos_bits, exc, exc_cap, n_exc, rec_bytes);
	if (t <= n) off32[t] = (u32)off[t];
}

// The run's status words and interval total go to the host by a STORE into mapped pinned memory: a copy would queue on the
// device-to-host copy engine behind another lane's bulk result transfer and hold the GPU's kernel turn for its duration.
__global__ void publish_status_kernel(const int *__restrict__ status, const long long *__restrict__ total, int *__restrict__ host)
{
	const int t = threadIdx.x;
	if (t < 8) host[t] = status[t];
	if (t == 8) { const long long v = *total; host[8] = (int)(u32)v; host[9] = (int)(u32)((unsigned long long)v >> 32); }
	__threadfence_system();
}

// dense 32-byte results of an earlier run -> 16-byte records + 32-bit offsets (smem_gpu_fetch_packed after smem_gpu_run_collect)
__global__ void intv16_from_dense_kernel(const Intv *__restrict__ in, long long total, const long long *__restrict__ off, long long n,
                                         uint4 *__restrict__ out, u32 *__restrict__ off32)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < total) out[t] = pack_intv16(ld_intv(&in[t]));
	if (t <= n) off32[t] = (u32)off[t];
}

__global__ void compact_list_packed_kernel(const Intv *__restrict__ big_slots, int big_cap, const int *__restrict__ list,
                                           const int *__restrict__ counts_k, int n_list, const long long *__restrict__ off, uint4 *__restrict__ out)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int k = (int)(t / big_cap), e = (int)(t % big_cap);
	if (k >= n_list || e >= counts_k[k]) return;
	Intv v = ld_intv(&big_slots[(size_t)k * big_cap + e]);
	v.info &= INFO_MASK;
	out[off[list[k]] + e] = pack_intv16(v);
}

// Overflow re-run placement: big_slots[k][big_cap] of read list[k] -> dense output at off[list[k]].
__global__ void compact_list_kernel(const Intv *__restrict__ big_slots, int big_cap, const int *__restrict__ list,
                                    const int *__restrict__ counts_k, int n_list, const long long *__restrict__ off,
                                    Intv *__restrict__ out, unsigned short *__restrict__ step_out, unsigned short *__restrict__ aux_out)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int k = (int)(t / big_cap), e = (int)(t % big_cap);
	if (k >= n_list || e >= counts_k[k]) return;
	const Intv v = ld_intv(&big_slots[(size_t)k * big_cap + e]);
	const long long o = off[list[k]] + e;
	st_intv(&out[o], v.x0, v.x1, v.x2, v.info & INFO_MASK);
	if (step_out) step_out[o] = (unsigned short)(v.info >> STEP_SHIFT);
	if (aux_out) aux_out[o] = (unsigned short)(v.info >> AUX_SHIFT);
}

__global__ void add_base_kernel(long long *off, long long n, long long base)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) off[t] += base;
}

// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// Seed -> reference position: bwt_sa (bwt.c:104-114) for a batch of suffix-array rows.
// Each step is bwt_invPsi (bwt.c:71-77) = the symbol at the row + one bwt_occ (bwt.c:125-147): one occ block,
// fetched by a lane pair with one request like everywhere else; the walk ends at a sampled row (expected
// sa_intv steps).  `sa` are the samples of bwt_cal_sa (bwt.c:79-101), sa[0] = -1.
// One bwt_invPsi step (bwt.c:71-77), executed by a converged warp of lane pairs; pairs without work ride along.
__device__ __forceinline__ u64 inv_psi_pair(const DevIndex &ix, u64 k, int half, int lane)
{
	const bool at_primary = k == ix.primary;                       // bwt_invPsi returns 0 there (bwt.c:76)
	const u64 kk = at_primary ? 0 : k - (k >= ix.primary);          // '$' is not stored (bwt.c:133); == k - (k > primary) here
	u32 w[8];
	ld_sector(w, ix.blk + (kk >> 7) * 4 + half * 2);
	const int pos = (int)(kk & 127), bit = pos & 63, sh = 31 - (bit & 31);
	const u32 hw = bit < 32 ? w[4] : w[5], lw = bit < 32 ? w[6] : w[7];
	const int cm = (int)((((hw >> sh) & 1u) << 1) | ((lw >> sh) & 1u));       // bwt_B0 (bwt.h:78), valid in the half that owns pos
	const int c = __shfl_sync(FULL_MASK, cm, (lane & ~1) | (pos >> 6));
	const int r = min(max(pos + 1 - 64 * half, 0), 64);             // symbols of my half that count
	const u32 m0 = __funnelshift_rc(0u, 0xffffffffu, r), m1 = __funnelshift_rc(0u, 0xffffffffu, max(r - 32, 0));
	const u32 h0 = (c & 2) ? w[4] : ~w[4], h1 = (c & 2) ? w[5] : ~w[5], l0 = (c & 1) ? w[6] : ~w[6], l1 = (c & 1) ? w[7] : ~w[7];
	u32 cnt = __popc(h0 & l0 & m0) + __popc(h1 & l1 & m1);
	cnt += __shfl_xor_sync(FULL_MASK, cnt, 1);
	const u64 base_mine = (c & 1) ? ((u64)w[2] | ((u64)w[3] << 32)) : ((u64)w[0] | ((u64)w[1] << 32));
	const u64 base = __shfl_sync(FULL_MASK, base_mine, (lane & ~1) | (c >> 1));
	return at_primary ? 0 : ix.L2[c] + base + cnt;
}

// Both walkers below are persistent: a pair retires its row when it reaches a sampled row and immediately takes
// the next one, so that a warp never waits for its longest walk (walk lengths are geometric, mean sa_intv).
__global__ void __launch_bounds__(128) sa_kernel(const DevIndex ix, const u64 *__restrict__ sa, int sa_shift, long long n,
                                                 const u64 *__restrict__ ks, u64 *__restrict__ out, int *__restrict__ status)
{
	const int lane = threadIdx.x & 31, half = lane & 1;
	const long long npairs = ((long long)gridDim.x * blockDim.x) >> 1;
	long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1, cur = 0;
	const u64 mask = (1ull << sa_shift) - 1;
	bool have = false;
	u64 k = 0, steps = 0;
	for (;;) {
		for (;;) {                                      // retire / refill (divergent between pairs)
			if (!have) { if (q >= n) break; cur = q; q += npairs; k = ks[cur]; steps = 0; have = true; }
			if ((k & mask) != 0) break;
			if (!half) out[cur] = steps + sa[k >> sa_shift];
			have = false;
		}
		__syncwarp();
		if (!__any_sync(FULL_MASK, have)) break;
		const u64 kn = inv_psi_pair(ix, k, half, lane);
		if (have) { k = kn; if (++steps > (1ull << 24)) { if (!half) atomicAdd(&status[2], 1); k = 0; } }   // guard: corrupt index
	}
}

// Text T = forward + reverse complement, one base per nibble, eight bases per 32-bit word, first base lowest -- the
// layout of the staged reads (pack_reads_kernel), so that PH_UW_TEXT compares eight bases with one XOR.  `pac` is the
// reference's .pac layout of the forward strand (bntseq.c: four bases per byte, first base in the top bits).
__global__ void __launch_bounds__(256) pack_text_nib_kernel(const uint8_t *__restrict__ pac, long long l_pac, u32 *__restrict__ tw, long long n_words)
{
	const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= n_words) return;
	const long long n = 2 * l_pac;
	u32 v = 0;
#pragma unroll
	for (int k = 0; k < 8; ++k) {
		const long long pos = 8 * w + k;
		u32 base = 0;
		if (pos < l_pac) base = (pac[pos >> 2] >> ((~pos & 3) << 1)) & 3u;
		else if (pos < n) { const long long o = n - 1 - pos; base = 3u - ((pac[o >> 2] >> ((~o & 3) << 1)) & 3u); }
		v |= base << (4 * k);
	}
	tw[w] = v;
}

// 33-bit table entries, seven to a 32-byte sector: words 0..6 = low 32 bits, word 7 = the seven high bits.  One sector is
// what one lane gathers with one request, so a lookup costs the same as with 8-byte entries at 4.57 bytes per entry.
__device__ __forceinline__ void uw_put33(u32 *tab, u64 idx, u64 val)          // table zeroed beforehand; entries are written once
{
	const u64 sec = idx / 7; const u32 slot = (u32)(idx % 7);
	tab[sec * 8 + slot] = (u32)val;
	if (val >> 32) atomicOr(&tab[sec * 8 + 7], 1u << slot);
}
__device__ __forceinline__ u64 uw_get33(const u32 *tab, u64 idx)
{
	const u64 sec = idx / 7; const u32 slot = (u32)(idx % 7);
	return (u64)tab[sec * 8 + slot] | ((u64)((tab[sec * 8 + 7] >> slot) & 1u) << 32);
}

// Full suffix array and sampled inverse from the samples (unique-walk tables, PH_UW_* of seed_kernel): every sampled row k * intv
// starts a walk r -> invPsi(r), p -> p - 1 that ends at the next sampled row; the walks partition the rows, so every row
// r gets fsa[r] = its text position exactly once, and every position seq_len - e * 2^isa_shift gets isa[e] = its row.
// Row 0 (the '$' suffix, sa[0] = -1 in the reference, bwt.c:97) is given position seq_len.  Persistent lane pairs, one occ
// block per step like sa_kernel.
__global__ void __launch_bounds__(128) fsa_build_kernel(const DevIndex ix, const u64 *__restrict__ sa, int sa_shift, long long n_sa,
                                                        u32 *__restrict__ fsa, u32 *__restrict__ isa, int isa_shift, int *__restrict__ status)
{
	const int lane = threadIdx.x & 31, half = lane & 1;
	const long long npairs = ((long long)gridDim.x * blockDim.x) >> 1;
	long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1;
	const u64 mask = (1ull << sa_shift) - 1, imask = (1ull << isa_shift) - 1;
	bool have = false;
	u64 r = 0, pos = 0, steps = 0;
	auto record = [&]() {
		uw_put33(fsa, r, pos);
		const u64 back = ix.seq_len - pos;
		if ((back & imask) == 0) uw_put33(isa, back >> isa_shift, r);
	};
	for (;;) {
		if (!have && q < n_sa) {
			r = (u64)q << sa_shift; pos = q ? sa[q] : ix.seq_len; q += npairs;
			if (!half) record();
			have = true; steps = 0;
		}
		__syncwarp();
		if (!__any_sync(FULL_MASK, have)) break;
		const u64 rn = inv_psi_pair(ix, r, half, lane);
		if (have) {
			r = rn; pos -= 1;
			if ((r & mask) == 0) have = false;
			else {
				if (!half) record();
				if (++steps > (1ull << 24)) { if (!half) atomicAdd(&status[2], 1); have = false; }   // guard: corrupt index
			}
		}
	}
}

// test hook: 33-bit table -> uint64 array
__global__ void uw_unpack_kernel(const u32 *__restrict__ tab, long long n, u64 *__restrict__ out)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = uw_get33(tab, (u64)i);
}

// Intervals -> seeds (bwamem.c:408-421 / 462-476): an interval with seed length >= min_seed_len and x[2] <= max_occ
// contributes x[2] seeds {rbeg = bwt_sa(x[0] + t), qbeg, len}.
struct __align__(16) Seed { long long rbeg; int qbeg, len; };      // == mem_seed_t (bwamem.c:316-319)

__global__ void seed_count_kernel(const Intv *__restrict__ iv, long long total, int min_seed_len, u64 max_occ, int *__restrict__ cnt)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i > total) return;
	int c = 0;
	if (i < total) {
		const Intv v = ld_intv(&iv[i]);
		const int slen = (int)(u32)v.info - (int)(v.info >> 32);
		if (slen >= min_seed_len && v.x2 <= max_occ) c = (int)v.x2;
	}
	cnt[i] = c;                      // cnt[total] = 0: the scan then leaves the grand total in soff[total]
}

__global__ void __launch_bounds__(128) seed_expand_kernel(const DevIndex ix, const u64 *__restrict__ sa, int sa_shift, const Intv *__restrict__ iv,
                                                          long long total, const long long *__restrict__ soff, long long n_seeds,
                                                          Seed *__restrict__ out, int *__restrict__ status)
{
	const int lane = threadIdx.x & 31, half = lane & 1;
	const long long npairs = ((long long)gridDim.x * blockDim.x) >> 1;
	long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 1, cur = 0;
	const u64 mask = (1ull << sa_shift) - 1;
	bool have = false;
	u64 k = 0, steps = 0;
	int qbeg = 0, len = 0;
	for (;;) {
		for (;;) {                                      // retire / refill (divergent between pairs)
			if (!have) {
				if (q >= n_seeds) break;
				cur = q; q += npairs;
				long long lo = 0, hi = total - 1;           // interval of seed `cur`: last i with soff[i] <= cur
				while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (soff[mid] <= cur) lo = mid; else hi = mid - 1; }
				const Intv v = ld_intv(&iv[lo]);
				k = v.x0 + (u64)(cur - soff[lo]);           // bwamem.c:420: bwt_sa(bwt, p->x[0] + k)
				qbeg = (int)(v.info >> 32); len = (int)(u32)v.info - qbeg;
				steps = 0; have = true;
			}
			if ((k & mask) != 0) break;
			if (!half) { Seed sd; sd.rbeg = (long long)(steps + sa[k >> sa_shift]); sd.qbeg = qbeg; sd.len = len; out[cur] = sd; }
			have = false;
		}
		__syncwarp();
		if (!__any_sync(FULL_MASK, have)) break;
		const u64 kn = inv_psi_pair(ix, k, half, lane);
		if (have) { k = kn; if (++steps > (1ull << 24)) { if (!half) atomicAdd(&status[2], 1); k = 0; } }
	}
}

// The same expansion when the unique-walk tables are resident: bwt_sa(k) is ONE gather from the full suffix array (33-bit entries,
// uw_get33) instead of a walk of sa_intv / 2 bwt_invPsi steps on average.  One thread per seed.
__global__ void __launch_bounds__(256) seed_expand_fsa_kernel(const u32 *__restrict__ fsa, const Intv *__restrict__ iv, long long total,
                                                              const long long *__restrict__ soff, long long n_seeds, Seed *__restrict__ out)
{
	const long long cur = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (cur >= n_seeds) return;
	long long lo = 0, hi = total - 1;               // interval of seed `cur`: last i with soff[i] <= cur
	while (lo < hi) { const long long mid = (lo + hi + 1) >> 1; if (soff[mid] <= cur) lo = mid; else hi = mid - 1; }
	const Intv v = ld_intv(&iv[lo]);
	Seed sd;
	sd.rbeg = (long long)uw_get33(fsa, v.x0 + (u64)(cur - soff[lo]));       // bwamem.c:420: bwt_sa(bwt, p->x[0] + k)
	sd.qbeg = (int)(v.info >> 32); sd.len = (int)(u32)v.info - sd.qbeg;
	out[cur] = sd;
}

// bwt_sa for a batch of rows from the full suffix array (row 0, the '$' suffix, is -1 in the reference, bwt.c:97)
__global__ void sa_from_fsa_kernel(const u32 *__restrict__ fsa, long long n, const u64 *__restrict__ k, u64 *__restrict__ out)
{
	const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = k[i] == 0 ? ~0ull : uw_get33(fsa, k[i]);
}

__global__ void seed_read_off_kernel(const long long *__restrict__ off, const long long *__restrict__ soff, long long n, long long base,
                                     long long *__restrict__ seed_off)
{
	const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (r <= n) seed_off[r] = soff[off[r]] + base;
}

// Random-access roofline probe: every thread walks a dependent chain of
// aligned BYTES-sized gathers over `n_units` units; the next address is a hash of the data just
// loaded, so nothing can be prefetched or coalesced -- the access pattern of bwt_occ4.
// VARIANT selects the load flavour so that the best the memory system can do is what gets reported:
//   0: 256-bit ld.global.nc.L1::no_allocate   1: 128-bit ld.global.nc.L1::no_allocate
//   2: 256-bit plain ld.global (L1 allocating) 3: 128-bit ld.global.cg   4: 128-bit ld.global.nc + L2::evict_first
template <int VARIANT>
__device__ __forceinline__ u32 probe_load32(const uint4 *ptr, u64 pol)
{
	u32 w[8];
	if (VARIANT == 0) {
		asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
	} else if (VARIANT == 2) {
		asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
	} else if (VARIANT == 1) {
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(ptr));
		asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr + 1));
	} else if (VARIANT == 3) {
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(ptr));
		asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr + 1));
	} else {
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "l"(ptr), "l"(pol));
		asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr + 1), "l"(pol));
	}
	return w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
}

template <int BYTES, int VARIANT>
__global__ void __launch_bounds__(256) gather_probe_kernel(const uint4 *__restrict__ base, u64 n_units, int steps, u64 *sink)
{
	u64 s = (u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	u64 acc = 0, pol = 0;
	if (VARIANT == 4) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
	for (int it = 0; it < steps; ++it) {
		s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
		const u64 u = __umul64hi(s, n_units);          // uniform in [0, n_units) without a 64-bit modulo
		const uint4 *ptr = base + u * (BYTES / 16);
		u32 f = probe_load32<VARIANT>(ptr, pol);
		if (BYTES >= 64) f ^= probe_load32<VARIANT>(ptr + 2, pol);
		if (BYTES >= 128) { f ^= probe_load32<VARIANT>(ptr + 4, pol); f ^= probe_load32<VARIANT>(ptr + 6, pol); }
		acc += f;
		s += f;          // the chain: next unit depends on the bytes just gathered
	}
	if (acc == 0x7fffffffffffffffull) *sink = acc;
}

// Cooperative flavour of the probe: LANES adjacent lanes share one chain and fetch one aligned
// (LANES * WIDTH)-byte unit with a single load instruction (WIDTH = 16 or 32 bytes per lane), the
// access shape of a sub-warp-per-read kernel.  Chains per SM = resident threads / LANES.
template <int LANES, int WIDTH>
__global__ void __launch_bounds__(256) gather_probe_coop_kernel(const uint4 *__restrict__ base, u64 n_units, int steps, u64 *sink)
{
	const int tid = blockIdx.x * blockDim.x + threadIdx.x;
	const int sub = tid % LANES;
	u64 s = (u64)(tid / LANES) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	u64 acc = 0;
	for (int it = 0; it < steps; ++it) {
		s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
		const u64 u = __umul64hi(s, n_units);
		const uint4 *ptr = base + u * (LANES * WIDTH / 16) + sub * (WIDTH / 16);
		u32 f;
		if (WIDTH == 32) f = probe_load32<0>(ptr, 0);
		else {
			u32 a, b, c, d;
			asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(ptr));
			f = a ^ b ^ c ^ d;
		}
#pragma unroll
		for (int m = 1; m < LANES; m <<= 1) f ^= __shfl_xor_sync(0xffffffffu, f, m);
		acc += f;
		s += f;
	}
	if (acc == 0x7fffffffffffffffull) *sink = acc;
}

// Bulk-copy flavour of the probe ("TMA ... behind an mbarrier"): every lane fetches its 64-byte unit with
// ONE cp.async.bulk (global -> shared, completion on the warp's mbarrier) instead of LDG, then reads it back from shared
// memory.  Answers two questions for the seed kernel: does the copy engine path sustain more DRAM-missing requests than
// 38 G/s, and does a 64-byte unit fetched by a single lane cost one request (a lane pair's LDG.256) or two (a lane's 2 x 32 B)?
__global__ void __launch_bounds__(256) gather_probe_bulk_kernel(const uint4 *__restrict__ base, u64 n_units, int steps, u64 *sink)
{
	__shared__ __align__(128) uint4 buf[256 * 4];                  // 64 bytes per thread
	__shared__ __align__(8) unsigned long long bar[8];             // one mbarrier per warp
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const u32 bar_a = (u32)__cvta_generic_to_shared(&bar[warp]);
	const u32 slot_a = (u32)__cvta_generic_to_shared(&buf[threadIdx.x * 4]);
	if (lane == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	__syncwarp();
	u64 s = (u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	u64 acc = 0;
	u32 phase = 0;
	for (int it = 0; it < steps; ++it) {
		s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
		const u64 u = __umul64hi(s, n_units);
		const uint4 *ptr = base + u * 4;
		if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(32u * 64u) : "memory");
		__syncwarp();
		asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 64, [%2];"
		             :: "r"(slot_a), "l"(ptr), "r"(bar_a) : "memory");
		u32 done = 0;
		while (!done) {
			asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
			             : "=r"(done) : "r"(bar_a), "r"(phase) : "memory");
		}
		phase ^= 1;
		const uint4 a = buf[threadIdx.x * 4], b = buf[threadIdx.x * 4 + 1], c = buf[threadIdx.x * 4 + 2], d = buf[threadIdx.x * 4 + 3];
		const u32 f = a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c.x ^ c.y ^ c.z ^ c.w ^ d.x ^ d.y ^ d.z ^ d.w;
		acc += f;
		s += f;
		__syncwarp();                                               // every lane has read its slot before the next round overwrites it
	}
	if (acc == 0x7fffffffffffffffull) *sink = acc;
}
