// smem_kernels.cuh -- the seeding kernel and its small companions (sm_100a).
#pragma once
#include "smem_device.cuh"

enum { MODE_COLLECT = 0, MODE_SMEM1 = 1 };
enum { PH_NEED_READ = 0, PH_NEXT_STEP, PH_INIT_CALL, PH_FWD, PH_FWD_DONE, PH_BWD, PH_CALL_DONE, PH_EXIT };

#define STEP_SHIFT 48        // inside the slots the smem_next2 step index rides in info bits 48..63
#define SEED_BLOCK 128       // threads per CTA of the seeding kernel

// One persistent thread = one read at a time.  Per-thread scratch (global, L1/L2 cached):
//   B0, B1 : prev/curr ping-pong of bwt_smem1 (B0 also receives the forward pass' pushes)
//   M1, M2 : `matches` / `sub` of smem_next2 in emission order (descending start)
// each of `scratch_cap` = max_read_len + 2 entries, which bounds every list of bwt.c:776-835
// (forward pushes <= len - x, backward emissions <= x + 2), so there is no overflow path.
template <int MODE, int MIN_BLOCKS>
__global__ void __launch_bounds__(SEED_BLOCK, MIN_BLOCKS) seed_kernel(const SeedParams p)
{
	const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
	Intv *const B0 = p.scratch + (size_t)gtid * 4 * p.scratch_cap;
	Intv *const B1 = B0 + p.scratch_cap, *const M1 = B1 + p.scratch_cap, *const M2 = M1 + p.scratch_cap;

	u64 policy = 0;
	if (p.hot_min_intv) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));

	int phase = PH_NEED_READ;
	// ---- per-read state
	int rk = 0, rid = 0;                 // position in the work list, read id
	const uint8_t *q = nullptr;
	int len = 0, start = 0, step = 0, ori_start = 0, split_len = 0, n_out = 0;
	Intv *slot = nullptr;
	// ---- per-call state (one bwt_smem1)
	int pass = 0, x = 0, i = 0, j = 0, n_prev = 0, n_curr = 0, n_mem = 0, ret = 0, cb = -1;
	u64 min_intv = 1;
	bool prev_rev = false;
	Intv *prev = B0, *curr = B1, *M = M1;
	u64 ka = 0, kb = 0, ks = 0;          // forward ik kept as (a = x[1], b = x[0], s = x[2]): is_back = 0 walks x[1]
	int kend = 0;
	u64 last_curr_s = 0;
	int last_mem_start = 0;
	int n_m1 = 0, keep_len = 0;          // |matches| and its longest length once pass 0 finished
	int max_len = 0, max_start = 0, max_end = 0;   // longest SMEM of this call (first max in start order)
	u64 max_s = 0;
	Intv pe;                             // current backward element prev[j]
	pe.x0 = pe.x1 = pe.x2 = pe.info = 0;
	int guard = 0;

	for (;;) {
		// ============================================================== bookkeeping (divergent, cheap)
		bool req = false;
		while (!req && phase != PH_EXIT) {
			switch (phase) {
			case PH_NEED_READ: {
				rk = atomicAdd(&p.status[0], 1);
				if ((long long)rk >= p.n) { phase = PH_EXIT; break; }
				rid = p.list ? p.list[rk] : rk;
				const long long o0 = p.offs[rid];
				q = p.seq + o0;
				len = (int)(p.offs[rid + 1] - o0);
				slot = p.slots + (size_t)rk * p.slot_cap;
				n_out = 0; start = 0; step = 0;
				if (MODE == MODE_COLLECT) {
					split_len = p.split_len_init < len ? p.split_len_init : len;     // bwamem.c:458
					phase = PH_NEXT_STEP;
				} else {
					x = p.xs[rid];
					const int mi = p.min_intvs[rid];
					min_intv = mi < 1 ? 1 : (u64)mi;                                 // bwt.c:784
					pass = 0; M = M1;
					if (x < 0 || x >= len || q[x] > 3) {                              // bwt.c:783
						p.ret[rid] = x + 1; p.counts[rk] = 0; phase = PH_NEED_READ;
					} else phase = PH_INIT_CALL;
				}
			} break;
			case PH_NEXT_STEP: {     // head of smem_next2, bwamem.c:249-258
				while (start < len && q[start] > 3) ++start;
				if (start >= len) {
					p.counts[rk] = n_out;
					if (n_out > p.slot_cap) { p.overflow_list[atomicAdd(&p.status[1], 1)] = rid; atomicMax(&p.status[3], n_out); }
					phase = PH_NEED_READ;
					break;
				}
				ori_start = start; x = start; pass = 0; M = M1;
				min_intv = p.start_width < 1 ? 1 : (u64)p.start_width;
				phase = PH_INIT_CALL;
			} break;
			case PH_INIT_CALL: {     // bwt_set_intv, bwt.c:788-789
				const int c0 = q[x];
				ka = p.ix.L2[3 - c0] + 1;
				kb = p.ix.L2[c0] + 1;
				ks = p.ix.L2[c0 + 1] - p.ix.L2[c0];
				kend = x + 1;
				i = x + 1; n_curr = 0; n_mem = 0;
				max_len = 0; max_start = max_end = 0; max_s = 0;
				guard = 2 * (len + 2) * (len + 2) + 64;   // > every extend one bwt_smem1 can issue
				phase = PH_FWD;
			} break;
			case PH_FWD: {
				if (i == len || q[i] > 3) {              // bwt.c:800-803 (ambiguous base) / :806 (end of read)
					st_intv(&B0[n_curr++], kb, ka, ks, (u64)kend);
					ret = kend;
					phase = PH_FWD_DONE;
				} else req = true;
			} break;
			case PH_FWD_DONE: {      // bwt.c:807-809: "reverse curr" == walk B0 from its tail
				if (MODE == MODE_COLLECT && pass == 0) start = ret;   // bwamem.c:262; pass 2 leaves start alone
				prev = B0; curr = B1; prev_rev = true; n_prev = n_curr; n_curr = 0;
				i = x - 1; j = 0;
				cb = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
				phase = PH_BWD;
			} break;
			case PH_BWD: {
				pe = ld_intv(&prev[prev_rev ? n_prev - 1 - j : j]);
				if (cb < 0) {
					// bwt.c:815-821 with c == -1 (read start or ambiguous base): nothing can enter curr and
					// only prev[0] can pass the containment test, so the round collapses to this.
					if (n_mem == 0 || i + 1 < last_mem_start) {
						st_intv(&M[n_mem++], pe.x0, pe.x1, pe.x2, pe.info | ((u64)(i + 1) << 32));
						last_mem_start = i + 1;
						const int l = (int)pe.info - (i + 1);
						if (l >= max_len) { max_len = l; max_start = i + 1; max_end = (int)pe.info; max_s = pe.x2; }
					}
					phase = PH_CALL_DONE;
				} else req = true;
			} break;
			case PH_CALL_DONE: {
				if (MODE == MODE_SMEM1) {
					for (int e = n_mem - 1, o = 0; e >= 0; --e, ++o)             // bwt.c:829: ascending start
						if (o < p.slot_cap) { const Intv v = ld_intv(&M1[e]); st_intv(&slot[o], v.x0, v.x1, v.x2, v.info); }
					p.counts[rk] = n_mem; p.ret[rid] = ret;
					if (n_mem > p.slot_cap) { p.overflow_list[atomicAdd(&p.status[1], 1)] = rid; atomicMax(&p.status[3], n_mem); }
					phase = PH_NEED_READ;
					break;
				}
				const u64 tag = (u64)step << STEP_SHIFT;
				if (pass == 0) {
					n_m1 = n_mem; keep_len = max_len;
					// bwamem.c:272: re-seed from the middle of the longest SMEM if it is long and (nearly) unique
					if (n_mem > 0 && split_len > 0 && max_len >= split_len && max_s <= (u64)p.split_width) {
						pass = 1; M = M2;
						x = (max_end + max_start) >> 1;
						min_intv = max_s + 1;
						phase = PH_INIT_CALL;
						break;
					}
					for (int e = n_m1 - 1; e >= 0; --e) {
						if (n_out < p.slot_cap) { const Intv v = ld_intv(&M1[e]); st_intv(&slot[n_out], v.x0, v.x1, v.x2, v.info | tag); }
						++n_out;
					}
				} else {
					// ordered merge, bwamem.c:281-301; both lists are walked in ascending start = reverse emission
					int a = n_m1 - 1, b = n_mem - 1;
					const int half = keep_len >> 1;
					Intv va, vb;
					va.x0 = va.x1 = va.x2 = va.info = 0; vb = va;
					bool have_a = false, have_b = false;
					while (a >= 0 || b >= 0) {
						if (a >= 0 && !have_a) { va = ld_intv(&M1[a]); have_a = true; }
						if (b >= 0 && !have_b) { vb = ld_intv(&M2[b]); have_b = true; }
						bool take_a;
						if (a >= 0 && b >= 0) {
							const long long xa = (long long)((va.info >> 32 << 32) | (u32)(len - (int)(u32)va.info));
							const long long xb = (long long)((vb.info >> 32 << 32) | (u32)(len - (int)(u32)vb.info));
							take_a = xa < xb;
						} else take_a = a >= 0;
						if (take_a) {
							if (n_out < p.slot_cap) st_intv(&slot[n_out], va.x0, va.x1, va.x2, va.info | tag);
							++n_out; --a; have_a = false;
						} else {
							const int sl = (int)(u32)vb.info - (int)(vb.info >> 32);
							if (sl >= half && (int)(u32)vb.info > ori_start) {
								if (n_out < p.slot_cap) st_intv(&slot[n_out], vb.x0, vb.x1, vb.x2, vb.info | tag);
								++n_out;
							}
							--b; have_b = false;
						}
					}
				}
				++step;
				phase = PH_NEXT_STEP;
			} break;
			default: break;
			}
		}
		if (phase == PH_EXIT) break;

		// ============================================================== one bwt_extend (warp re-converges here)
		Ext ok;
		if (phase == PH_FWD) ok = extend(p.ix, ka, kb, ks, 3 - (int)q[i], p.hot_min_intv, policy);   // bwt.c:793
		else                 ok = extend(p.ix, pe.x0, pe.x1, pe.x2, cb, p.hot_min_intv, policy);     // bwt.c:812
		if (--guard < 0) { atomicAdd(&p.status[2], 1); p.counts[rk] = 0; phase = PH_NEED_READ; continue; }

		// ============================================================== consume the result
		if (phase == PH_FWD) {                               // bwt.c:794-799
			if (ok.s != ks) {
				st_intv(&B0[n_curr++], kb, ka, ks, (u64)kend);
				ret = kend;
				if (ok.s < min_intv) { phase = PH_FWD_DONE; continue; }
			}
			ka = ok.a; kb = ok.b; ks = ok.s; kend = i + 1;
			++i;
		} else {                                             // bwt.c:813-824
			if (ok.s < min_intv) {
				if (n_curr == 0 && (n_mem == 0 || i + 1 < last_mem_start)) {
					st_intv(&M[n_mem++], pe.x0, pe.x1, pe.x2, pe.info | ((u64)(i + 1) << 32));
					last_mem_start = i + 1;
					const int l = (int)pe.info - (i + 1);
					if (l >= max_len) { max_len = l; max_start = i + 1; max_end = (int)pe.info; max_s = pe.x2; }
				}
			} else if (n_curr == 0 || ok.s != last_curr_s) {
				st_intv(&curr[n_curr++], ok.a, ok.b, ok.s, pe.info);
				last_curr_s = ok.s;
			}
			if (++j == n_prev) {                             // bwt.c:826-827
				if (n_curr == 0) phase = PH_CALL_DONE;
				else {
					Intv *t = prev; prev = curr; curr = t;
					prev_rev = false;
					n_prev = n_curr; n_curr = 0; j = 0; --i;
					cb = i < 0 ? -1 : (q[i] < 4 ? (int)q[i] : -1);
				}
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------
// counts -> CSR offsets is done with one CUB exclusive scan on the host side (plumbing).
// Compaction: slots[n][slot_cap] -> dense intv[total] (+ optional step array), info restored.
__global__ void compact_kernel(const Intv *__restrict__ slots, int slot_cap, const int *__restrict__ counts,
                               const long long *__restrict__ off, long long n, Intv *__restrict__ out,
                               unsigned short *__restrict__ step_out)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const long long r = t / slot_cap;
	const int e = (int)(t - r * slot_cap);
	if (r >= n) return;
	const int c = counts[r];
	if (e >= c) return;          // entries beyond slot_cap are filled in by the overflow re-run
	const Intv v = ld_intv(&slots[(size_t)r * slot_cap + e]);
	const long long o = off[r] + e;
	st_intv(&out[o], v.x0, v.x1, v.x2, v.info & ((1ull << STEP_SHIFT) - 1));
	if (step_out) step_out[o] = (unsigned short)(v.info >> STEP_SHIFT);
}

// Overflow re-run placement: big_slots[k][big_cap] of read list[k] -> dense output at off[list[k]].
__global__ void compact_list_kernel(const Intv *__restrict__ big_slots, int big_cap, const int *__restrict__ list,
                                    const int *__restrict__ counts_k, int n_list, const long long *__restrict__ off,
                                    Intv *__restrict__ out, unsigned short *__restrict__ step_out)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	const int k = (int)(t / big_cap), e = (int)(t % big_cap);
	if (k >= n_list || e >= counts_k[k]) return;
	const Intv v = ld_intv(&big_slots[(size_t)k * big_cap + e]);
	const long long o = off[list[k]] + e;
	st_intv(&out[o], v.x0, v.x1, v.x2, v.info & ((1ull << STEP_SHIFT) - 1));
	if (step_out) step_out[o] = (unsigned short)(v.info >> STEP_SHIFT);
}

__global__ void add_base_kernel(long long *off, long long n, long long base)
{
	const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
	if (t < n) off[t] += base;
}

// ---------------------------------------------------------------------------------------------
// Random-access roofline probe: every thread walks a dependent chain of
// aligned BYTES-sized gathers over `n_units` units; the next address is a hash of the data just
// loaded, so nothing can be prefetched or coalesced -- the access pattern of bwt_occ4.
template <int BYTES>
__global__ void __launch_bounds__(256) gather_probe_kernel(const uint4 *__restrict__ base, u64 n_units, int steps, u64 *sink)
{
	u64 s = (u64)(blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull + 0x1234567ull;
	u64 acc = 0;
	for (int it = 0; it < steps; ++it) {
		s ^= s >> 29; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 32;
		const u64 u = s % n_units;
		const uint4 *ptr = base + u * (BYTES / 16);
		u32 w[16];
		asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
		             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(ptr));
		u32 f = w[0] ^ w[1] ^ w[2] ^ w[3] ^ w[4] ^ w[5] ^ w[6] ^ w[7];
		if (BYTES == 64) {
			asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
			             : "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15]) : "l"(ptr + 2));
			f ^= w[8] ^ w[9] ^ w[10] ^ w[11] ^ w[12] ^ w[13] ^ w[14] ^ w[15];
		}
		acc += f;
		s += f;          // the chain: next unit depends on the bytes just gathered
	}
	if (acc == 0x7fffffffffffffffull) *sink = acc;
}
