// smem_gpu.cu -- host side of the C ABI declared in include/smem_gpu.h.
//
// Replaces, for the SMEM path only, the reference's accelerator plumbing:
//   HelloALINLB.cpp:309-451 (service + workspace allocation, CSR programming)  -> smem_gpu_create
//   bwa.c:289-301           (index memcpy into the workspace + handshake 2)    -> smem_gpu_upload_index
//   bwt.c:559-749 + fastmap.c:335-420 (pack, handshake, manager copy, unpack)  -> stage / run / fetch
// No CPU fallback exists here: every failure is an error code.
#include "../../include/smem_gpu.h"
#include "smem_kernels.cuh"
#include "smem_chain.cuh"
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct DeviceCtx {
	int dev = 0, sm_count = 0;
	size_t smem_per_sm = 0, smem_per_block_optin = 0;
	cudaStream_t stream = nullptr;
	cudaEvent_t ev0 = nullptr, ev0s = nullptr, ev1 = nullptr, ev2 = nullptr;   // call start | seed kernel start | seed kernel end | call end
	cudaEvent_t ev_sync = nullptr;   // blocking-sync event: host waits sleep instead of spinning (see stream_wait)
	// index
	uint4 *d_index = nullptr;
	uint4 *d_sec = nullptr;          // the index as 32-byte sectors with 32-bit checkpoints (one lane per read); owned with d_index
	size_t index_bytes = 0;
	DevIndex ix{};
	bool has_index = false;
	bool owns_index = true;          // false: d_index aliases the copy of an earlier context on the same GPU
	// suffix-array samples and the seed-level API (section 8f-1)
	u64 *d_sa = nullptr; bool owns_sa = true; int sa_shift = -1; u64 n_sa = 0;
	// repeat filter of the re-seeding pass (smem_repeat.cuh)
	u32 *d_rf = nullptr; bool owns_rf = true; int rf_k = 0, rf_log2 = 0;
	// unique-walk tables (PH_UW_* of seed_kernel): nibble text, full suffix array, inverse suffix array
	u64 *d_uw_text = nullptr; u32 *d_fsa = nullptr, *d_isa = nullptr; bool owns_uw = true; u64 uw_text_len = 0; int uw_isa_shift = 2; size_t uw_bytes = 0;
	u64 rf_text_len = 0;             // length of the text the filter was built from: it is only used with an index of that seq_len
	int64_t pass2_skipped = 0, uw_walks = 0;
	uint64_t turn_epoch = 0;         // epoch of the last run in which this lane took part in the kernel turn
	bool holds_turn = false;         // this lane's seed kernel is part of the call that currently owns the GPU's kernel turn
	u64 *d_k = nullptr, *d_kout = nullptr; size_t k_cap = 0;
	int *d_scnt = nullptr; long long *d_soff = nullptr, *d_sroff = nullptr; size_t s_cap = 0, sroff_cap = 0;
	Seed *d_seeds = nullptr; size_t seeds_cap = 0; long long n_seeds = 0;
	bool seeds_valid = false;        // d_seeds / d_soff describe the resident intervals (set by smem_gpu_seeds, cleared by a new run)
	// chaining (section 8f-3): scratch per seed slot, outputs
	int *d_cwork = nullptr; FltRec *d_flt = nullptr; unsigned char *d_keep = nullptr; size_t cwork_cap = 0;
	int *d_bt = nullptr; size_t bt_nodes = 0;   // B-tree nodes of the chain builder (smem_chain.cuh)
	int *d_nch = nullptr, *d_nkept = nullptr; long long *d_coff = nullptr, *d_koff = nullptr; size_t cread_cap = 0;
	Chain *d_chains = nullptr; Seed *d_cseeds = nullptr; size_t chains_cap = 0;
	long long n_chains = 0, n_cseeds = 0;
	float chain_ms = 0;
	// pipeline lanes
	int lane = 0, lanes_on_dev = 1;  // pipeline lane of this context on its GPU
	DeviceCtx *prev_lane = nullptr;  // the lane whose seed kernel runs right before this one's
	struct smem_gpu *owner = nullptr;
	uint64_t seed_issued = 0;        // epoch of the last run whose seed kernel + ev1 have been enqueued (guarded by smem_gpu::lane_mu)
	uint64_t stage_issued = 0;       // epoch of the last staging whose H2D copies have been enqueued (same guard)
	// batch buffers (capacity fixed at create)
	int64_t read_cap = 0;
	size_t seq_cap = 0;
	uint8_t *d_seq = nullptr;        // the caller's reads as they arrived: CSR bytes (in_fmt 0) or 2-bit records (in_fmt 1)
	uint4 *d_qpack = nullptr; size_t qpack_bytes = 0;   // reads re-packed one base per nibble at a fixed stride (pack pre-pass)
	int *d_rlen = nullptr;           // read lengths (pack pre-pass)
	long long *d_offs = nullptr;
	unsigned short *d_lens = nullptr; AmbEntry *d_amb = nullptr; size_t amb_cap = 0;   // compact wire format: lengths, ambiguous-base list
	int in_fmt = 0, r2_stride = 0, r2_read_len = 0; bool r2_has_lens = false; long long n_amb = 0;
	uint4 *d_outp = nullptr; size_t outp_cap = 0; u32 *d_off32 = nullptr;               // 16- or 12-byte result records + 32-bit offsets
	Exc12 *d_exc = nullptr; size_t exc_cap = 0; long long n_exc = 0; int pos_bits = 0;  // 12-byte records: intervals whose size does not fit the record
	int want_packed = 0;             // result form of the next run: 0 = 32-byte bwtintv_t, 1 = 16-byte records, 2 = 12-byte records
	int outp_fmt = 0;                // what d_outp holds (0 = nothing valid)
	bool out_valid = false;
	int64_t h2d = 0, d2h = 0;        // bytes of the last stage / fetch of this context
	double acc_stage_ms = 0, acc_turn_ms = 0, acc_run_ms = 0, acc_fetch_ms = 0; int64_t acc_calls = 0, acc_h2d = 0, acc_d2h = 0;   // wall-time accumulators of the one-call forms
	int *d_x = nullptr, *d_mi = nullptr, *d_ret = nullptr;
	int *d_counts = nullptr, *d_overflow = nullptr, *d_status = nullptr;
	long long *d_off = nullptr;
	Intv *d_slots = nullptr;
	int slots_cap_alloc = 0;
	Intv *d_scratch = nullptr;
	size_t scratch_entries = 0;
	Intv *d_out = nullptr;
	unsigned short *d_step = nullptr, *d_aux = nullptr;
	size_t out_cap = 0;
	void *d_tmp = nullptr;
	size_t tmp_bytes = 0;
	Intv *d_big = nullptr;
	size_t big_entries = 0;
	int *d_counts_k = nullptr;
	size_t counts_k_cap = 0;
	int *h_status = nullptr;          // pinned: [0..4] status, [6..7] int64 total
	// latency path of small one-call batches (TINY_MAX reads, do_collect): ONE H2D copy of a pinned blob, three launches, results
	// stored into a pinned arena by the last kernel -- 7 driver calls per batch instead of 25 (with 16 host threads calling at
	// once the driver's serialisation of those calls, not the GPU, set the latency)
	bool tiny = false, status_dirty = true;
	uint8_t *h_tiny_in = nullptr, *d_tiny_in = nullptr; size_t tiny_in_bytes = 0;
	uint8_t *h_arena = nullptr; size_t arena_entries = 0, arena_reads = 0;
	size_t tiny_o_x = 0, tiny_o_mi = 0, tiny_o_seq = 0;
	// current shard
	int64_t lo = 0, hi = 0, n = 0;
	long long seq_base = 0;
	long long total = 0;
	int mode = MODE_COLLECT;
	// last-run measurements
	float seed_ms = 0, total_ms = 0;
	int64_t launches = 0, overflow = 0;
	std::string err;
	int rc = 0;
};

} // namespace

// Host wait for everything enqueued on the context's stream.  cudaStreamSynchronize spins a core; with one process per GPU
// and several worker threads each (8 GPUs x 2 threads on a 16-core host) the spinning threads starve the ones that have
// work, so such hosts can route the wait through a blocking-sync event and sleep ("blocking_sync" = 1; measured on this box: each wake-up costs ~1.5 ms, so it is off by default).
static bool g_blocking_sync = false;
static inline cudaError_t stream_wait(DeviceCtx &d)
{
	if (!g_blocking_sync) return cudaStreamSynchronize(d.stream);
	cudaError_t e = cudaEventRecord(d.ev_sync, d.stream);
	return e != cudaSuccess ? e : cudaEventSynchronize(d.ev_sync);
}

struct smem_gpu {
	std::vector<DeviceCtx> devs;
	int64_t max_batch = 0;
	int max_len = 0;
	int64_t staged = -1;
	bool ran = false;
	int block_threads = SEED_BLOCK, blocks_per_sm = 9, slot_cap = 128, b_cap = 17;
	int64_t hot_min_intv = 0;
	int l2_mode = 0;                 // see SeedParams::l2_mode
	int repeat_filter = 1;           // use the repeat filter (if built) to skip void re-seeding passes in MODE_COLLECT
	int unique_walk = 1;             // use the unique-walk tables if smem_gpu_build_text_index built them
	int uw_isa_shift = 2;            // the inverse suffix array of the next smem_gpu_build_text_index is sampled every 2^this positions
	int uw_min_left = 8, uw_min_run = 3;   // ... for walks with at least this many read bases left, after this many extends of a unique interval
	int spec_walk = 1;               // speculative longest-only backward walk in pass-1 calls (smem_kernels.cuh PH_SPEC)
	int sa_from_tables = 1;          // bwt_sa from the full suffix array of the unique-walk tables when they are resident (one gather instead of a walk)
	int tiny_path = 1;               // batches of up to TINY_MAX reads through the one-call forms take the latency path (DeviceCtx::tiny)
	int lanes_per_read = 4;          // 4 = lane pairs on the 32-byte sector index, backward sweep split over the two lanes (default; wide indices take 2);
	                                 // 2 = lane pairs on the 64-byte blocks, 3 = lane pairs on the sector index, 1 = one lane per read on the sector index
	bool build_sectors = true;       // smem_gpu_upload_index also builds the sector form (when every checkpoint fits 32 bits); "lanes_per_read" = 2 before the upload turns it off
	int count_skips = 0;             // debug: count the skipped passes (one atomic each)
	int probe_variant = 0;
	int force_wide = 0;
	int spare_sms = 0;               // SMs the seed kernel leaves empty when a GPU has several lanes
	int chain_lanes = 0;             // 1: lane k's seed kernel waits for lane k-1's (no tail overlap)
	int64_t turn_min_reads = 16384;  // calls with fewer reads per lane do not take the GPU's kernel turn: their few CTAs run next to other calls' kernels
	bool use_turn = true;
	bool stage_wait = true;          // ctx_stage_inner waits for its H2D copies (always in the split form)
	int64_t h2d_bytes = 0, d2h_bytes = 0;
	uint64_t epoch = 0;              // bumped per run; orders the lanes' seed kernels
	uint64_t stage_epoch = 0;        // bumped per staging; orders the lanes' H2D copies
	std::mutex lane_mu;
	std::condition_variable lane_cv;
	std::string err;
};

namespace {

#define CK(call)                                                                                          \
	do {                                                                                                  \
		cudaError_t e_ = (call);                                                                          \
		if (e_ != cudaSuccess) {                                                                          \
			char b_[512];                                                                                 \
			snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
			d.err = b_;                                                                                   \
			return e_ == cudaErrorMemoryAllocation ? SMEM_GPU_E_NOMEM : SMEM_GPU_E_CUDA;                 \
		}                                                                                                 \
	} while (0)

template <typename T> int dev_alloc(DeviceCtx &d, T **p, size_t count)
{
	CK(cudaMalloc((void **)p, std::max<size_t>(count, 1) * sizeof(T)));
	return 0;
}

// Batch buffers of a context: everything whose size follows max_batch_reads / max_read_len (smem_gpu_create, smem_gpu_resize).
// Device memory per lane: read_cap * (max_len + slot_cap * 32 + 16 * 32 + ~50) bytes of batch buffers plus
// sm_count * 9 * 64 * 3 * (max_len + 2) * 32 bytes of per-pair list scratch (0.85 GB at 101 bp, 2.1 GB at 250 bp, 8.4 GB at 1024 bp).
void ctx_free_batch(DeviceCtx &d)
{
	cudaFree(d.d_k); cudaFree(d.d_kout); cudaFree(d.d_scnt); cudaFree(d.d_soff); cudaFree(d.d_sroff); cudaFree(d.d_seeds);
	cudaFree(d.d_bt); d.d_bt = nullptr; d.bt_nodes = 0;
	cudaFree(d.d_cwork); cudaFree(d.d_flt); cudaFree(d.d_keep); cudaFree(d.d_nch); cudaFree(d.d_nkept); cudaFree(d.d_coff); cudaFree(d.d_koff);
	cudaFree(d.d_chains); cudaFree(d.d_cseeds);
	cudaFree(d.d_seq); cudaFree(d.d_qpack); cudaFree(d.d_offs); cudaFree(d.d_rlen); cudaFree(d.d_lens); cudaFree(d.d_amb); cudaFree(d.d_outp); cudaFree(d.d_off32); cudaFree(d.d_exc);
	cudaFree(d.d_x); cudaFree(d.d_mi); cudaFree(d.d_ret);
	cudaFree(d.d_counts); cudaFree(d.d_overflow); cudaFree(d.d_off); cudaFree(d.d_slots);
	cudaFree(d.d_scratch); cudaFree(d.d_out); cudaFree(d.d_step); cudaFree(d.d_aux); cudaFree(d.d_tmp); cudaFree(d.d_big); cudaFree(d.d_counts_k);
	d.d_k = d.d_kout = nullptr; d.k_cap = 0; d.d_scnt = nullptr; d.d_soff = d.d_sroff = nullptr; d.s_cap = d.sroff_cap = 0; d.d_seeds = nullptr; d.seeds_cap = 0;
	d.d_cwork = nullptr; d.d_flt = nullptr; d.d_keep = nullptr; d.cwork_cap = 0; d.d_nch = d.d_nkept = nullptr; d.d_coff = d.d_koff = nullptr; d.cread_cap = 0;
	d.d_chains = nullptr; d.d_cseeds = nullptr; d.chains_cap = 0;
	d.d_seq = nullptr; d.d_qpack = nullptr; d.qpack_bytes = 0; d.d_offs = nullptr; d.d_rlen = nullptr; d.d_lens = nullptr; d.d_amb = nullptr; d.amb_cap = 0;
	d.d_outp = nullptr; d.outp_cap = 0; d.d_off32 = nullptr; d.d_exc = nullptr; d.exc_cap = 0; d.d_x = d.d_mi = d.d_ret = nullptr; d.d_counts = d.d_overflow = nullptr; d.d_off = nullptr;
	d.d_slots = nullptr; d.slots_cap_alloc = 0; d.d_scratch = nullptr; d.scratch_entries = 0; d.d_out = nullptr; d.d_step = d.d_aux = nullptr; d.out_cap = 0;
	d.d_tmp = nullptr; d.tmp_bytes = 0; d.d_big = nullptr; d.big_entries = 0; d.d_counts_k = nullptr; d.counts_k_cap = 0;
	d.read_cap = 0; d.seq_cap = 0; d.n = 0; d.lo = d.hi = 0; d.seeds_valid = d.out_valid = false; d.outp_fmt = 0;
}

int ctx_alloc_batch(DeviceCtx &d, int64_t read_cap, int max_len, int slot_cap)
{
	CK(cudaSetDevice(d.dev));
	d.read_cap = read_cap;
	d.seq_cap = (size_t)read_cap * (size_t)max_len + 64;
	int rc;
	if ((rc = dev_alloc(d, &d.d_seq, d.seq_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_offs, (size_t)read_cap + 1))) return rc;
	if ((rc = dev_alloc(d, &d.d_rlen, (size_t)read_cap + 1))) return rc;
	if ((rc = dev_alloc(d, &d.d_x, (size_t)read_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_mi, (size_t)read_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_ret, (size_t)read_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_counts, (size_t)read_cap + 1))) return rc;
	if ((rc = dev_alloc(d, &d.d_overflow, (size_t)read_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_off, (size_t)read_cap + 1))) return rc;
	if ((rc = dev_alloc(d, &d.d_slots, (size_t)read_cap * slot_cap))) return rc;
	d.slots_cap_alloc = slot_cap;
	d.out_cap = (size_t)read_cap * 24 + 1024;      // (grown on demand; TRACE keeps ~19 raw entries per 101 bp read, COLLECT 11)
	if ((rc = dev_alloc(d, &d.d_out, d.out_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_step, d.out_cap))) return rc;
	if ((rc = dev_alloc(d, &d.d_aux, d.out_cap))) return rc;
	// per-pair scratch for the default launch geometry (re-grown in ctx_run if blocks_per_sm is raised): allocating
	// it lazily would delay the first lane's first kernel by a cudaMalloc
	d.scratch_entries = (size_t)std::min<int64_t>((int64_t)d.sm_count * 9, (read_cap + 63) / 64 + 1) * 64 * 3 * (size_t)(max_len + 2);   // (a batch never needs more CTAs than it has groups of 64 reads)
	CK(cudaMalloc((void **)&d.d_scratch, d.scratch_entries * sizeof(Intv)));
	// block sums of the counts -> offsets scan
	d.tmp_bytes = ((size_t)(read_cap + 1) / SCAN_PER_BLOCK + 2) * sizeof(long long);
	CK(cudaMalloc(&d.d_tmp, d.tmp_bytes));
	{   // reads re-packed one base per nibble | window flags of the repeat filter (sizes as in ctx_run_inner)
		const size_t q_stride = (size_t)(((max_len + 1) / 2 + 15) / 16 * 16);
		d.qpack_bytes = (size_t)read_cap * q_stride + (size_t)read_cap * (q_stride >> 4) * 4;
		CK(cudaMalloc((void **)&d.d_qpack, std::max<size_t>(d.qpack_bytes, 16)));
	}
	return 0;
}

int ctx_init(DeviceCtx &d, int dev, int lane, int64_t read_cap, int max_len, int slot_cap)
{
	d.dev = dev;
	CK(cudaSetDevice(dev));
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, dev));
	d.sm_count = prop.multiProcessorCount;
	d.smem_per_sm = prop.sharedMemPerMultiprocessor;
	d.smem_per_block_optin = prop.sharedMemPerBlockOptin;
	{
		// lanes of one GPU: earlier lane = higher stream priority.  Pending seed kernels then start in lane order
		// (the next lane's CTAs fill the tail of the previous one) and a finished lane's scan / compaction kernels
		// are scheduled ahead of the following lanes' persistent CTAs (measured: with equal priorities they starve
		// until every seed kernel has drained), so its D2H copy overlaps the next lane's kernel.
		int lo = 0, hi = 0;
		CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));          // lo = least (numerically largest), hi = greatest
		CK(cudaStreamCreateWithPriority(&d.stream, cudaStreamNonBlocking, std::min(lo, hi + lane)));
	}
	CK(cudaEventCreate(&d.ev0)); CK(cudaEventCreate(&d.ev0s)); CK(cudaEventCreate(&d.ev1)); CK(cudaEventCreate(&d.ev2));
	CK(cudaEventCreateWithFlags(&d.ev_sync, cudaEventBlockingSync | cudaEventDisableTiming));
	int rc;
	if ((rc = dev_alloc(d, &d.d_status, 8))) return rc;
	CK(cudaMallocHost((void **)&d.h_status, 128));
	return ctx_alloc_batch(d, read_cap, max_len, slot_cap);
}

void ctx_free(DeviceCtx &d)
{
	cudaSetDevice(d.dev);
	if (d.owns_index) { cudaFree(d.d_index); cudaFree(d.d_sec); }
	if (d.owns_sa) cudaFree(d.d_sa);
	if (d.owns_rf) cudaFree(d.d_rf);
	if (d.owns_uw) { cudaFree(d.d_uw_text); cudaFree(d.d_fsa); cudaFree(d.d_isa); }
	ctx_free_batch(d);
	cudaFree(d.d_status);
	if (d.h_status) cudaFreeHost(d.h_status);
	if (d.h_tiny_in) cudaFreeHost(d.h_tiny_in);
	if (d.h_arena) cudaFreeHost(d.h_arena);
	if (d.d_tiny_in) cudaFree(d.d_tiny_in);
	d.h_tiny_in = d.h_arena = d.d_tiny_in = nullptr; d.tiny_in_bytes = 0; d.arena_entries = d.arena_reads = 0;
	if (d.ev0) cudaEventDestroy(d.ev0);
	if (d.ev0s) cudaEventDestroy(d.ev0s);
	if (d.ev1) cudaEventDestroy(d.ev1);
	if (d.ev2) cudaEventDestroy(d.ev2);
	if (d.ev_sync) cudaEventDestroy(d.ev_sync);
	if (d.stream) cudaStreamDestroy(d.stream);
}

// The index lives in HBM as re-packed 64-byte blocks (see smem_device.cuh); the packed bwt_t::bwt the
// caller hands over is staged in a temporary device buffer and re-arranged by repack_kernel, once.
int ctx_upload_index(DeviceCtx &d, const smem_index_desc_t *ix, int src_device)
{
	CK(cudaSetDevice(d.dev));
	const u64 n_blocks = (ix->seq_len + 127) / 128;
	const u64 last_words = ((ix->seq_len - (n_blocks - 1) * 128) + 15) / 16;
	if (ix->seq_len == 0 || ix->bwt_size < (n_blocks - 1) * 16 + 8 + last_words) { d.err = "bwt_size is inconsistent with seq_len"; return SMEM_GPU_E_ARG; }
	const size_t bytes = (size_t)ix->bwt_size * 4;
	const size_t out_bytes = (size_t)(n_blocks + 1) * 64;            // one spare block keeps idle lanes in bounds
	if (d.d_index && d.owns_index) { CK(cudaFree(d.d_index)); if (d.d_sec) CK(cudaFree(d.d_sec)); }
	d.d_index = nullptr; d.d_sec = nullptr; d.owns_index = true;
	d.has_index = false;
	CK(cudaMalloc((void **)&d.d_index, out_bytes));
	CK(cudaMemsetAsync((char *)d.d_index + out_bytes - 64, 0, 64, d.stream));
	const u32 *src = nullptr;
	u32 *tmp = nullptr;
	if (src_device == d.dev) src = ix->bwt;
	else {
		CK(cudaMalloc((void **)&tmp, bytes + 64));
		if (src_device < 0) CK(cudaMemcpyAsync(tmp, ix->bwt, bytes, cudaMemcpyHostToDevice, d.stream));
		else CK(cudaMemcpyPeerAsync(tmp, d.dev, ix->bwt, src_device, bytes, d.stream));
		src = tmp;
	}
	repack_kernel<<<(unsigned)((n_blocks + 255) / 256), 256, 0, d.stream>>>(src, n_blocks, ix->seq_len, d.d_index);
	CK(cudaGetLastError());
	CK(stream_wait(d));
	if (tmp) CK(cudaFree(tmp));
	d.index_bytes = (size_t)n_blocks * 64;
	// the sector form for the one-lane-per-read kernel: every checkpoint must fit 32 bits
	bool narrow = d.owner->build_sectors;
	for (int c = 0; c < 4; ++c) if (ix->L2[c + 1] - ix->L2[c] >= (1ull << 32)) narrow = false;
	if (narrow) {
		CK(cudaMalloc((void **)&d.d_sec, out_bytes));
		sectors_from_blocks_kernel<<<(unsigned)((n_blocks + 1 + 255) / 256), 256, 0, d.stream>>>(d.d_index, n_blocks + 1, d.d_sec);
		CK(cudaGetLastError());
		CK(stream_wait(d));
	}
	d.ix.sec = d.d_sec;
	d.ix.blk = d.d_index;
	d.ix.primary = ix->primary;
	for (int i = 0; i < 5; ++i) d.ix.L2[i] = ix->L2[i];
	d.ix.seq_len = ix->seq_len;
	d.has_index = true;
	return 0;
}


// Unique-walk tables: the text, one base per nibble (pack_text_nib_kernel), the full suffix array and the inverse suffix
// array sampled every 2^isa_shift positions, expanded from the samples on the device (fsa_build_kernel) as 33-bit entries:
// 0.5 + 4.57 + 4.57 / 2^isa_shift bytes per text position (3.1 Gbp, shift 2: 3.1 + 28.3 + 7.1 = 38.5 GB).
int ctx_build_text_index(DeviceCtx &d, const uint8_t *pac, long long l_pac, int src_device, int isa_shift)
{
	CK(cudaSetDevice(d.dev));
	if (d.owns_uw) { cudaFree(d.d_uw_text); cudaFree(d.d_fsa); cudaFree(d.d_isa); }
	d.d_uw_text = nullptr; d.d_fsa = d.d_isa = nullptr; d.owns_uw = true; d.uw_text_len = 0; d.uw_bytes = 0;
	const long long n = 2 * l_pac;
	if ((unsigned long long)n >= (1ull << 33)) { d.err = "text too long for 33-bit suffix-array entries (seq_len >= 2^33): unique-walk tables not built"; return SMEM_GPU_E_ARG; }
	const size_t pac_bytes = (size_t)((l_pac + 3) / 4);
	uint8_t *d_pac = nullptr;
	u64 *tw = nullptr;
	u32 *fsa = nullptr, *isa = nullptr;
	auto fail = [&](int rc) { cudaFree(d_pac == pac ? nullptr : d_pac); cudaFree(tw); cudaFree(fsa); cudaFree(isa); return rc; };
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); d.err = b_; return fail(e_ == cudaErrorMemoryAllocation ? SMEM_GPU_E_NOMEM : SMEM_GPU_E_CUDA); } } while (0)
	if (src_device == d.dev) d_pac = const_cast<uint8_t *>(pac);
	else {
		CKT(cudaMalloc((void **)&d_pac, pac_bytes));
		if (src_device < 0) CKT(cudaMemcpyAsync(d_pac, pac, pac_bytes, cudaMemcpyHostToDevice, d.stream));
		else CKT(cudaMemcpyPeerAsync(d_pac, d.dev, pac, src_device, pac_bytes, d.stream));
	}
	const long long n_words = ((n + 63) / 64 + 2) * 8;                   // 32-bit words of eight bases: whole 32-byte sectors, two spare ones past the end
	const size_t fsa_bytes = ((size_t)(n + 1) / 7 + 2) * 32, isa_bytes = (((size_t)n >> isa_shift) / 7 + 2) * 32;
	CKT(cudaMalloc((void **)&tw, (size_t)n_words * 4));
	CKT(cudaMalloc((void **)&fsa, fsa_bytes));
	CKT(cudaMalloc((void **)&isa, isa_bytes));
	CKT(cudaMemsetAsync(fsa, 0, fsa_bytes, d.stream));
	CKT(cudaMemsetAsync(isa, 0, isa_bytes, d.stream));
	pack_text_nib_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, d.stream>>>(d_pac, l_pac, reinterpret_cast<u32 *>(tw), n_words);
	CKT(cudaGetLastError());
	CKT(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));
	d.status_dirty = true;
	const long long n_sa = (long long)d.n_sa;
	const int grid = (int)std::min<long long>((long long)d.sm_count * 8, (n_sa + 63) / 64);
	fsa_build_kernel<<<grid, 128, 0, d.stream>>>(d.ix, d.d_sa, d.sa_shift, n_sa, fsa, isa, isa_shift, d.d_status);
	CKT(cudaGetLastError());
	CKT(cudaMemcpyAsync(d.h_status, d.d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
	CKT(stream_wait(d));
#undef CKT
	if (d_pac != pac) cudaFree(d_pac);
	d_pac = nullptr;
	if (d.h_status[2] != 0) { d.err = "suffix-array walk did not terminate (corrupt index or samples?)"; return fail(SMEM_GPU_E_INTERNAL); }
	d.d_uw_text = tw; d.d_fsa = fsa; d.d_isa = isa; d.uw_text_len = (u64)n; d.uw_isa_shift = isa_shift;
	d.uw_bytes = (size_t)n_words * 4 + fsa_bytes + isa_bytes;
	return 0;
}

// Repeat filter (smem_repeat.cuh) from the 2-bit forward text: every K-mer of T = forward + reverse complement goes
// through an open-addressing table, one hash group per pass so that the table stays small; a K-mer met twice sets its bit.
int ctx_build_repeat_filter(DeviceCtx &d, const uint8_t *pac, long long l_pac, int src_device, int K, int log2_bits, bool auto_bits)
{
	CK(cudaSetDevice(d.dev));
	if (d.d_rf && d.owns_rf) cudaFree(d.d_rf);
	d.d_rf = nullptr; d.owns_rf = true; d.rf_k = 0; d.rf_log2 = 0; d.rf_text_len = 0;
	const long long n = 2 * l_pac;
	const size_t pac_bytes = (size_t)((l_pac + 3) / 4);
	uint8_t *d_pac = nullptr;
	u64 *tw = nullptr, *slots = nullptr;
	u32 *bits = nullptr, *cur = nullptr, *nxt = nullptr;
	auto fail = [&](int rc) { cudaFree(d_pac == pac ? nullptr : d_pac); cudaFree(tw); cudaFree(slots); cudaFree(bits); cudaFree(cur); cudaFree(nxt); return rc; };
#define CKT(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { char b_[512]; snprintf(b_, sizeof b_, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); d.err = b_; return fail(e_ == cudaErrorMemoryAllocation ? SMEM_GPU_E_NOMEM : SMEM_GPU_E_CUDA); } } while (0)
	if (src_device == d.dev) d_pac = const_cast<uint8_t *>(pac);
	else {
		CKT(cudaMalloc((void **)&d_pac, pac_bytes));
		if (src_device < 0) CKT(cudaMemcpyAsync(d_pac, pac, pac_bytes, cudaMemcpyHostToDevice, d.stream));
		else CKT(cudaMemcpyPeerAsync(d_pac, d.dev, pac, src_device, pac_bytes, d.stream));
	}
	const long long n_words = (n + 31) / 32 + 2;
	int log2_slots = 12;
	while (log2_slots < 28 && (1ll << log2_slots) < 2 * n) ++log2_slots;          // <= 2 GB of slots
	const unsigned G = (unsigned)((2 * n + (1ll << log2_slots) - 1) >> log2_slots);  // distinct k-mers per pass <= slots / 2
	CKT(cudaMalloc((void **)&tw, (size_t)n_words * 8));
	CKT(cudaMalloc((void **)&slots, (size_t)8 << log2_slots));
	CKT(cudaMalloc((void **)&bits, (size_t)1 << (log2_bits - 3)));
	CKT(cudaMemsetAsync(bits, 0, (size_t)1 << (log2_bits - 3), d.stream));
	pack_text_kernel<<<(unsigned)((n_words + 255) / 256), 256, 0, d.stream>>>(d_pac, l_pac, tw, n_words);
	CKT(cudaGetLastError());
	for (unsigned g = 0; g < G; ++g) {
		CKT(cudaMemsetAsync(slots, 0xff, (size_t)8 << log2_slots, d.stream));
		rf_insert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(tw, n, K, G, g, slots, log2_slots, bits, log2_bits);
		CKT(cudaGetLastError());
	}
	CKT(stream_wait(d));
	cudaFree(slots); slots = nullptr;
	if (auto_bits) {
		// fold down while the share of set bits stays <= 1/256 (and never below 2^16 bits)
		cur = bits; bits = nullptr;
		unsigned long long *d_pop = (unsigned long long *)tw, h_pop = 0;          // (the text words are no longer needed)
		while (log2_bits > 16) {
			const long long n_dst = 1ll << (log2_bits - 6);
			if (!nxt) CKT(cudaMalloc((void **)&nxt, (size_t)n_dst * 4));
			CKT(cudaMemsetAsync(d_pop, 0, 8, d.stream));
			rf_fold_kernel<<<(unsigned)((n_dst + 255) / 256), 256, 0, d.stream>>>(cur, nxt, n_dst, d_pop);
			CKT(cudaMemcpyAsync(&h_pop, d_pop, 8, cudaMemcpyDeviceToHost, d.stream));
			CKT(stream_wait(d));
			if (h_pop * 256 > (1ull << (log2_bits - 1))) break;                   // too full: keep `cur`
			std::swap(cur, nxt);                                                   // nxt now holds the larger table; its buffer is reused
			--log2_bits;
		}
		cudaFree(nxt); nxt = nullptr;
		// keep only as many bytes as the chosen size needs
		CKT(cudaMalloc((void **)&bits, (size_t)1 << (log2_bits - 3)));
		CKT(cudaMemcpyAsync(bits, cur, (size_t)1 << (log2_bits - 3), cudaMemcpyDeviceToDevice, d.stream));
		CKT(stream_wait(d));
		cudaFree(cur); cur = nullptr;
	}
#undef CKT
	if (d_pac != pac) cudaFree(d_pac);
	cudaFree(tw);
	d.d_rf = bits; d.rf_k = K; d.rf_log2 = log2_bits; d.rf_text_len = (u64)n;
	return 0;
}

// What the caller staged: CSR bytes (fmt 0: smem_gpu_collect / smem1 / trace) or the compact wire format (fmt 1: smem_reads2_t).
struct BatchIn {
	int fmt = 0;
	const uint8_t *seq = nullptr; const int64_t *offs = nullptr; const int32_t *x = nullptr, *mi = nullptr;
	const smem_reads2_t *r2 = nullptr;
};

int ctx_stage_inner(DeviceCtx &d, const BatchIn &in);

// Lanes of one GPU enqueue their H2D copies in lane order (the copy engine is FIFO), so lane 0's reads land
// first and the seed kernels become ready in priority order.
int ctx_stage(DeviceCtx &d, smem_gpu &h, const BatchIn &in)
{
	if (d.prev_lane) {
		std::unique_lock<std::mutex> lk(h.lane_mu);
		h.lane_cv.wait(lk, [&] { return d.prev_lane->stage_issued == h.stage_epoch; });
	}
	const int rc = ctx_stage_inner(d, in);
	if (d.stage_issued != h.stage_epoch) { std::lock_guard<std::mutex> lk(h.lane_mu); d.stage_issued = h.stage_epoch; h.lane_cv.notify_all(); }
	return rc;
}

enum { TINY_FALLBACK = 1000 };      // internal: the latency path could not finish the batch, take the ordinary one

// pinned blob + device twin for the reads of a small batch, pinned arena for its results (sizes follow the handle's limits)
int ensure_tiny(DeviceCtx &d, int max_len)
{
	const size_t reads = (size_t)std::min<int64_t>(d.read_cap, TINY_MAX);
	const size_t in_bytes = (reads + 1) * 8 + reads * 8 + 512 + reads * (size_t)max_len + 128;
	if (in_bytes > d.tiny_in_bytes) {
		if (d.h_tiny_in) CK(cudaFreeHost(d.h_tiny_in));
		if (d.d_tiny_in) CK(cudaFree(d.d_tiny_in));
		d.h_tiny_in = d.d_tiny_in = nullptr; d.tiny_in_bytes = 0;
		CK(cudaMallocHost((void **)&d.h_tiny_in, in_bytes));
		CK(cudaMalloc((void **)&d.d_tiny_in, in_bytes));
		d.tiny_in_bytes = in_bytes;
	}
	const size_t entries = reads * 48 + 1024;
	if (entries > d.arena_entries || reads > d.arena_reads) {
		if (d.h_arena) CK(cudaFreeHost(d.h_arena));
		d.h_arena = nullptr; d.arena_entries = d.arena_reads = 0;
		CK(cudaMallocHost((void **)&d.h_arena, 64 + (reads + 1) * 8 + reads * 4 + 64 + entries * 36 + 64));
		d.arena_entries = entries; d.arena_reads = reads;
	}
	return 0;
}

struct ArenaView { int *status; long long *off; int *ret; Intv *intv; unsigned short *step, *aux; };
static ArenaView arena_view(const DeviceCtx &d, int64_t n)
{
	ArenaView v;
	uint8_t *p = d.h_arena;
	v.status = reinterpret_cast<int *>(p); p += 64;
	v.off = reinterpret_cast<long long *>(p); p += (size_t)(n + 1) * 8;
	v.ret = reinterpret_cast<int *>(p); p += (size_t)n * 4;
	p = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(p) + 31) & ~(uintptr_t)31);
	v.intv = reinterpret_cast<Intv *>(p); p += d.arena_entries * 32;
	v.step = reinterpret_cast<unsigned short *>(p); p += d.arena_entries * 2;
	v.aux = reinterpret_cast<unsigned short *>(p);
	return v;
}

// H2D only: nothing here walks the batch on the CPU -- read lengths are validated by the pack pre-pass on the device
// (status[5], reported by the run).
int ctx_stage_inner(DeviceCtx &d, const BatchIn &in)
{
	CK(cudaSetDevice(d.dev));
	d.n = d.hi - d.lo;
	d.total = 0;
	d.in_fmt = in.fmt;
	d.h2d = 0;
	if (d.n == 0) return 0;
	if (d.n > d.read_cap) { d.err = "batch exceeds the capacity given to smem_gpu_create"; return SMEM_GPU_E_CAPACITY; }
	if (d.tiny) {
		// latency path: offsets, raw-call arguments and bases gathered into one pinned blob, ONE copy
		d.seq_base = in.offs[d.lo];
		const long long span = in.offs[d.hi] - in.offs[d.lo];
		if (span < 0) { d.err = "offs is not monotone"; return SMEM_GPU_E_ARG; }
		const size_t nbytes = (size_t)span, n = (size_t)d.n;
		d.tiny_o_x = (n + 1) * 8; d.tiny_o_mi = d.tiny_o_x + n * 4; d.tiny_o_seq = (d.tiny_o_mi + n * 4 + 255) & ~(size_t)255;
		if (nbytes > d.seq_cap || d.tiny_o_seq + nbytes + 64 > d.tiny_in_bytes) { d.err = "batch exceeds the capacity given to smem_gpu_create (a read longer than max_read_len?)"; return SMEM_GPU_E_CAPACITY; }
		memcpy(d.h_tiny_in, in.offs + d.lo, (n + 1) * 8);
		if (in.x) { memcpy(d.h_tiny_in + d.tiny_o_x, in.x + d.lo, n * 4); memcpy(d.h_tiny_in + d.tiny_o_mi, in.mi + d.lo, n * 4); }
		memcpy(d.h_tiny_in + d.tiny_o_seq, in.seq + d.seq_base, nbytes);
		CK(cudaMemcpyAsync(d.d_tiny_in, d.h_tiny_in, d.tiny_o_seq + nbytes, cudaMemcpyHostToDevice, d.stream));
		d.h2d = (int64_t)(d.tiny_o_seq + nbytes);
		return 0;
	}
	if (in.fmt == 0) {
		d.seq_base = in.offs[d.lo];
		const long long span = in.offs[d.hi] - in.offs[d.lo];
		if (span < 0) { d.err = "offs is not monotone"; return SMEM_GPU_E_ARG; }
		const size_t nbytes = (size_t)span;
		if (nbytes > d.seq_cap) { d.err = "batch exceeds the capacity given to smem_gpu_create (a read longer than max_read_len?)"; return SMEM_GPU_E_CAPACITY; }
		if (nbytes) CK(cudaMemcpyAsync(d.d_seq, in.seq + d.seq_base, nbytes, cudaMemcpyHostToDevice, d.stream));
		CK(cudaMemcpyAsync(d.d_offs, in.offs + d.lo, (size_t)(d.n + 1) * 8, cudaMemcpyHostToDevice, d.stream));
		d.h2d = (int64_t)nbytes + (d.n + 1) * 8;
		if (in.x) {
			CK(cudaMemcpyAsync(d.d_x, in.x + d.lo, (size_t)d.n * 4, cudaMemcpyHostToDevice, d.stream));
			CK(cudaMemcpyAsync(d.d_mi, in.mi + d.lo, (size_t)d.n * 4, cudaMemcpyHostToDevice, d.stream));
			d.h2d += d.n * 8;
		}
	} else {
		const smem_reads2_t &r = *in.r2;
		const size_t nbytes = (size_t)d.n * (size_t)r.stride;
		if (nbytes > d.seq_cap) { d.err = "batch exceeds the capacity given to smem_gpu_create"; return SMEM_GPU_E_CAPACITY; }
		d.seq_base = 0; d.r2_stride = r.stride; d.r2_read_len = r.read_len; d.r2_has_lens = r.lens != nullptr;
		CK(cudaMemcpyAsync(d.d_seq, r.seq2 + (size_t)d.lo * (size_t)r.stride, nbytes, cudaMemcpyHostToDevice, d.stream));
		d.h2d = (int64_t)nbytes;
		if (r.lens) {
			if (!d.d_lens) CK(cudaMalloc((void **)&d.d_lens, ((size_t)d.read_cap + 1) * sizeof(unsigned short)));
			CK(cudaMemcpyAsync(d.d_lens, r.lens + d.lo, (size_t)d.n * 2, cudaMemcpyHostToDevice, d.stream));
			d.h2d += d.n * 2;
		}
		// ambiguous bases of this shard: the list is sorted by read, so the shard's entries are one contiguous range
		const smem_amb_t *a0 = r.amb, *a1 = r.amb + (r.amb ? r.n_amb : 0);
		if (r.amb && r.n_amb > 0) {
			a0 = std::lower_bound(r.amb, r.amb + r.n_amb, (uint32_t)d.lo, [](const smem_amb_t &e, uint32_t v) { return e.read < v; });
			a1 = std::lower_bound(a0, r.amb + r.n_amb, (uint32_t)d.hi, [](const smem_amb_t &e, uint32_t v) { return e.read < v; });
		}
		d.n_amb = (long long)(a1 - a0);
		if (d.n_amb > 0) {
			if ((size_t)d.n_amb > d.amb_cap) {
				if (d.d_amb) CK(cudaFree(d.d_amb));
				d.d_amb = nullptr; d.amb_cap = 0;
				const size_t cap = (size_t)d.n_amb + (size_t)d.n_amb / 2 + 1024;
				CK(cudaMalloc((void **)&d.d_amb, cap * sizeof(AmbEntry)));
				d.amb_cap = cap;
			}
			CK(cudaMemcpyAsync(d.d_amb, a0, (size_t)d.n_amb * sizeof(AmbEntry), cudaMemcpyHostToDevice, d.stream));
			d.h2d += d.n_amb * (int64_t)sizeof(AmbEntry);
		}
	}
	{ std::lock_guard<std::mutex> lk(d.owner->lane_mu); d.stage_issued = d.owner->stage_epoch; d.owner->lane_cv.notify_all(); }
	// Small one-call batches (no kernel turn to hold, see do_collect) leave the copies queued: the kernels follow on the same
	// stream and the caller's buffers stay valid until the call returns, so the host round trip is saved.  Otherwise the reads
	// are on the device before the call asks for the GPU's kernel turn (it must not hold the turn while its copy is in flight).
	if (d.owner->stage_wait) CK(stream_wait(d));
	return 0;
}

// Launch-bounds variants of seed_kernel that are compiled in: the default geometry (9 CTAs of 128 threads per SM, 56
// registers), one step down (8, 64 registers) and a roomy fallback (6, 80 registers).  make ALL_BOUNDS=1 adds the rest for
// occupancy sweeps; "blocks_per_sm" only accepts what was compiled.
#ifdef SMEM_ALL_BOUNDS
static const int kSeedBounds[] = {3, 4, 5, 6, 7, 8, 9, 10, 12};
#else
static const int kSeedBounds[] = {6, 8, 9};
#endif

#ifndef SEED_SPLIT_BLOCKS
#define SEED_SPLIT_BLOCKS 9      // CTAs per SM the default kernel is compiled for (56 registers)
#endif

template <int MODE, bool WIDE>
int launch_seed_w(DeviceCtx &d, const SeedParams &p, int blocks_per_sm, int grid, size_t smem, int lpr)
{
	if (lpr == 1) {          // one lane per read: 64-thread CTAs, as many as shared memory admits (nine at 101 bp); up to 112 registers
		if constexpr (WIDE) { d.err = "one lane per read needs 32-bit occurrence counts"; return SMEM_GPU_E_INTERNAL; }
		else {
			CK(cudaFuncSetAttribute(seed_kernel<MODE, 9, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			seed_kernel<MODE, 9, false, 1><<<grid, SEED_BLOCK1, smem, d.stream>>>(p);
			CK(cudaGetLastError());
			++d.launches;
			return 0;
		}
	}
	if (lpr == 3 || lpr == 4) {          // lane pairs, each lane on the 32-byte sector form of the index; 4: the backward sweep split over the two lanes
		if constexpr (WIDE) { d.err = "the sector form needs 32-bit occurrence counts"; return SMEM_GPU_E_INTERNAL; }
		else {
			if (lpr == 3) {
				CK(cudaFuncSetAttribute(seed_kernel<MODE, 9, false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				seed_kernel<MODE, 9, false, 2, true><<<grid, SEED_BLOCK, smem, d.stream>>>(p);
			} else {
				CK(cudaFuncSetAttribute(seed_kernel<MODE, SEED_SPLIT_BLOCKS, false, 2, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
				seed_kernel<MODE, SEED_SPLIT_BLOCKS, false, 2, true, true><<<grid, SEED_BLOCK, smem, d.stream>>>(p);
			}
			CK(cudaGetLastError());
			++d.launches;
			return 0;
		}
	}
#define LAUNCH(B)                                                                                                         \
	do {                                                                                                                  \
		CK(cudaFuncSetAttribute(seed_kernel<MODE, B, WIDE, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
		seed_kernel<MODE, B, WIDE, 2><<<grid, SEED_BLOCK, smem, d.stream>>>(p);                                             \
	} while (0)
	switch (blocks_per_sm) {
#ifdef SMEM_ALL_BOUNDS
	case 3: LAUNCH(3); break;
	case 4: LAUNCH(4); break;
	case 5: LAUNCH(5); break;
	case 7: LAUNCH(7); break;
	case 10: LAUNCH(10); break;
	case 12: LAUNCH(12); break;
#endif
	case 6: LAUNCH(6); break;
	case 8: LAUNCH(8); break;
	case 9: LAUNCH(9); break;
	default: d.err = "blocks_per_sm variant not compiled in"; return SMEM_GPU_E_ARG;
	}
#undef LAUNCH
	CK(cudaGetLastError());
	++d.launches;
	return 0;
}

template <int MODE>
int launch_seed(DeviceCtx &d, const SeedParams &p, int blocks_per_sm, int grid, size_t smem, bool wide, int lpr)
{
	return wide ? launch_seed_w<MODE, true>(d, p, blocks_per_sm, grid, smem, 2) : launch_seed_w<MODE, false>(d, p, blocks_per_sm, grid, smem, lpr);
}

int launch_seed_mode(DeviceCtx &d, int mode, const SeedParams &p, int bps, int grid, size_t smem, bool wide, int lpr)
{
	return mode == MODE_COLLECT ? launch_seed<MODE_COLLECT>(d, p, bps, grid, smem, wide, lpr)
	     : mode == MODE_SMEM1 ? launch_seed<MODE_SMEM1>(d, p, bps, grid, smem, wide, lpr) : launch_seed<MODE_TRACE>(d, p, bps, grid, smem, wide, lpr);
}

// Several handles on one GPU (one per host thread, smem_gpu_share_index): the seed kernels of ONE call run as a block.
// Without this the lanes of concurrent calls interleave on the GPU, all calls finish together, and their copies
// (H2D before, D2H after) find the GPU idle; with it call B stages its reads while call A's kernels run and starts
// its kernels while call A's results travel back.
struct DevTurn { std::mutex mu; std::condition_variable cv; const smem_gpu *owner = nullptr; int remaining = 0; };
DevTurn g_turn[64];

void turn_acquire(DeviceCtx &d, smem_gpu &h)
{
	if (!h.use_turn) return;
	DevTurn &t = g_turn[d.dev & 63];
	std::unique_lock<std::mutex> lk(t.mu);
	t.cv.wait(lk, [&] { return t.owner == nullptr || t.owner == &h; });
	if (t.owner == nullptr) {
		t.owner = &h; t.remaining = 0;
		for (auto &o : h.devs) t.remaining += o.dev == d.dev && o.hi > o.lo;
	}
	d.holds_turn = true;
	d.turn_epoch = h.epoch;
}

void turn_release(DeviceCtx &d)
{
	if (!d.holds_turn) return;
	d.holds_turn = false;
	DevTurn &t = g_turn[d.dev & 63];
	std::lock_guard<std::mutex> lk(t.mu);
	if (--t.remaining <= 0) { t.owner = nullptr; t.remaining = 0; t.cv.notify_all(); }
}

// Lanes of one GPU run their seed kernels strictly in lane order: lane k waits (on the device) for lane k-1's
// seed kernel of the same epoch.  The host-side hand-shake only makes sure the event was recorded before it is waited on.
void lane_mark_issued(DeviceCtx &d, smem_gpu &h)
{
	std::lock_guard<std::mutex> lk(h.lane_mu);
	d.seed_issued = h.epoch;
	h.lane_cv.notify_all();
}

int ctx_run_inner(DeviceCtx &d, smem_gpu &h, int mode, const smem_seed_opt_t *opt);

int ctx_run(DeviceCtx &d, smem_gpu &h, int mode, const smem_seed_opt_t *opt)
{
	const int rc = ctx_run_inner(d, h, mode, opt);
	turn_release(d);                          // (normally released right after the seed kernel; this covers the error paths)
	if (rc && d.hi > d.lo && d.turn_epoch != h.epoch) { turn_acquire(d, h); turn_release(d); }   // failed before its turn: keep the lane count right
	if (d.seed_issued != h.epoch) {           // early return (no reads / error): never leave the next lane waiting
		cudaSetDevice(d.dev);
		cudaEventRecord(d.ev1, d.stream);
		lane_mark_issued(d, h);
	}
	return rc;
}

// 16-byte result records need every SA coordinate < 2^33 and query positions < 2^14 (smem_intv16_t)
static bool packed_out_ok(const DeviceCtx &d, const smem_gpu &h) { return d.ix.seq_len < (1ull << 33) && h.max_len < (1 << 14); }

int ensure_packed_out(DeviceCtx &d)
{
	if (d.outp_cap < d.out_cap) {
		if (d.d_outp) CK(cudaFree(d.d_outp));
		d.d_outp = nullptr; d.outp_cap = 0;
		CK(cudaMalloc((void **)&d.d_outp, d.out_cap * sizeof(uint4)));
		d.outp_cap = d.out_cap;
	}
	if (!d.d_off32) CK(cudaMalloc((void **)&d.d_off32, ((size_t)d.read_cap + 2) * sizeof(u32)));
	return 0;
}

// 12-byte records: query positions take pos_bits each, what is left of the third word holds the interval size
static int pos_bits_for(int max_len) { int b = 1; while ((1 << b) < max_len) ++b; return b; }
static bool packed12_ok(const DeviceCtx &d, const smem_gpu &h, int fmt = 2) { return d.ix.seq_len < (1ull << 33) && pos_bits_for(h.max_len) <= (fmt == 3 ? 9 : 13); }

int ensure_exc(DeviceCtx &d, size_t want)
{
	static_assert(sizeof(Exc12) == sizeof(smem_x2exc_t) && sizeof(smem_intv12_t) == 12 && sizeof(smem_intv11_t) == 11, "compact record layouts");
	if (d.exc_cap < want) {
		if (d.d_exc) CK(cudaFree(d.d_exc));
		d.d_exc = nullptr; d.exc_cap = 0;
		CK(cudaMalloc((void **)&d.d_exc, want * sizeof(Exc12)));
		d.exc_cap = want;
	}
	return 0;
}

int ctx_run_inner(DeviceCtx &d, smem_gpu &h, int mode, const smem_seed_opt_t *opt)
{
	CK(cudaSetDevice(d.dev));
	d.mode = mode; d.launches = 0; d.overflow = 0; d.seed_ms = d.total_ms = 0; d.total = 0;
	d.seeds_valid = false; d.out_valid = false; d.outp_fmt = 0; d.n_exc = 0;
	if (d.n == 0) return 0;
	if (!d.has_index) { d.err = "no index uploaded"; return SMEM_GPU_E_NOINDEX; }
	const bool packed_out = d.want_packed != 0, packed12 = d.want_packed >= 2;      // (2 = 12-byte records, 3 = 11-byte records: same machinery)
	const int rec_bytes = d.want_packed == 3 ? -11 : 12;       // (11-byte records are laid out as 12 in d_out first, then squeezed)
	auto rec_dst = [&]() -> u32 * { return d.want_packed == 3 ? reinterpret_cast<u32 *>(d.d_out) : reinterpret_cast<u32 *>(d.d_outp); };   // (both may be re-sized below)
	if (packed_out && (mode != MODE_COLLECT || !packed_out_ok(d, h))) { d.err = "16-byte result records need seq_len < 2^33 and max_read_len < 2^14"; return SMEM_GPU_E_ARG; }
	if (packed12 && !packed12_ok(d, h, d.want_packed)) { d.err = "12-byte result records need seq_len < 2^33 and max_read_len <= 8192 (11-byte records: <= 512)"; return SMEM_GPU_E_ARG; }
	d.pos_bits = pos_bits_for(h.max_len);
	const int bps = h.blocks_per_sm;
	// several lanes on this GPU: the persistent seed kernel leaves `spare_sms` SMs empty, so that the finished
	// lane's scan / compaction kernels run next to the following lane's seed kernel (measured: with every SM
	// occupied they wait for it to drain, even when CTA slots are free)
	const int spare = d.lanes_on_dev > 1 ? std::min(h.spare_sms, d.sm_count / 4) : 0;
	// 16-byte packed prev/curr entries need every SA coordinate < 2^36 and read positions < 2^20
	// ... and the 32-bit occurrence counts of the narrow extend need every base to occur fewer than 2^32 times
	bool wide = h.force_wide || d.ix.seq_len >= (1ull << 36) || h.max_len >= (1 << 20);
	for (int c = 0; c < 4; ++c) if (d.ix.L2[c + 1] - d.ix.L2[c] >= (1ull << 32)) wide = true;
	if (h.lanes_per_read != 2 && !wide && !d.ix.sec) { d.err = "lanes_per_read 1 / 3 need the sector form of the index: set the parameter before smem_gpu_upload_index"; return SMEM_GPU_E_ARG; }
	const int lpr = (h.lanes_per_read != 2 && !wide) ? h.lanes_per_read : 2;      // 3 = lane pairs on the sector index; wide indices always take the lane-pair kernel
	const int pairs_per_cta = lpr == 1 ? SEED_BLOCK1 : SEED_BLOCK / 2;            // reads in flight per CTA
	const int max_grid = d.sm_count * std::max(bps, 9);                              // (the one-lane kernel runs up to nine CTAs per SM whatever bps says)
	const int grid = (int)std::min<int64_t>((int64_t)(d.sm_count - spare) * (lpr == 4 ? SEED_SPLIT_BLOCKS : lpr != 2 ? 9 : bps), (d.n + pairs_per_cta - 1) / pairs_per_cta);
	const int scratch_cap = h.max_len + 2;
	const size_t need = (size_t)std::min<int64_t>(max_grid, (d.read_cap + pairs_per_cta - 1) / pairs_per_cta + 1) * pairs_per_cta * 3 * scratch_cap;
	const int q_stride = ((h.max_len + 1) / 2 + 15) / 16 * 16;      // two bases per byte; keeps pair_stride a multiple of 16
	// b_cap is honoured even if that leaves room for fewer CTAs per SM than blocks_per_sm asks for (the
	// hardware then simply runs fewer); it only shrinks when a single CTA would not fit at all.
	const size_t smem_budget = d.smem_per_block_optin;
	const size_t entry = wide ? 32 : 16;
	int b_cap = h.b_cap;
	if ((size_t)pairs_per_cta * (COLD_BYTES + b_cap * entry + q_stride) > smem_budget) {
		const long long fit = ((long long)(smem_budget / pairs_per_cta) - q_stride - COLD_BYTES) / (long long)entry;
		if (fit < 2) { d.err = "read length too large for the shared-memory staging of one CTA"; return SMEM_GPU_E_CAPACITY; }
		b_cap = (int)fit;
	}
	const int pair_stride = (int)(COLD_BYTES + q_stride + b_cap * entry);
	const size_t smem = (size_t)pairs_per_cta * pair_stride;   // (the grid, not shared memory, keeps the spare CTA slot free)
	if (need > d.scratch_entries) {
		if (d.d_scratch) CK(cudaFree(d.d_scratch));
		d.d_scratch = nullptr; d.scratch_entries = 0;
		CK(cudaMalloc((void **)&d.d_scratch, need * sizeof(Intv)));
		d.scratch_entries = need;
	}
	if (h.slot_cap > d.slots_cap_alloc) {
		CK(cudaFree(d.d_slots)); d.d_slots = nullptr;
		CK(cudaMalloc((void **)&d.d_slots, (size_t)d.read_cap * h.slot_cap * sizeof(Intv)));
		d.slots_cap_alloc = h.slot_cap;
	}
	SeedParams p{};
	p.ix = d.ix;
	p.rlen = d.d_rlen;
	p.n = d.n;
	p.list = nullptr;
	p.xs = d.d_x; p.min_intvs = d.d_mi; p.ret = d.d_ret;
	p.slots = d.d_slots; p.slot_cap = h.slot_cap;
	p.counts = d.d_counts; p.overflow_list = d.d_overflow; p.status = d.d_status;
	p.scratch = d.d_scratch; p.scratch_cap = scratch_cap;
	p.b_cap = b_cap; p.q_stride = q_stride; p.pair_stride = pair_stride;
	if (opt) {
		// bwamem.c:456, the path's only FP: mem_opt_t::split_factor is a float, so the product is a float product
		p.split_len_init = (int)((double)((float)opt->min_seed_len * (float)opt->split_factor) + .499);
		p.split_width = opt->split_width; p.start_width = opt->start_width;
	}
	p.hot_min_intv = (u64)h.hot_min_intv;
	p.l2_mode = h.l2_mode;
	const bool use_rf = mode == MODE_COLLECT && h.repeat_filter && d.d_rf && d.rf_text_len == d.ix.seq_len;   // (a filter of another text is ignored)
	p.qflags = nullptr; p.rf_k = d.rf_k; p.count_skips = h.count_skips; p.spec_walk = h.spec_walk;
	const bool use_uw = mode != MODE_SMEM1 && h.unique_walk && d.d_fsa && d.uw_text_len == d.ix.seq_len;
	p.uw_text = use_uw ? reinterpret_cast<const uint4 *>(d.d_uw_text) : nullptr; p.uw_fsa = d.d_fsa; p.uw_isa = d.d_isa; p.uw_isa_shift = d.uw_isa_shift; p.uw_min_left = h.uw_min_left;
	// (the PH_UW_LF steps read the bases in front of the sampled row from the read: the pattern must be longer than the sampling distance)
	p.uw_min_run = use_uw ? std::max(h.uw_min_run, (1 << d.uw_isa_shift) - 1) : 0x7fffffff;
	{
		const size_t bytes_q = (size_t)d.read_cap * q_stride, bytes = bytes_q + (size_t)d.read_cap * (q_stride >> 4) * 4;   // packed reads | window flags
		if (bytes > d.qpack_bytes) {
			if (d.d_qpack) CK(cudaFree(d.d_qpack));
			d.d_qpack = nullptr; d.qpack_bytes = 0;
			CK(cudaMalloc((void **)&d.d_qpack, bytes));
			d.qpack_bytes = bytes;
		}
		p.qpack = d.d_qpack;
		if (use_rf) p.qflags = reinterpret_cast<u32 *>(reinterpret_cast<uint8_t *>(d.d_qpack) + bytes_q);
	}
	if (packed_out) { const int rc0 = ensure_packed_out(d); if (rc0) return rc0; }
	if (packed12) { const int rc0 = ensure_exc(d, d.out_cap / 16 + 1024); if (rc0) return rc0; }

	{
		const auto tw0 = std::chrono::steady_clock::now();
		turn_acquire(d, h);
		d.acc_turn_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tw0).count();
	}
	const bool tiny = d.tiny;
	if (tiny) {
		if (d.status_dirty) CK(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));    // (the latency path's last kernel leaves them zeroed)
		p.xs = reinterpret_cast<const int *>(d.d_tiny_in + d.tiny_o_x); p.min_intvs = reinterpret_cast<const int *>(d.d_tiny_in + d.tiny_o_mi);
	} else {
		CK(cudaEventRecord(d.ev0, d.stream));
		CK(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));
		CK(cudaMemsetAsync(d.d_counts + d.n, 0, sizeof(int), d.stream));
	}
	d.status_dirty = true;
	{   // pack pre-pass: the caller's reads -> one base per nibble at a fixed stride + lengths (+ the repeat filter's window flags)
		const int cpr = q_stride >> 4;
		const long long chunks = (long long)d.n * cpr;
		const unsigned gridp = (unsigned)((chunks + 255) / 256);
		if (tiny)
			pack_bytes_kernel<<<gridp, 256, 0, d.stream>>>(d.d_tiny_in + d.tiny_o_seq - d.seq_base, reinterpret_cast<const long long *>(d.d_tiny_in), d.n, cpr, h.max_len,
			                                               d.d_qpack, d.d_rlen, d.d_status);
		else if (d.in_fmt == 0)
			pack_bytes_kernel<<<gridp, 256, 0, d.stream>>>(d.d_seq - d.seq_base, d.d_offs, d.n, cpr, h.max_len, d.d_qpack, d.d_rlen, d.d_status);
		else {
			unpack2_kernel<<<gridp, 256, 0, d.stream>>>(d.d_seq, d.r2_stride, d.r2_has_lens ? d.d_lens : nullptr, d.r2_read_len, d.n, cpr, h.max_len,
			                                            d.d_qpack, d.d_rlen, d.d_status);
			if (d.n_amb > 0) {
				amb_patch_kernel<<<(unsigned)((d.n_amb + 255) / 256), 256, 0, d.stream>>>(d.d_amb, d.n_amb, d.lo, d.n, q_stride,
				                                                                          reinterpret_cast<u32 *>(d.d_qpack), d.d_rlen, d.d_status);
				++d.launches;
			}
		}
		CK(cudaGetLastError());
		++d.launches;
		if (use_rf) {
			window_flags_kernel<<<gridp, 256, 0, d.stream>>>(d.d_qpack, d.d_rlen, d.n, cpr, d.d_rf, d.rf_k, d.rf_log2, const_cast<u32 *>(p.qflags));
			CK(cudaGetLastError());
			++d.launches;
		}
	}
	if (d.prev_lane && h.chain_lanes) {
		std::unique_lock<std::mutex> lk(h.lane_mu);
		h.lane_cv.wait(lk, [&] { return d.prev_lane->seed_issued == h.epoch; });
		lk.unlock();
		CK(cudaStreamWaitEvent(d.stream, d.prev_lane->ev1, 0));
	}
	int rc = 0;
	if (!tiny) CK(cudaEventRecord(d.ev0s, d.stream));          // after the pack pre-pass: seed_ms is the seed kernel's own duration
	rc = launch_seed_mode(d, mode, p, bps, grid, smem, wide, lpr);
	if (rc) return rc;
	if (!tiny) CK(cudaEventRecord(d.ev1, d.stream));
	lane_mark_issued(d, h);
	if (tiny) {
		const ArenaView av = arena_view(d, d.n);
		tiny_tail_kernel<<<1, TINY_TPB, 0, d.stream>>>(d.d_slots, h.slot_cap, d.d_counts, (int)d.n, d.d_status, mode == MODE_SMEM1 ? d.d_ret : nullptr, d.d_off, d.d_out,
		                                               d.d_step, mode == MODE_TRACE ? d.d_aux : nullptr, (long long)d.out_cap, av.status, av.off,
		                                               mode == MODE_SMEM1 ? av.ret : nullptr, av.intv, av.step, mode == MODE_TRACE ? av.aux : nullptr,
		                                               (long long)d.arena_entries);
		CK(cudaGetLastError());
		++d.launches;
		turn_release(d);
		CK(stream_wait(d));
		d.status_dirty = false;
		if (av.status[5] & 1) { d.err = "a read is longer than max_read_len (or offs is not monotone / a length exceeds the record stride)"; return SMEM_GPU_E_CAPACITY; }
		if (av.status[2] != 0) { d.err = "device guard tripped (extend budget exceeded)"; return SMEM_GPU_E_INTERNAL; }
		if (!av.status[11]) return TINY_FALLBACK;          // slots outgrown or more intervals than the buffers hold: the ordinary path sorts that out
		d.overflow = 0; d.pass2_skipped = av.status[6]; d.uw_walks = av.status[7];
		d.total = (long long)((unsigned long long)(u32)av.status[8] | ((unsigned long long)(u32)av.status[9] << 32));
		d.out_valid = true;
		d.seed_ms = d.total_ms = 0;                         // (no events on this path: they are four driver calls)
		return 0;
	}
	static const bool trace = getenv("SMEM_GPU_TRACE") != nullptr;
	const auto tt0 = std::chrono::steady_clock::now();
	auto tms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tt0).count(); };
	// The common case needs ONE host round trip: counts -> offsets (scan), slots -> dense output (compaction) and the copies
	// of the status words and of the total are all queued behind the seed kernel, then the host waits once.  Only a run with
	// reads that outgrew their slots, or with more intervals than the output buffers hold, takes the second leg below -- the
	// offsets come from the exact counts either way, so the entries already placed stay where they are.
	auto queue_scan = [&]() -> int {
		const long long n1 = d.n + 1;                  // counts[n] = 0, so off[n] = total
		const int nb = (int)((n1 + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK);
		long long *bsum = (long long *)d.d_tmp;
		scan_local_kernel<<<nb, SCAN_TPB, 0, d.stream>>>(d.d_counts, n1, d.d_off, bsum);
		scan_bsum_kernel<<<1, SCAN_TPB, 0, d.stream>>>(bsum, nb);
		scan_add_kernel<<<(unsigned)((n1 + SCAN_TPB - 1) / SCAN_TPB), SCAN_TPB, 0, d.stream>>>(d.d_off, n1, bsum);
		CK(cudaGetLastError());
		d.launches += 3;
		return 0;
	};
	auto queue_compact = [&]() -> int {
		const long long threads = (long long)d.n * 8;
		if (packed12)
			compact_packed12_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, d.stream>>>(d.d_slots, h.slot_cap, d.d_counts, d.d_off, d.n,
			                                                                                 rec_dst(), d.d_off32, (long long)d.out_cap,
			                                                                                 d.pos_bits, d.d_exc, (long long)d.exc_cap, d.d_status + 4, rec_bytes);
		else if (packed_out)
			compact_packed_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, d.stream>>>(d.d_slots, h.slot_cap, d.d_counts, d.d_off, d.n,
			                                                                               d.d_outp, d.d_off32, (long long)d.out_cap);
		else
			compact_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, d.stream>>>(d.d_slots, h.slot_cap, d.d_counts, d.d_off, d.n,
			                                                                        d.d_out, d.d_step, mode == MODE_TRACE ? d.d_aux : nullptr, (long long)d.out_cap);
		CK(cudaGetLastError());
		++d.launches;
		return 0;
	};
	if ((rc = queue_scan())) return rc;
	if ((rc = queue_compact())) return rc;
	auto queue_squeeze = [&]() -> int {          // 11-byte records: the 12-byte layout in d_out -> byte-packed records in d_outp
		if (d.want_packed != 3) return 0;
		squeeze11_kernel<<<d.sm_count * 8, 256, 0, d.stream>>>(reinterpret_cast<const u32 *>(d.d_out), d.d_off + d.n, reinterpret_cast<u32 *>(d.d_outp));
		CK(cudaGetLastError());
		++d.launches;
		return 0;
	};
	if ((rc = queue_squeeze())) return rc;
	publish_status_kernel<<<1, 32, 0, d.stream>>>(d.d_status, d.d_off + d.n, d.h_status);     // (h_status is pinned: mapped under unified addressing)
	CK(cudaGetLastError());
	++d.launches;
	CK(cudaEventRecord(d.ev2, d.stream));
	CK(stream_wait(d));
	const double t_first = tms();
	if (d.h_status[5] & 1) { d.err = "a read is longer than max_read_len (or offs is not monotone / a length exceeds the record stride)"; return SMEM_GPU_E_CAPACITY; }
	if (d.h_status[5] & 2) { d.err = "an ambiguous-base entry names a read outside its shard (is the list sorted by read?)"; return SMEM_GPU_E_ARG; }
	if (d.h_status[2] != 0) { d.err = "device guard tripped (extend budget exceeded)"; return SMEM_GPU_E_INTERNAL; }
	const int n_over = d.h_status[1];
	d.overflow = n_over;
	d.pass2_skipped = d.h_status[6]; d.uw_walks = d.h_status[7];
	memcpy(&d.total, d.h_status + 8, 8);
	const bool regrow = (size_t)d.total > d.out_cap;
	bool exc_grow = packed12 && (size_t)d.h_status[4] > d.exc_cap;       // (counted before any re-compaction)
	const long long exc_first = d.h_status[4];
	d.n_exc = exc_first;
	if (packed_out && (unsigned long long)d.total >= (1ull << 32)) { d.err = "more than 2^32 intervals in one shard: use the 32-byte form"; return SMEM_GPU_E_CAPACITY; }
	if (n_over > 0 || regrow || exc_grow) {
		int big_cap = 0;
		if (n_over > 0) {
			// Reads that outgrew their result slot are seeded again into slots of the largest count measured.
			std::vector<int> list(n_over);
			CK(cudaMemcpyAsync(list.data(), d.d_overflow, (size_t)n_over * 4, cudaMemcpyDeviceToHost, d.stream));
			CK(stream_wait(d));
			std::sort(list.begin(), list.end());       // deterministic re-run order
			if ((size_t)n_over > d.counts_k_cap) {
				if (d.d_counts_k) CK(cudaFree(d.d_counts_k));
				d.d_counts_k = nullptr; d.counts_k_cap = 0;
				CK(cudaMalloc((void **)&d.d_counts_k, (size_t)n_over * 2 * 4));       // counts + this pass' overflow list
				d.counts_k_cap = n_over;
			}
			CK(cudaMemcpyAsync(d.d_overflow, list.data(), (size_t)n_over * 4, cudaMemcpyHostToDevice, d.stream));
			big_cap = d.h_status[3];
			for (int round = 0; round < 3; ++round) {
				const size_t need_big = (size_t)n_over * big_cap;
				if (need_big > d.big_entries) {
					if (d.d_big) CK(cudaFree(d.d_big));
					d.d_big = nullptr; d.big_entries = 0;
					CK(cudaMalloc((void **)&d.d_big, need_big * sizeof(Intv)));
					d.big_entries = need_big;
				}
				CK(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));
				SeedParams p2 = p;
				p2.n = n_over; p2.list = d.d_overflow; p2.slots = d.d_big; p2.slot_cap = big_cap; p2.counts = d.d_counts_k;
				p2.overflow_list = d.d_counts_k + n_over;
				const int grid2 = (int)std::min<int64_t>(std::min<int64_t>(max_grid, (int64_t)d.sm_count * 4), (n_over + pairs_per_cta - 1) / pairs_per_cta);
				rc = launch_seed_mode(d, mode, p2, bps, grid2, smem, wide, lpr);
				if (rc) return rc;
				CK(cudaMemcpyAsync(d.h_status, d.d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
				CK(stream_wait(d));
				if (d.h_status[2] != 0) { d.err = "device guard tripped in the overflow re-run"; return SMEM_GPU_E_INTERNAL; }
				if (d.h_status[1] == 0) break;
				if (round == 2) { d.err = "overflow re-run did not converge"; return SMEM_GPU_E_INTERNAL; }
				big_cap = d.h_status[3];               // exact now: nothing was abandoned with B in global memory
			}
			// (the re-run reproduces the counts the first pass measured: the offsets already computed stay valid)
		}
		if (regrow) {
			CK(cudaFree(d.d_out)); CK(cudaFree(d.d_step)); CK(cudaFree(d.d_aux)); d.d_out = nullptr; d.d_step = nullptr; d.d_aux = nullptr;
			d.out_cap = (size_t)d.total + (size_t)d.total / 8 + 1024;
			CK(cudaMalloc((void **)&d.d_out, d.out_cap * sizeof(Intv)));
			CK(cudaMalloc((void **)&d.d_step, d.out_cap * sizeof(unsigned short)));
			CK(cudaMalloc((void **)&d.d_aux, d.out_cap * sizeof(unsigned short)));
			if (packed_out) { const int rc0 = ensure_packed_out(d); if (rc0) return rc0; }
		}
		if (packed12) {
			// room for the first compaction's entries plus every entry the overflow reads can add (they are placed a second time, whole)
			const size_t need_exc = (size_t)exc_first + (size_t)n_over * big_cap + 1024;
			if (need_exc > d.exc_cap) exc_grow = true;         // (a new buffer: the first compaction runs again to fill it)
			if (regrow || exc_grow) { const int rc0 = ensure_exc(d, std::max(need_exc + need_exc / 8, d.out_cap / 16 + 1024)); if (rc0) return rc0; }
		}
		if (packed12) CK(cudaMemsetAsync(d.d_status + 4, 0, sizeof(int), d.stream));   // the re-compaction (or the overflow reads' part) counts again
		if (regrow || exc_grow) { if ((rc = queue_compact())) return rc; }       // the first compaction stopped at the old capacity
		else if (packed12) CK(cudaMemcpyAsync(d.d_status + 4, d.h_status + 4, sizeof(int), cudaMemcpyHostToDevice, d.stream));   // keep the first pass' entries
		if (n_over > 0) {
			const long long threads = (long long)n_over * big_cap;
			if (packed12)
				compact_list_packed12_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, d.stream>>>(d.d_big, big_cap, d.d_overflow, d.d_counts_k, n_over, d.d_off,
				                                                                                      rec_dst(), d.pos_bits, d.d_exc,
				                                                                                      (long long)d.exc_cap, d.d_status + 4, rec_bytes);
			else if (packed_out)
				compact_list_packed_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, d.stream>>>(d.d_big, big_cap, d.d_overflow, d.d_counts_k, n_over,
				                                                                                    d.d_off, d.d_outp);
			else
				compact_list_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, d.stream>>>(d.d_big, big_cap, d.d_overflow, d.d_counts_k, n_over,
				                                                                             d.d_off, d.d_out, d.d_step, mode == MODE_TRACE ? d.d_aux : nullptr);
			CK(cudaGetLastError());
			++d.launches;
		}
		if ((rc = queue_squeeze())) return rc;
		if (packed12) { publish_status_kernel<<<1, 32, 0, d.stream>>>(d.d_status, d.d_off + d.n, d.h_status); CK(cudaGetLastError()); }
		CK(cudaEventRecord(d.ev2, d.stream));
		turn_release(d);
		CK(stream_wait(d));
		if (packed12) {
			d.n_exc = d.h_status[4];
			if ((size_t)d.n_exc > d.exc_cap) { d.err = "exception list of the 12-byte records outgrew its buffer twice"; return SMEM_GPU_E_INTERNAL; }
		}
	}
	// the scan / compaction kernels are queued ahead of whatever the next call launches: a persistent seed kernel that got
	// there first would hold every SM and they would wait for it to drain
	turn_release(d);
	if (packed_out) d.outp_fmt = d.want_packed; else d.out_valid = true;
	if (trace) fprintf(stderr, "[smem_gpu trace]   lane %d run: first round trip +%.2f, done +%.2f ms after the seed launch (%d overflow reads%s)\n", d.lane, t_first, tms(), n_over, regrow ? ", output buffers regrown" : "");
	CK(cudaEventElapsedTime(&d.seed_ms, d.ev0s, d.ev1));
	CK(cudaEventElapsedTime(&d.total_ms, d.ev0, d.ev2));
	return 0;
}

int ctx_fetch(DeviceCtx &d, smem_intv_t *intv_out, int64_t *read_off, uint16_t *step_out, int32_t *ret, long long base, uint16_t *aux_out = nullptr)
{
	CK(cudaSetDevice(d.dev));
	d.d2h = 0;
	if (d.n == 0) return 0;
	if (!d.out_valid) { d.err = "the resident results are in the 16-byte form: fetch them with smem_gpu_fetch_packed"; return SMEM_GPU_E_ARG; }
	CK(cudaMemcpyAsync(read_off + d.lo, d.d_off, (size_t)d.n * 8, cudaMemcpyDeviceToHost, d.stream));   // read_off[n] is set by the caller
	d.d2h = d.n * 8;
	if (intv_out && d.total) { CK(cudaMemcpyAsync(intv_out + base, d.d_out, (size_t)d.total * sizeof(Intv), cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.total * 32; }
	if (step_out && d.total) { CK(cudaMemcpyAsync(step_out + base, d.d_step, (size_t)d.total * 2, cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.total * 2; }
	if (aux_out && d.total) { CK(cudaMemcpyAsync(aux_out + base, d.d_aux, (size_t)d.total * 2, cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.total * 2; }
	if (ret) { CK(cudaMemcpyAsync(ret + d.lo, d.d_ret, (size_t)d.n * 4, cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.n * 4; }
	CK(stream_wait(d));
	if (base) for (int64_t i = d.lo; i < d.hi; ++i) read_off[i] += base;
	return 0;
}

// 16-byte records + 32-bit offsets; results of a run that compacted into the 32-byte form are converted on the device first
// fmt 2 = 12-byte records: `out` is then a smem_intv12_t array, exc_out (room for exc_room entries from exc_base on) receives the exception list
int ctx_fetch_packed(DeviceCtx &d, smem_gpu &h, void *out, uint32_t *read_off, long long base, int fmt = 1, smem_x2exc_t *exc_out = nullptr,
                     long long exc_base = 0, long long exc_room = 0)
{
	CK(cudaSetDevice(d.dev));
	d.d2h = 0;
	if (d.n == 0) return 0;
	if (d.outp_fmt != fmt) {
		if (!d.out_valid || d.mode != MODE_COLLECT || !packed_out_ok(d, h) || (unsigned long long)d.total >= (1ull << 32) || (fmt >= 2 && !packed12_ok(d, h, fmt))) {
			d.err = "no resident results that fit the compact records (a run that compacted into one compact form keeps only that form)"; return SMEM_GPU_E_ARG;
		}
		const int rc0 = ensure_packed_out(d);
		if (rc0) return rc0;
		const long long threads = std::max<long long>(d.total, d.n + 1);
		if (fmt >= 2) {
			const int rec_bytes = fmt == 3 ? 11 : 12;
			d.pos_bits = pos_bits_for(h.max_len);
			for (int round = 0; round < 2; ++round) {
				const int rc1 = ensure_exc(d, round == 0 ? d.out_cap / 16 + 1024 : (size_t)d.n_exc + 1024);
				if (rc1) return rc1;
				CK(cudaMemsetAsync(d.d_status + 4, 0, sizeof(int), d.stream));
				d.status_dirty = true;
				intv12_from_dense_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, d.stream>>>(d.d_out, d.total, d.d_off, d.n, reinterpret_cast<u32 *>(d.d_outp),
				                                                                                 d.d_off32, d.pos_bits, d.d_exc, (long long)d.exc_cap, d.d_status + 4, rec_bytes);
				CK(cudaGetLastError());
				CK(cudaMemcpyAsync(d.h_status + 4, d.d_status + 4, sizeof(int), cudaMemcpyDeviceToHost, d.stream));
				CK(stream_wait(d));
				d.n_exc = d.h_status[4];
				if ((size_t)d.n_exc <= d.exc_cap) break;
			}
		} else {
			intv16_from_dense_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, d.stream>>>(d.d_out, d.total, d.d_off, d.n, d.d_outp, d.d_off32);
			CK(cudaGetLastError());
		}
		d.outp_fmt = fmt;
	}
	const size_t rec = fmt == 3 ? 11 : fmt == 2 ? 12 : 16;
	CK(cudaMemcpyAsync(read_off + d.lo, d.d_off32, (size_t)d.n * 4, cudaMemcpyDeviceToHost, d.stream));
	d.d2h = d.n * 4;
	if (out && d.total) { CK(cudaMemcpyAsync((uint8_t *)out + (size_t)base * rec, d.d_outp, (size_t)d.total * rec, cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.total * rec; }
	const bool exc_fits = fmt >= 2 && exc_out && d.n_exc <= exc_room;
	if (exc_fits && d.n_exc) { CK(cudaMemcpyAsync(exc_out + exc_base, d.d_exc, (size_t)d.n_exc * sizeof(Exc12), cudaMemcpyDeviceToHost, d.stream)); d.d2h += d.n_exc * (int64_t)sizeof(Exc12); }
	CK(stream_wait(d));
	if (base) {
		for (int64_t i = d.lo; i < d.hi; ++i) read_off[i] += (uint32_t)base;
		if (exc_fits) for (long long k = 0; k < d.n_exc; ++k) exc_out[exc_base + k].index += (uint32_t)base;
	}
	return 0;
}

template <typename F> int for_each_device(smem_gpu *h, F f)
{
	if (h->devs.size() == 1) { h->devs[0].rc = f(h->devs[0]); }
	else {
		std::vector<std::thread> th;
		for (auto &d : h->devs) th.emplace_back([&d, &f]() { d.rc = f(d); });
		for (auto &t : th) t.join();
	}
	for (auto &d : h->devs) if (d.rc) { h->err = "device " + std::to_string(d.dev) + ": " + d.err; return d.rc; }
	return 0;
}

void shard(smem_gpu *h, int64_t n)
{
	const int g = (int)h->devs.size();
	const int64_t per = (n + g - 1) / g;
	for (int k = 0; k < g; ++k) {
		h->devs[k].lo = std::min<int64_t>(n, k * per);
		h->devs[k].hi = std::min<int64_t>(n, (k + 1) * per);
	}
}

int check_batch(smem_gpu *h, const BatchIn &in, int64_t n)
{
	if (!h || n < 0) { if (h) h->err = "bad argument"; return SMEM_GPU_E_ARG; }
	if (in.fmt == 0) { if (n > 0 && (!in.seq || !in.offs)) { h->err = "bad argument"; return SMEM_GPU_E_ARG; } }
	else {
		const smem_reads2_t *r = in.r2;
		if (!r || r->n_reads != n || (n > 0 && !r->seq2) || r->stride < 1 || (!r->lens && (r->read_len < 0 || r->read_len > 4 * r->stride)) ||
		    r->n_amb < 0 || (r->n_amb > 0 && !r->amb) || n >= (1ll << 32)) { h->err = "bad smem_reads2_t"; return SMEM_GPU_E_ARG; }
	}
	if (n > h->max_batch) { h->err = "batch larger than max_batch_reads"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

int64_t sum_h2d(const smem_gpu *h) { int64_t t = 0; for (auto &d : h->devs) t += d.h2d; return t; }
int64_t sum_d2h(const smem_gpu *h) { int64_t t = 0; for (auto &d : h->devs) t += d.d2h; return t; }

int do_stage(smem_gpu *h, int64_t n, const BatchIn &in)
{
	int rc = check_batch(h, in, n);
	if (rc) return rc;
	shard(h, n);
	{ std::lock_guard<std::mutex> lk(h->lane_mu); ++h->stage_epoch; }
	h->stage_wait = true;
	rc = for_each_device(h, [&](DeviceCtx &d) { return ctx_stage(d, *h, in); });
	if (rc) return rc;
	h->staged = n; h->ran = false;
	h->h2d_bytes = sum_h2d(h);
	return 0;
}

int do_run(smem_gpu *h, int mode, const smem_seed_opt_t *opt, int64_t *total_out, int packed_out = 0)
{
	if (h->staged < 0) { h->err = "nothing staged"; return SMEM_GPU_E_ARG; }
	{ std::lock_guard<std::mutex> lk(h->lane_mu); ++h->epoch; }
	h->use_turn = h->staged >= h->turn_min_reads * (int64_t)h->devs.size();
	int rc = for_each_device(h, [&](DeviceCtx &d) { d.want_packed = packed_out; return ctx_run(d, *h, mode, opt); });
	if (rc) return rc;
	int64_t tot = 0;
	for (auto &d : h->devs) tot += d.total;
	if (total_out) *total_out = tot;
	h->ran = true;
	return 0;
}

int do_fetch(smem_gpu *h, smem_intv_t *intv_out, int64_t cap, int64_t *read_off, uint16_t *step_out, int32_t *ret, int64_t *total_out)
{
	if (!h->ran) { h->err = "no results: call a run function first"; return SMEM_GPU_E_ARG; }
	if (!read_off) { h->err = "read_off is required"; return SMEM_GPU_E_ARG; }
	int64_t tot = 0;
	std::vector<long long> base(h->devs.size());
	for (size_t k = 0; k < h->devs.size(); ++k) { base[k] = tot; tot += h->devs[k].total; }
	if (total_out) *total_out = tot;
	const bool fits = tot <= cap && (intv_out || tot == 0);
	read_off[0] = 0;
	int rc = for_each_device(h, [&](DeviceCtx &d) {
		const size_t k = &d - &h->devs[0];
		return ctx_fetch(d, fits ? intv_out : nullptr, read_off, fits ? step_out : nullptr, ret, base[k]);
	});
	if (rc) return rc;
	read_off[h->staged] = tot;
	h->d2h_bytes = sum_d2h(h);
	if (!fits) { h->err = "intv_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

struct Exc12Out { smem_x2exc_t *exc = nullptr; int64_t cap = 0; int64_t *n_out = nullptr; int32_t *pos_bits_out = nullptr; };

int do_fetch_packed(smem_gpu *h, void *out, int64_t cap, uint32_t *read_off, int64_t *total_out, int fmt = 1, const Exc12Out &xo = Exc12Out())
{
	if (!h->ran) { h->err = "no results: call a run function first"; return SMEM_GPU_E_ARG; }
	if (!read_off) { h->err = "read_off is required"; return SMEM_GPU_E_ARG; }
	int64_t tot = 0;
	std::vector<long long> base(h->devs.size());
	for (size_t k = 0; k < h->devs.size(); ++k) { base[k] = tot; tot += h->devs[k].total; }
	if (total_out) *total_out = tot;
	if (tot >= (1ll << 32)) { h->err = "more than 2^32 intervals: use the 32-byte form"; return SMEM_GPU_E_CAPACITY; }
	const bool fits = tot <= cap && (out || tot == 0);
	read_off[0] = 0;
	// (exception lists are placed device after device; a device that still has to convert its results learns its count inside ctx_fetch_packed,
	//  so with several devices the lists are fetched one device at a time)
	int rc = 0;
	int64_t n_exc = 0;
	if (fmt >= 2 && h->devs.size() > 1) {
		for (auto &d : h->devs) {
			rc = ctx_fetch_packed(d, *h, fits ? out : nullptr, read_off, base[&d - &h->devs[0]], fmt, xo.exc, n_exc, std::max<int64_t>(xo.cap - n_exc, 0));
			if (rc) { h->err = "device " + std::to_string(d.dev) + ": " + d.err; return rc; }
			n_exc += d.n_exc;
		}
	} else {
		rc = for_each_device(h, [&](DeviceCtx &d) { return ctx_fetch_packed(d, *h, fits ? out : nullptr, read_off, base[&d - &h->devs[0]], fmt, xo.exc, 0, xo.cap); });
		if (rc) return rc;
		if (fmt >= 2) n_exc = h->devs[0].n_exc;
	}
	read_off[h->staged] = (uint32_t)tot;
	h->d2h_bytes = sum_d2h(h);
	if (xo.n_out) *xo.n_out = n_exc;
	if (xo.pos_bits_out) *xo.pos_bits_out = pos_bits_for(h->max_len);
	if (!fits) { h->err = "intv_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	if (fmt >= 2 && (n_exc > xo.cap || (n_exc > 0 && !xo.exc))) { h->err = "exc_cap too small for the exception list"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

// One-call form: every context (GPU, or pipeline lane of a GPU) runs stage -> run -> fetch for its shard on its own
// host thread, so that one lane's copies overlap another lane's kernels.  A lane needs the interval totals of all
// lanes before it to know where its results go; those finish first anyway (lanes run in order).
// out16 / read_off32 != nullptr selects the 16-byte result records.
struct CollectOut {
	smem_intv_t *intv = nullptr; int64_t *read_off = nullptr; uint16_t *step = nullptr, *aux = nullptr; int32_t *ret = nullptr;
	void *intv16 = nullptr; uint32_t *read_off32 = nullptr;      // compact records: smem_intv16_t (fmt 1) or smem_intv12_t (fmt 2)
	int fmt = 1; Exc12Out xo;
	int64_t cap = 0; int64_t *total = nullptr;
};

// Latency path (see DeviceCtx::tiny): one context, byte reads, 32-byte results, at most TINY_MAX reads.
int tiny_collect(smem_gpu *h, int mode, int64_t n, const BatchIn &in, const smem_seed_opt_t *opt, const CollectOut &out)
{
	DeviceCtx &d = h->devs[0];
	CK(cudaSetDevice(d.dev));
	int rc = ensure_tiny(d, h->max_len);
	if (rc) return rc;
	shard(h, n);
	{ std::lock_guard<std::mutex> lk(h->lane_mu); ++h->epoch; ++h->stage_epoch; }
	h->use_turn = false; h->stage_wait = false;
	d.want_packed = 0;
	d.tiny = true;
	rc = ctx_stage(d, *h, in);
	if (!rc) rc = ctx_run(d, *h, mode, opt);
	d.tiny = false;
	h->staged = n; h->ran = rc == 0;
	if (rc) return rc;
	const ArenaView av = arena_view(d, n);
	h->h2d_bytes = d.h2d;
	if (out.total) *out.total = d.total;
	if (!av.status[10]) {                          // more intervals than the arena holds: fetch them the ordinary way
		out.read_off[0] = 0;
		const bool fits = out.intv && d.total <= out.cap;
		rc = ctx_fetch(d, fits ? out.intv : nullptr, out.read_off, fits ? out.step : nullptr, out.ret, 0, fits ? out.aux : nullptr);
		if (rc) return rc;
		out.read_off[n] = d.total;
		h->d2h_bytes = d.d2h;
		if (!fits && d.total > 0) { h->err = "intv_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
		return 0;
	}
	memcpy(out.read_off, av.off, (size_t)(n + 1) * 8);
	if (out.ret) memcpy(out.ret, av.ret, (size_t)n * 4);
	d.d2h = 64 + (n + 1) * 8 + (out.ret ? n * 4 : 0);
	if (d.total > 0 && (!out.intv || d.total > out.cap)) { h->d2h_bytes = d.d2h; h->err = "intv_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	memcpy(out.intv, av.intv, (size_t)d.total * 32);
	if (out.step) memcpy(out.step, av.step, (size_t)d.total * 2);
	if (out.aux) memcpy(out.aux, av.aux, (size_t)d.total * 2);
	d.d2h += d.total * (32 + 2 + (mode == MODE_TRACE ? 2 : 0));
	h->d2h_bytes = d.d2h;
	return 0;
}

int do_collect(smem_gpu *h, int mode, int64_t n, const BatchIn &in, const smem_seed_opt_t *opt, const CollectOut &out)
{
	const bool packed = out.read_off32 != nullptr;
	if (!h || (!packed && !out.read_off)) { if (h) h->err = "bad argument"; return SMEM_GPU_E_ARG; }
	int rc = check_batch(h, in, n);
	if (rc) return rc;
	if (h->tiny_path && h->devs.size() == 1 && !packed && in.fmt == 0 && n > 0 && n <= TINY_MAX) {
		rc = tiny_collect(h, mode, n, in, opt, out);
		if (rc != TINY_FALLBACK) return rc;
	}
	shard(h, n);
	{ std::lock_guard<std::mutex> lk(h->lane_mu); ++h->epoch; ++h->stage_epoch; }
	const size_t G = h->devs.size();
	h->use_turn = n >= h->turn_min_reads * (int64_t)G;
	h->stage_wait = h->use_turn;
	std::vector<long long> totals(G, 0), excs(G, 0);
	std::vector<char> ran(G, 0);
	bool exc_overflow = false;
	const int want_fmt = packed ? out.fmt : 0;
	std::mutex mu;
	std::condition_variable cv;
	bool overflow = false;
	if (packed) out.read_off32[0] = 0; else out.read_off[0] = 0;
	static const bool trace = getenv("SMEM_GPU_TRACE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	auto ms = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	const bool have_out = packed ? out.intv16 != nullptr : out.intv != nullptr;
	rc = for_each_device(h, [&](DeviceCtx &d) {
		const size_t k = &d - &h->devs[0];
		const double ta = ms();
		d.want_packed = want_fmt;
		int r = ctx_stage(d, *h, in);
		const double tb = ms();
		const double turn0 = d.acc_turn_ms;
		if (!r) r = ctx_run(d, *h, mode, opt);
		else if (d.hi > d.lo) { turn_acquire(d, *h); turn_release(d); }    // keep the turn's lane count right
		const double tc = ms();
		long long base = 0, exc_base = 0;
		{
			std::unique_lock<std::mutex> lk(mu);
			totals[k] = r ? 0 : d.total; excs[k] = r ? 0 : d.n_exc; ran[k] = 1;
			cv.notify_all();
			cv.wait(lk, [&] { for (size_t j = 0; j < k; ++j) if (!ran[j]) return false; return true; });
			for (size_t j = 0; j < k; ++j) { base += totals[j]; exc_base += excs[j]; }
			if (!r && base + d.total > out.cap) overflow = true;
			if (!r && want_fmt >= 2 && (exc_base + d.n_exc > out.xo.cap || (d.n_exc > 0 && !out.xo.exc))) exc_overflow = true;
		}
		if (r) return r;
		const bool fits = have_out && base + d.total <= out.cap;
		const double tc2 = ms();
		if (packed) {
			if (base + d.total >= (1ll << 32)) { d.err = "more than 2^32 intervals: use the 32-byte form"; return (int)SMEM_GPU_E_CAPACITY; }
			r = ctx_fetch_packed(d, *h, fits ? out.intv16 : nullptr, out.read_off32, base, want_fmt, out.xo.exc, exc_base, std::max<int64_t>(out.xo.cap - exc_base, 0));
		} else
			r = ctx_fetch(d, fits ? out.intv : nullptr, out.read_off, fits ? out.step : nullptr, out.ret, base, fits ? out.aux : nullptr);
		const double td = ms();
		d.acc_stage_ms += tb - ta; d.acc_run_ms += (tc - tb) - (d.acc_turn_ms - turn0); d.acc_fetch_ms += td - tc2; d.acc_calls += 1;
		d.acc_h2d += d.h2d; d.acc_d2h += d.d2h;
		if (trace) fprintf(stderr, "[smem_gpu trace] lane %zu: start %.2f staged %.2f ran %.2f (seed %.2f ms, dev %.2f ms) fetched %.2f\n", k, ta, tb, tc,
		                   d.seed_ms, d.total_ms, td);
		return r;
	});
	h->staged = n; h->ran = rc == 0;
	if (rc) return rc;
	long long tot = 0;
	for (auto t : totals) tot += t;
	if (packed) out.read_off32[n] = (uint32_t)tot; else out.read_off[n] = tot;
	if (out.total) *out.total = tot;
	h->h2d_bytes = sum_h2d(h);
	h->d2h_bytes = sum_d2h(h);
	if (want_fmt >= 2) {
		long long ne = 0;
		for (auto e : excs) ne += e;
		if (out.xo.n_out) *out.xo.n_out = ne;
		if (out.xo.pos_bits_out) *out.xo.pos_bits_out = pos_bits_for(h->max_len);
	}
	if (overflow || (tot > 0 && !have_out)) { h->err = "intv_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	if (exc_overflow) { h->err = "exc_cap too small for the exception list"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

} // namespace

// ------------------------------------------------------------------------------------------- C ABI
extern "C" {

int smem_gpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
	return n;
}

// result slots per read of the first pass: 101 bp reads emit 11 intervals on average, 250 bp reads 48; rarer, longer lists take
// the overflow re-run (ctx_run_inner), so this only trades HBM (read_cap * slot_cap * 32 bytes) against how often that runs
static int default_slot_cap(int max_read_len) { return std::min(256, std::max(64, (max_read_len * 3 / 4 + 31) / 32 * 32)); }

int smem_gpu_create(smem_gpu_t **out, int n_devices, const int *device_ids, int64_t max_batch_reads, int max_read_len)
{
	if (!out || n_devices < 1 || n_devices > 64 || max_batch_reads < 1 || max_read_len < 1 || max_read_len > 65535) return SMEM_GPU_E_ARG;
	*out = nullptr;
	if (smem_gpu_device_count() < 1) return SMEM_GPU_E_NODEVICE;
	smem_gpu *h = new (std::nothrow) smem_gpu();
	if (!h) return SMEM_GPU_E_NOMEM;
	h->max_batch = max_batch_reads; h->max_len = max_read_len;
	h->slot_cap = default_slot_cap(max_read_len);
	h->devs.resize(n_devices);
	const int64_t per = (max_batch_reads + n_devices - 1) / n_devices;
	for (int k = 0; k < n_devices; ++k) {
		const int dev = device_ids ? device_ids[k] : k;
		int lane = 0;
		for (int j = 0; j < k; ++j) lane += (device_ids ? device_ids[j] : j) == dev;
		h->devs[k].lane = lane;
		h->devs[k].owner = h;
		for (int j = k - 1; j >= 0; --j) if ((device_ids ? device_ids[j] : j) == dev) { h->devs[k].prev_lane = &h->devs[j]; break; }
		int rc = ctx_init(h->devs[k], dev, lane, per, max_read_len, h->slot_cap);
		if (rc) {
			fprintf(stderr, "[smem_gpu] create failed on device %d: %s\n", h->devs[k].dev, h->devs[k].err.c_str());
			for (auto &d : h->devs) ctx_free(d);
			delete h;
			return rc;
		}
	}
	for (auto &d : h->devs) { d.lanes_on_dev = 0; for (auto &o : h->devs) d.lanes_on_dev += o.dev == d.dev; }
	*out = h;
	return 0;
}

// Re-size the batch buffers of a handle in place (index, samples and accelerator tables stay): the adapter grows its
// handles when a batch or a read arrives that is larger than what they were created for.
int smem_gpu_resize(smem_gpu_t *h, int64_t max_batch_reads, int max_read_len)
{
	if (!h || max_batch_reads < 1 || max_read_len < 1 || max_read_len > 65535) return SMEM_GPU_E_ARG;
	const int64_t per = (max_batch_reads + (int64_t)h->devs.size() - 1) / (int64_t)h->devs.size();
	const int slot_cap = default_slot_cap(max_read_len);
	for (auto &d : h->devs) {
		cudaSetDevice(d.dev);
		cudaStreamSynchronize(d.stream);
		ctx_free_batch(d);
		const int rc = ctx_alloc_batch(d, per, max_read_len, slot_cap);
		if (rc) { h->err = "device " + std::to_string(d.dev) + ": " + d.err; h->staged = -1; h->ran = false; return rc; }
	}
	h->max_batch = max_batch_reads; h->max_len = max_read_len; h->slot_cap = slot_cap;
	h->staged = -1; h->ran = false;
	return 0;
}

int smem_gpu_destroy(smem_gpu_t *h)
{
	if (!h) return SMEM_GPU_E_ARG;
	for (auto &d : h->devs) ctx_free(d);
	delete h;
	return 0;
}

int ctx_upload_sa(DeviceCtx &d, int sa_shift, u64 n_sa, const uint64_t *sa, int src_device)
{
	CK(cudaSetDevice(d.dev));
	if (d.d_sa && d.owns_sa) CK(cudaFree(d.d_sa));
	d.d_sa = nullptr; d.owns_sa = true; d.sa_shift = -1;
	CK(cudaMalloc((void **)&d.d_sa, (size_t)n_sa * 8));
	if (src_device < 0) CK(cudaMemcpyAsync(d.d_sa, sa, (size_t)n_sa * 8, cudaMemcpyHostToDevice, d.stream));
	else if (src_device == d.dev) CK(cudaMemcpyAsync(d.d_sa, sa, (size_t)n_sa * 8, cudaMemcpyDeviceToDevice, d.stream));
	else CK(cudaMemcpyPeerAsync(d.d_sa, d.dev, sa, src_device, (size_t)n_sa * 8, d.stream));
	CK(stream_wait(d));
	d.sa_shift = sa_shift; d.n_sa = n_sa;
	return 0;
}

int ctx_sa(DeviceCtx &d, const uint64_t *k, uint64_t *out)
{
	CK(cudaSetDevice(d.dev));
	const int64_t n = d.hi - d.lo;
	if (n == 0) return 0;
	if (!d.has_index || d.sa_shift < 0) { d.err = "index or suffix-array samples not uploaded"; return SMEM_GPU_E_NOINDEX; }
	if ((size_t)n > d.k_cap) {
		if (d.d_k) CK(cudaFree(d.d_k));
		if (d.d_kout) CK(cudaFree(d.d_kout));
		d.d_k = d.d_kout = nullptr; d.k_cap = 0;
		CK(cudaMalloc((void **)&d.d_k, (size_t)n * 8)); CK(cudaMalloc((void **)&d.d_kout, (size_t)n * 8));
		d.k_cap = (size_t)n;
	}
	CK(cudaMemcpyAsync(d.d_k, k + d.lo, (size_t)n * 8, cudaMemcpyHostToDevice, d.stream));
	CK(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));
	d.status_dirty = true;
	if (d.owner->sa_from_tables && d.d_fsa && d.uw_text_len == d.ix.seq_len)
		sa_from_fsa_kernel<<<(unsigned)((n + 255) / 256), 256, 0, d.stream>>>(d.d_fsa, n, d.d_k, d.d_kout);
	else {
		const int grid = (int)std::min<int64_t>((int64_t)d.sm_count * 8, (n + 63) / 64);
		sa_kernel<<<grid, 128, 0, d.stream>>>(d.ix, d.d_sa, d.sa_shift, n, d.d_k, d.d_kout, d.d_status);
	}
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(out + d.lo, d.d_kout, (size_t)n * 8, cudaMemcpyDeviceToHost, d.stream));
	CK(cudaMemcpyAsync(d.h_status, d.d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
	CK(stream_wait(d));
	if (d.h_status[2] != 0) { d.err = "suffix-array walk did not terminate (corrupt index?)"; return SMEM_GPU_E_INTERNAL; }
	return 0;
}

// counts (int) -> exclusive offsets (int64) with the library's own three-kernel scan; n1 elements
int run_scan(DeviceCtx &d, const int *counts, long long n1, long long *off)
{
	const int nb = (int)((n1 + SCAN_PER_BLOCK - 1) / SCAN_PER_BLOCK);
	if ((size_t)(nb + 2) * sizeof(long long) > d.tmp_bytes) {
		CK(cudaFree(d.d_tmp)); d.d_tmp = nullptr;
		d.tmp_bytes = (size_t)(nb + 2) * sizeof(long long) * 2;
		CK(cudaMalloc(&d.d_tmp, d.tmp_bytes));
	}
	long long *bsum = (long long *)d.d_tmp;
	scan_local_kernel<<<nb, SCAN_TPB, 0, d.stream>>>(counts, n1, off, bsum);
	scan_bsum_kernel<<<1, SCAN_TPB, 0, d.stream>>>(bsum, nb);
	scan_add_kernel<<<(unsigned)((n1 + SCAN_TPB - 1) / SCAN_TPB), SCAN_TPB, 0, d.stream>>>(off, n1, bsum);
	CK(cudaGetLastError());
	d.launches += 3;
	return 0;
}

// intervals of the last run (d_out, d_off) -> seeds in HBM; d.n_seeds
int ctx_seeds_run(DeviceCtx &d, int min_seed_len, u64 max_occ)
{
	CK(cudaSetDevice(d.dev));
	d.n_seeds = 0;
	d.seeds_valid = false;
	if (d.n == 0) { d.seeds_valid = true; return 0; }
	if (d.sa_shift < 0) { d.err = "suffix-array samples not uploaded"; return SMEM_GPU_E_NOINDEX; }
	if (!d.out_valid) { d.err = "the resident results are 16-byte records (smem_gpu_collect_packed): run smem_gpu_run_collect for the seed-level API"; return SMEM_GPU_E_ARG; }
	const long long total = d.total;
	if ((size_t)total + 1 > d.s_cap) {
		if (d.d_scnt) CK(cudaFree(d.d_scnt));
		if (d.d_soff) CK(cudaFree(d.d_soff));
		d.d_scnt = nullptr; d.d_soff = nullptr; d.s_cap = 0;
		const size_t cap = (size_t)total + (size_t)total / 8 + 1024;
		CK(cudaMalloc((void **)&d.d_scnt, cap * sizeof(int))); CK(cudaMalloc((void **)&d.d_soff, cap * sizeof(long long)));
		d.s_cap = cap;
	}
	if ((size_t)d.n + 1 > d.sroff_cap) {
		if (d.d_sroff) CK(cudaFree(d.d_sroff));
		d.d_sroff = nullptr; d.sroff_cap = 0;
		CK(cudaMalloc((void **)&d.d_sroff, ((size_t)d.n + 1) * sizeof(long long)));
		d.sroff_cap = (size_t)d.n + 1;
	}
	seed_count_kernel<<<(unsigned)((total + 1 + 255) / 256), 256, 0, d.stream>>>(d.d_out, total, min_seed_len, max_occ, d.d_scnt);
	CK(cudaGetLastError());
	int rc = run_scan(d, d.d_scnt, total + 1, d.d_soff);
	if (rc) return rc;
	CK(cudaMemcpyAsync(d.h_status + 8, d.d_soff + total, 8, cudaMemcpyDeviceToHost, d.stream));
	CK(stream_wait(d));
	memcpy(&d.n_seeds, d.h_status + 8, 8);
	if ((size_t)d.n_seeds > d.seeds_cap) {
		if (d.d_seeds) CK(cudaFree(d.d_seeds));
		d.d_seeds = nullptr; d.seeds_cap = 0;
		const size_t cap = (size_t)d.n_seeds + (size_t)d.n_seeds / 8 + 1024;
		CK(cudaMalloc((void **)&d.d_seeds, cap * sizeof(Seed)));
		d.seeds_cap = cap;
	}
	CK(cudaMemsetAsync(d.d_status, 0, 8 * sizeof(int), d.stream));
	d.status_dirty = true;
	if (d.n_seeds > 0 && total > 0) {
		if (d.owner->sa_from_tables && d.d_fsa && d.uw_text_len == d.ix.seq_len)
			seed_expand_fsa_kernel<<<(unsigned)((d.n_seeds + 255) / 256), 256, 0, d.stream>>>(d.d_fsa, d.d_out, total, d.d_soff, d.n_seeds, d.d_seeds);
		else {
			const int grid = (int)std::min<int64_t>((int64_t)d.sm_count * 8, (d.n_seeds + 63) / 64);
			seed_expand_kernel<<<grid, 128, 0, d.stream>>>(d.ix, d.d_sa, d.sa_shift, d.d_out, total, d.d_soff, d.n_seeds, d.d_seeds, d.d_status);
		}
		CK(cudaGetLastError());
	}
	CK(cudaMemcpyAsync(d.h_status, d.d_status, 4 * sizeof(int), cudaMemcpyDeviceToHost, d.stream));
	CK(stream_wait(d));
	if (d.h_status[2] != 0) { d.err = "suffix-array walk did not terminate (corrupt index?)"; return SMEM_GPU_E_INTERNAL; }
	d.seeds_valid = true;
	return 0;
}

// resident seeds of the last ctx_seeds_run -> chains in HBM (smem_chain.cuh); d.n_chains, d.n_cseeds
int ctx_chains_run(DeviceCtx &d, const ChainOpt &o)
{
	CK(cudaSetDevice(d.dev));
	d.n_chains = d.n_cseeds = 0; d.chain_ms = 0;
	if (d.n == 0) return 0;
	if (!d.seeds_valid) { d.err = "no resident seeds: call smem_gpu_seeds first"; return SMEM_GPU_E_ARG; }
	const size_t need = (size_t)d.n_seeds + 1;
	if (need > d.cwork_cap) {
		cudaFree(d.d_cwork); cudaFree(d.d_flt); cudaFree(d.d_keep); cudaFree(d.d_chains); cudaFree(d.d_cseeds);
		d.d_cwork = nullptr; d.d_flt = nullptr; d.d_keep = nullptr; d.d_chains = nullptr; d.d_cseeds = nullptr; d.cwork_cap = 0;
		const size_t cap = need + need / 8 + 1024;
		CK(cudaMalloc((void **)&d.d_cwork, cap * 5 * sizeof(int)));
		CK(cudaMalloc((void **)&d.d_flt, cap * sizeof(FltRec)));
		CK(cudaMalloc((void **)&d.d_keep, cap));
		CK(cudaMalloc((void **)&d.d_chains, cap * sizeof(Chain)));       // a chain has at least one seed
		CK(cudaMalloc((void **)&d.d_cseeds, cap * sizeof(Seed)));
		d.cwork_cap = cap;
	}
	if ((size_t)d.n + 1 > d.cread_cap) {
		cudaFree(d.d_nch); cudaFree(d.d_nkept); cudaFree(d.d_coff); cudaFree(d.d_koff);
		d.d_nch = d.d_nkept = nullptr; d.d_coff = d.d_koff = nullptr; d.cread_cap = 0;
		const size_t cap = (size_t)d.n + 1;
		CK(cudaMalloc((void **)&d.d_nch, cap * sizeof(int))); CK(cudaMalloc((void **)&d.d_nkept, cap * sizeof(int)));
		CK(cudaMalloc((void **)&d.d_coff, cap * sizeof(long long))); CK(cudaMalloc((void **)&d.d_koff, cap * sizeof(long long)));
		d.cread_cap = cap;
	}
	ChainWork cw;
	cw.seeds = d.d_seeds; cw.off = d.d_off; cw.soff = d.d_soff; cw.n = d.n;
	{
		const size_t nodes = (size_t)d.n_seeds / 7 + 2 * ((size_t)d.n + 1) + 4;      // read r owns nodes soff / 7 + 2 r .. : at most n_seeds_r / 7 + 2
		if (nodes > d.bt_nodes) {
			cudaFree(d.d_bt); d.d_bt = nullptr; d.bt_nodes = 0;
			const size_t cap = nodes + nodes / 8 + 64;
			CK(cudaMalloc((void **)&d.d_bt, cap * BT_INTS * sizeof(int)));
			d.bt_nodes = cap;
		}
	}
	cw.bt = d.d_bt;
	cw.ord = d.d_cwork; cw.ord2 = cw.ord + d.cwork_cap; cw.c_last = cw.ord2 + d.cwork_cap; cw.c_n = cw.c_last + d.cwork_cap; cw.s_next = cw.c_n + d.cwork_cap;
	cw.flt = d.d_flt; cw.keep = d.d_keep; cw.n_chains = d.d_nch; cw.n_kept = d.d_nkept;
	const unsigned grid = (unsigned)(((d.n + 1) * 32 + CHAIN_TPB - 1) / CHAIN_TPB);
	CK(cudaEventRecord(d.ev0, d.stream));
	chain_build_kernel<<<grid, CHAIN_TPB, 0, d.stream>>>(cw, o);
	CK(cudaGetLastError());
	int rc = run_scan(d, d.d_nch, d.n + 1, d.d_coff);
	if (rc) return rc;
	rc = run_scan(d, d.d_nkept, d.n + 1, d.d_koff);
	if (rc) return rc;
	CK(cudaMemcpyAsync(d.h_status + 8, d.d_coff + d.n, 8, cudaMemcpyDeviceToHost, d.stream));
	CK(cudaMemcpyAsync(d.h_status + 10, d.d_koff + d.n, 8, cudaMemcpyDeviceToHost, d.stream));
	chain_emit_kernel<<<grid, CHAIN_TPB, 0, d.stream>>>(cw, d.d_coff, d.d_koff, 0, d.d_chains, d.d_cseeds);
	CK(cudaGetLastError());
	CK(cudaEventRecord(d.ev1, d.stream));
	CK(stream_wait(d));
	memcpy(&d.n_chains, d.h_status + 8, 8);
	memcpy(&d.n_cseeds, d.h_status + 10, 8);
	CK(cudaEventElapsedTime(&d.chain_ms, d.ev0, d.ev1));
	d.launches += 2;
	return 0;
}

// chain_off[lo .. hi) (+ chain_base), chains (+ seed_base in seed_first) and their seeds to the caller's arrays
int ctx_chains_fetch(DeviceCtx &d, smem_chain_t *chains_out, smem_seed_t *seeds_out, int64_t *chain_off, long long chain_base, long long seed_base)
{
	CK(cudaSetDevice(d.dev));
	if (d.n == 0) return 0;
	if (chain_base) {                        // offsets of this shard in the caller's numbering (d_coff is rebuilt by every run)
		add_base_kernel<<<(unsigned)((d.n + 255) / 256), 256, 0, d.stream>>>(d.d_coff, d.n, chain_base);
		CK(cudaGetLastError());
	}
	CK(cudaMemcpyAsync(chain_off + d.lo, d.d_coff, (size_t)d.n * 8, cudaMemcpyDeviceToHost, d.stream));
	if (chains_out && d.n_chains) CK(cudaMemcpyAsync(chains_out + chain_base, d.d_chains, (size_t)d.n_chains * sizeof(Chain), cudaMemcpyDeviceToHost, d.stream));
	if (seeds_out && d.n_cseeds) CK(cudaMemcpyAsync(seeds_out + seed_base, d.d_cseeds, (size_t)d.n_cseeds * sizeof(Seed), cudaMemcpyDeviceToHost, d.stream));
	CK(stream_wait(d));
	if (chains_out && seed_base) for (long long k = 0; k < d.n_chains; ++k) chains_out[chain_base + k].seed_first += seed_base;
	return 0;
}

int ctx_seeds_fetch(DeviceCtx &d, smem_seed_t *seeds_out, int64_t *seed_off, long long base)
{
	CK(cudaSetDevice(d.dev));
	if (d.n == 0) return 0;
	seed_read_off_kernel<<<(unsigned)((d.n + 1 + 255) / 256), 256, 0, d.stream>>>(d.d_off, d.d_soff, d.n, base, d.d_sroff);
	CK(cudaGetLastError());
	CK(cudaMemcpyAsync(seed_off + d.lo, d.d_sroff, (size_t)d.n * 8, cudaMemcpyDeviceToHost, d.stream));
	if (seeds_out && d.n_seeds) CK(cudaMemcpyAsync(seeds_out + base, d.d_seeds, (size_t)d.n_seeds * sizeof(Seed), cudaMemcpyDeviceToHost, d.stream));
	CK(stream_wait(d));
	return 0;
}

// One copy of the index per physical GPU: contexts that share a device (several pipeline lanes on one GPU,
// see smem_gpu_create) alias the first one's copy.
static int upload_all(smem_gpu_t *h, const smem_index_desc_t *ix, int src_device)
{
	// a repeat filter belongs to the text of the index it was built next to: a new index drops it (build it after the upload)
	for (auto &d : h->devs) {
		if (d.d_rf && d.owns_rf) { cudaSetDevice(d.dev); cudaFree(d.d_rf); }
		d.d_rf = nullptr; d.owns_rf = true; d.rf_k = d.rf_log2 = 0; d.rf_text_len = 0;
		if (d.owns_uw) { cudaSetDevice(d.dev); cudaFree(d.d_uw_text); cudaFree(d.d_fsa); cudaFree(d.d_isa); }
		d.d_uw_text = nullptr; d.d_fsa = d.d_isa = nullptr; d.owns_uw = true; d.uw_text_len = 0; d.uw_bytes = 0;
	}
	int rc = for_each_device(h, [&](DeviceCtx &d) {
		for (auto &o : h->devs) { if (&o == &d) break; if (o.dev == d.dev) return 0; }     // not the first on its GPU
		return ctx_upload_index(d, ix, src_device);
	});
	if (rc) return rc;
	for (auto &d : h->devs)
		for (auto &o : h->devs) {
			if (&o == &d) break;
			if (o.dev == d.dev) {
				if (d.d_index && d.owns_index) { cudaSetDevice(d.dev); cudaFree(d.d_index); cudaFree(d.d_sec); }
				d.d_index = o.d_index; d.d_sec = o.d_sec; d.owns_index = false; d.index_bytes = o.index_bytes; d.ix = o.ix; d.has_index = true;
				break;
			}
		}
	return 0;
}

int smem_gpu_upload_index(smem_gpu_t *h, const smem_index_desc_t *ix)
{
	if (!h || !ix || !ix->bwt || ix->bwt_size < 16) return SMEM_GPU_E_ARG;
	return upload_all(h, ix, -1);
}

int smem_gpu_upload_index_device(smem_gpu_t *h, const smem_index_desc_t *ix, int src_device)
{
	if (!h || !ix || !ix->bwt || ix->bwt_size < 16 || src_device < 0) return SMEM_GPU_E_ARG;
	return upload_all(h, ix, src_device);
}

int smem_gpu_share_index(smem_gpu_t *dst, const smem_gpu_t *src)
{
	if (!dst || !src || dst == src) return SMEM_GPU_E_ARG;
	for (auto &d : dst->devs) {
		const DeviceCtx *o = nullptr;
		for (auto &c : src->devs) if (c.dev == d.dev && c.has_index) { o = &c; break; }
		if (!o) { dst->err = "source handle has no index on device " + std::to_string(d.dev); return SMEM_GPU_E_NOINDEX; }
		cudaSetDevice(d.dev);
		if (d.d_index && d.owns_index) { cudaFree(d.d_index); cudaFree(d.d_sec); }
		if (d.d_sa && d.owns_sa) cudaFree(d.d_sa);
		d.d_index = o->d_index; d.d_sec = o->d_sec; d.owns_index = false; d.index_bytes = o->index_bytes; d.ix = o->ix; d.has_index = true;
		d.d_sa = o->d_sa; d.owns_sa = false; d.sa_shift = o->sa_shift; d.n_sa = o->n_sa;
		if (d.d_rf && d.owns_rf) cudaFree(d.d_rf);
		d.d_rf = o->d_rf; d.owns_rf = false; d.rf_k = o->rf_k; d.rf_log2 = o->rf_log2; d.rf_text_len = o->rf_text_len;
		if (d.owns_uw) { cudaFree(d.d_uw_text); cudaFree(d.d_fsa); cudaFree(d.d_isa); }
		d.d_uw_text = o->d_uw_text; d.d_fsa = o->d_fsa; d.d_isa = o->d_isa; d.owns_uw = false; d.uw_text_len = o->uw_text_len; d.uw_isa_shift = o->uw_isa_shift; d.uw_bytes = o->uw_bytes;
	}
	return 0;
}

int smem_gpu_build_text_index(smem_gpu_t *h, const uint8_t *pac, int64_t l_pac, int src_device)
{
	if (!h || !pac || l_pac < 32) return SMEM_GPU_E_ARG;
	for (auto &d : h->devs) {
		if (!d.has_index || d.sa_shift < 0) { h->err = "upload the index and the suffix-array samples first"; return SMEM_GPU_E_NOINDEX; }
		if (d.ix.seq_len != (u64)(2 * l_pac)) { h->err = "2 * l_pac differs from the seq_len of the uploaded index"; return SMEM_GPU_E_ARG; }
	}
	int rc = for_each_device(h, [&](DeviceCtx &d) {
		for (auto &o : h->devs) { if (&o == &d) break; if (o.dev == d.dev) return 0; }     // not the first on its GPU
		return ctx_build_text_index(d, pac, l_pac, src_device, h->uw_isa_shift);
	});
	if (rc) return rc;
	for (auto &d : h->devs)
		for (auto &o : h->devs) {
			if (&o == &d) break;
			if (o.dev == d.dev) { d.d_uw_text = o.d_uw_text; d.d_fsa = o.d_fsa; d.d_isa = o.d_isa; d.owns_uw = false; d.uw_text_len = o.uw_text_len; d.uw_isa_shift = o.uw_isa_shift; d.uw_bytes = o.uw_bytes; break; }
		}
	return 0;
}

int smem_gpu_get_text_index(smem_gpu_t *h, int which, uint64_t *out, int64_t n_out)
{
	if (!h || !out || which < 0 || which > 1) return SMEM_GPU_E_ARG;
	DeviceCtx &d = h->devs[0];
	if (!d.d_fsa) { h->err = "no text index built"; return SMEM_GPU_E_NOINDEX; }
	const int64_t want = which ? (int64_t)(d.uw_text_len >> d.uw_isa_shift) + 1 : (int64_t)d.uw_text_len + 1;
	if (n_out != want) return SMEM_GPU_E_ARG;
	cudaSetDevice(d.dev);
	u64 *tmp = nullptr;
	if (cudaMalloc((void **)&tmp, (size_t)n_out * 8) != cudaSuccess) return SMEM_GPU_E_NOMEM;
	uw_unpack_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, d.stream>>>(which ? d.d_isa : d.d_fsa, n_out, tmp);
	const bool ok = cudaMemcpyAsync(out, tmp, (size_t)n_out * 8, cudaMemcpyDeviceToHost, d.stream) == cudaSuccess && stream_wait(d) == cudaSuccess;
	cudaFree(tmp);
	return ok ? 0 : SMEM_GPU_E_CUDA;
}

int smem_gpu_build_repeat_filter(smem_gpu_t *h, const uint8_t *pac, int64_t l_pac, int src_device, int kmer_len, int log2_bits)
{
	if (!h || !pac || l_pac < 32) return SMEM_GPU_E_ARG;
	const long long n = 2 * l_pac;
	int lg = 1;
	while ((1ll << lg) < n) ++lg;                                  // ceil(log2 n)
	if (kmer_len <= 0) kmer_len = (lg + 1) / 2 + 5;                // 4^K ~ 1000 x the text length: 22 on a 3.1 Gbp index
	const bool auto_bits = log2_bits <= 0;                        // fill a table of 8 bits per text position, then fold it down (smem_repeat.cuh)
	if (auto_bits) log2_bits = std::min(36, std::max(17, lg + 3));
	if (kmer_len < 8 || kmer_len > 32 || n <= kmer_len || log2_bits < 10 || log2_bits > 40) return SMEM_GPU_E_ARG;
	for (auto &d : h->devs)
		if (d.has_index && d.ix.seq_len != (u64)n) { h->err = "2 * l_pac differs from the seq_len of the uploaded index"; return SMEM_GPU_E_ARG; }
	int rc = for_each_device(h, [&](DeviceCtx &d) {
		for (auto &o : h->devs) { if (&o == &d) break; if (o.dev == d.dev) return 0; }     // not the first on its GPU
		return ctx_build_repeat_filter(d, pac, l_pac, src_device, kmer_len, log2_bits, auto_bits);
	});
	if (rc) return rc;
	for (auto &d : h->devs)
		for (auto &o : h->devs) {
			if (&o == &d) break;
			if (o.dev == d.dev) { d.d_rf = o.d_rf; d.owns_rf = false; d.rf_k = o.rf_k; d.rf_log2 = o.rf_log2; d.rf_text_len = o.rf_text_len; break; }
		}
	return 0;
}

int smem_gpu_get_repeat_filter(smem_gpu_t *h, uint32_t *out, int64_t out_words)
{
	if (!h || !out) return SMEM_GPU_E_ARG;
	DeviceCtx &d = h->devs[0];
	if (!d.d_rf) { h->err = "no repeat filter built"; return SMEM_GPU_E_NOINDEX; }
	if (out_words != (int64_t)1 << (d.rf_log2 - 5)) return SMEM_GPU_E_ARG;
	cudaSetDevice(d.dev);
	return cudaMemcpy(out, d.d_rf, (size_t)out_words * 4, cudaMemcpyDeviceToHost) == cudaSuccess ? 0 : SMEM_GPU_E_CUDA;
}

int smem_gpu_upload_sa(smem_gpu_t *h, int sa_intv, uint64_t n_sa, const uint64_t *sa, int src_device)
{
	if (!h || !sa || n_sa < 1 || sa_intv < 1 || (sa_intv & (sa_intv - 1))) return SMEM_GPU_E_ARG;   // power of two (bwt.c:85-86)
	int shift = 0;
	while ((1 << shift) < sa_intv) ++shift;
	int rc = for_each_device(h, [&](DeviceCtx &d) {
		for (auto &o : h->devs) { if (&o == &d) break; if (o.dev == d.dev) return 0; }
		return ctx_upload_sa(d, shift, n_sa, sa, src_device);
	});
	if (rc) return rc;
	for (auto &d : h->devs)
		for (auto &o : h->devs) {
			if (&o == &d) break;
			if (o.dev == d.dev) {
				if (d.d_sa && d.owns_sa) { cudaSetDevice(d.dev); cudaFree(d.d_sa); }
				d.d_sa = o.d_sa; d.owns_sa = false; d.sa_shift = o.sa_shift; d.n_sa = o.n_sa;
				break;
			}
		}
	return 0;
}

int smem_gpu_sa(smem_gpu_t *h, int64_t n, const uint64_t *k, uint64_t *out)
{
	if (!h || n < 0 || (n > 0 && (!k || !out))) return SMEM_GPU_E_ARG;
	shard(h, n);
	h->staged = -1; h->ran = false;
	return for_each_device(h, [&](DeviceCtx &d) { return ctx_sa(d, k, out); });
}

int smem_gpu_seeds(smem_gpu_t *h, int min_seed_len, int64_t max_occ, smem_seed_t *seeds_out, int64_t seeds_cap, int64_t *seed_off,
                   int64_t *total_out)
{
	if (!h || !seed_off || max_occ < 0) return SMEM_GPU_E_ARG;
	if (!h->ran || h->staged < 0) { h->err = "no resident results: run smem_gpu_collect or smem_gpu_run_collect first"; return SMEM_GPU_E_ARG; }
	static_assert(sizeof(Seed) == sizeof(smem_seed_t), "seed layout");
	int rc = for_each_device(h, [&](DeviceCtx &d) { return ctx_seeds_run(d, min_seed_len, (u64)max_occ); });
	if (rc) return rc;
	long long tot = 0;
	std::vector<long long> base(h->devs.size());
	for (size_t k = 0; k < h->devs.size(); ++k) { base[k] = tot; tot += h->devs[k].n_seeds; }
	if (total_out) *total_out = tot;
	const bool fits = tot <= seeds_cap && (seeds_out || tot == 0);
	rc = for_each_device(h, [&](DeviceCtx &d) { return ctx_seeds_fetch(d, fits ? seeds_out : nullptr, seed_off, base[&d - &h->devs[0]]); });
	if (rc) return rc;
	seed_off[h->staged] = tot;
	if (!fits) { h->err = "seeds_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

int smem_gpu_chains(smem_gpu_t *h, const smem_chain_opt_t *opt, int64_t l_pac, smem_chain_t *chains_out, int64_t chains_cap,
                    smem_seed_t *seeds_out, int64_t seeds_cap, int64_t *chain_off, int64_t *n_chains_out, int64_t *n_seeds_out)
{
	if (!h || !opt || !chain_off || l_pac < 0) return SMEM_GPU_E_ARG;
	if (!h->ran || h->staged < 0) { h->err = "no resident results: run smem_gpu_collect and smem_gpu_seeds first"; return SMEM_GPU_E_ARG; }
	static_assert(sizeof(Chain) == sizeof(smem_chain_t), "chain layout");
	ChainOpt o;
	o.w = opt->w; o.max_chain_gap = opt->max_chain_gap; o.min_seed_len = opt->min_seed_len;
	o.mask_level = opt->mask_level; o.chain_drop_ratio = opt->chain_drop_ratio; o.l_pac = l_pac; o.do_flt = opt->filter != 0;
	int rc = for_each_device(h, [&](DeviceCtx &d) { return ctx_chains_run(d, o); });
	if (rc) return rc;
	long long tc = 0, ts = 0;
	std::vector<long long> cb(h->devs.size()), sb(h->devs.size());
	for (size_t k = 0; k < h->devs.size(); ++k) { cb[k] = tc; sb[k] = ts; tc += h->devs[k].n_chains; ts += h->devs[k].n_cseeds; }
	if (n_chains_out) *n_chains_out = tc;
	if (n_seeds_out) *n_seeds_out = ts;
	const bool fits = tc <= chains_cap && ts <= seeds_cap && (chains_out || tc == 0) && (seeds_out || ts == 0);
	rc = for_each_device(h, [&](DeviceCtx &d) {
		const size_t k = (size_t)(&d - &h->devs[0]);
		return ctx_chains_fetch(d, fits ? chains_out : nullptr, fits ? seeds_out : nullptr, chain_off, cb[k], sb[k]);
	});
	if (rc) return rc;
	chain_off[h->staged] = tc;
	if (!fits) { h->err = "chains_cap / seeds_cap too small for the result"; return SMEM_GPU_E_CAPACITY; }
	return 0;
}

static BatchIn bytes_in(const uint8_t *seq, const int64_t *offs, const int32_t *x = nullptr, const int32_t *mi = nullptr)
{
	BatchIn in; in.fmt = 0; in.seq = seq; in.offs = offs; in.x = x; in.mi = mi; return in;
}
static BatchIn packed_in(const smem_reads2_t *r) { BatchIn in; in.fmt = 1; in.r2 = r; return in; }

int smem_gpu_stage_reads(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs)
{
	return do_stage(h, n_reads, bytes_in(seq, offs));
}

int smem_gpu_stage_reads_packed(smem_gpu_t *h, const smem_reads2_t *reads)
{
	if (!h || !reads) return SMEM_GPU_E_ARG;
	return do_stage(h, reads->n_reads, packed_in(reads));
}

int smem_gpu_run_collect(smem_gpu_t *h, const smem_seed_opt_t *opt, int64_t *total_out)
{
	if (!h || !opt) return SMEM_GPU_E_ARG;
	return do_run(h, MODE_COLLECT, opt, total_out);
}

int smem_gpu_fetch(smem_gpu_t *h, smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off, uint16_t *step_out, int64_t *total_out)
{
	if (!h) return SMEM_GPU_E_ARG;
	return do_fetch(h, intv_out, intv_cap, read_off, step_out, nullptr, total_out);
}

int smem_gpu_fetch_packed(smem_gpu_t *h, smem_intv16_t *intv_out, int64_t intv_cap, uint32_t *read_off, int64_t *total_out)
{
	if (!h) return SMEM_GPU_E_ARG;
	return do_fetch_packed(h, intv_out, intv_cap, read_off, total_out);
}

int smem_gpu_collect(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs, const smem_seed_opt_t *opt,
                     smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off, uint16_t *step_out, int64_t *total_out)
{
	if (!h || !opt || !read_off) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv = intv_out; o.cap = intv_cap; o.read_off = read_off; o.step = step_out; o.total = total_out;
	return do_collect(h, MODE_COLLECT, n_reads, bytes_in(seq, offs), opt, o);
}

int smem_gpu_collect_packed(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv16_t *intv_out, int64_t intv_cap,
                            uint32_t *read_off, int64_t *total_out)
{
	if (!h || !reads || !opt || !read_off) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv16 = intv_out; o.cap = intv_cap; o.read_off32 = read_off; o.total = total_out;
	return do_collect(h, MODE_COLLECT, reads->n_reads, packed_in(reads), opt, o);
}

int smem_gpu_collect_packed12(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv12_t *intv_out, int64_t intv_cap,
                              uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap, int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out)
{
	if (!h || !reads || !opt || !read_off || exc_cap < 0) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv16 = intv_out; o.cap = intv_cap; o.read_off32 = read_off; o.total = total_out; o.fmt = 2;
	o.xo.exc = exc_out; o.xo.cap = exc_cap; o.xo.n_out = n_exc_out; o.xo.pos_bits_out = pos_bits_out;
	return do_collect(h, MODE_COLLECT, reads->n_reads, packed_in(reads), opt, o);
}

int smem_gpu_fetch_packed12(smem_gpu_t *h, smem_intv12_t *intv_out, int64_t intv_cap, uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap,
                            int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out)
{
	if (!h || exc_cap < 0) return SMEM_GPU_E_ARG;
	Exc12Out xo; xo.exc = exc_out; xo.cap = exc_cap; xo.n_out = n_exc_out; xo.pos_bits_out = pos_bits_out;
	return do_fetch_packed(h, intv_out, intv_cap, read_off, total_out, 2, xo);
}

int smem_gpu_collect_packed11(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv11_t *intv_out, int64_t intv_cap,
                              uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap, int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out)
{
	if (!h || !reads || !opt || !read_off || exc_cap < 0) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv16 = intv_out; o.cap = intv_cap; o.read_off32 = read_off; o.total = total_out; o.fmt = 3;
	o.xo.exc = exc_out; o.xo.cap = exc_cap; o.xo.n_out = n_exc_out; o.xo.pos_bits_out = pos_bits_out;
	return do_collect(h, MODE_COLLECT, reads->n_reads, packed_in(reads), opt, o);
}

int smem_gpu_fetch_packed11(smem_gpu_t *h, smem_intv11_t *intv_out, int64_t intv_cap, uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap,
                            int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out)
{
	if (!h || exc_cap < 0) return SMEM_GPU_E_ARG;
	Exc12Out xo; xo.exc = exc_out; xo.cap = exc_cap; xo.n_out = n_exc_out; xo.pos_bits_out = pos_bits_out;
	return do_fetch_packed(h, intv_out, intv_cap, read_off, total_out, 3, xo);
}

int smem_gpu_trace(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs, const smem_seed_opt_t *opt,
                   smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off, uint16_t *tag_out, uint16_t *ret_out, int64_t *total_out)
{
	if (!h || !opt || !read_off || !tag_out || !ret_out) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv = intv_out; o.cap = intv_cap; o.read_off = read_off; o.step = tag_out; o.aux = ret_out; o.total = total_out;
	return do_collect(h, MODE_TRACE, n_reads, bytes_in(seq, offs), opt, o);
}

int smem_gpu_smem1(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs, const int32_t *x, const int32_t *min_intv,
                   smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off, int32_t *ret, int64_t *total_out)
{
	if (!h || !read_off || !ret || (n_reads > 0 && (!x || !min_intv))) return SMEM_GPU_E_ARG;
	CollectOut o; o.intv = intv_out; o.cap = intv_cap; o.read_off = read_off; o.ret = ret; o.total = total_out;
	return do_collect(h, MODE_SMEM1, n_reads, bytes_in(seq, offs, x, min_intv), nullptr, o);
}

// Host-side helper of the compact wire format: CSR bytes -> fixed-stride 2-bit records (+ lengths, + the ambiguous-base
// list, sorted by read because reads are walked in order and the per-thread lists are concatenated in thread order).
int smem_gpu_pack_reads(int64_t n_reads, const uint8_t *seq, const int64_t *offs, int32_t stride, uint8_t *seq2, uint16_t *lens,
                        smem_amb_t *amb, int64_t amb_cap, int64_t *n_amb_out, int n_threads)
{
	if (n_reads < 0 || (n_reads > 0 && (!seq || !offs || !seq2)) || stride < 1 || n_reads >= (1ll << 32)) return SMEM_GPU_E_ARG;
	const int T = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(n_threads < 1 ? 1 : n_threads, 64), (n_reads + 4095) / 4096));
	std::vector<std::vector<smem_amb_t>> found(T);
	std::vector<int> bad(T, 0);
	auto work = [&](int t) {
		const int64_t lo = n_reads * t / T, hi = n_reads * (t + 1) / T;
		for (int64_t r = lo; r < hi; ++r) {
			const int64_t l = offs[r + 1] - offs[r];
			if (l < 0 || l > 4ll * stride || l > 65535) { bad[t] = 1; return; }
			if (lens) lens[r] = (uint16_t)l;
			const uint8_t *q = seq + offs[r];
			uint8_t *o = seq2 + (size_t)r * (size_t)stride;
			int64_t i = 0;
			for (; i + 4 <= l; i += 4) {
				const uint8_t a = q[i], b = q[i + 1], c = q[i + 2], e = q[i + 3];
				if ((a | b | c | e) > 3) for (int k = 0; k < 4; ++k) if (q[i + k] > 3) found[t].push_back(smem_amb_t{(uint32_t)r, (uint16_t)(i + k), 0});
				o[i >> 2] = (uint8_t)(((a & 3) << 6) | ((b & 3) << 4) | ((c & 3) << 2) | (e & 3));
			}
			if (i < l) {
				uint8_t v = 0;
				for (int k = 0; k < 4; ++k) {
					const uint8_t a = i + k < l ? q[i + k] : 0;
					if (a > 3) found[t].push_back(smem_amb_t{(uint32_t)r, (uint16_t)(i + k), 0});
					v = (uint8_t)((v << 2) | (a & 3));
				}
				o[i >> 2] = v; i += 4;
			}
			for (int64_t k = i >> 2; k < stride; ++k) o[k] = 0;
		}
	};
	if (T == 1) work(0);
	else {
		std::vector<std::thread> th;
		for (int t = 0; t < T; ++t) th.emplace_back(work, t);
		for (auto &x : th) x.join();
	}
	for (int t = 0; t < T; ++t) if (bad[t]) return SMEM_GPU_E_ARG;
	int64_t tot = 0;
	for (auto &v : found) tot += (int64_t)v.size();
	if (n_amb_out) *n_amb_out = tot;
	if (tot > amb_cap || (tot > 0 && !amb)) return SMEM_GPU_E_CAPACITY;
	int64_t k = 0;
	for (auto &v : found) { if (!v.empty()) memcpy(amb + k, v.data(), v.size() * sizeof(smem_amb_t)); k += (int64_t)v.size(); }
	return 0;
}

int smem_gpu_host_alloc(void **ptr, size_t bytes)
{
	if (!ptr) return SMEM_GPU_E_ARG;
	return cudaMallocHost(ptr, bytes ? bytes : 1) == cudaSuccess ? 0 : SMEM_GPU_E_NOMEM;
}
int smem_gpu_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? 0 : SMEM_GPU_E_CUDA; }

int smem_gpu_last_timing(const smem_gpu_t *h, smem_gpu_timing_t *t)
{
	if (!h || !t) return SMEM_GPU_E_ARG;
	memset(t, 0, sizeof *t);
	for (auto &d : h->devs) {
		t->seed_kernel_ms = std::max<double>(t->seed_kernel_ms, d.seed_ms);
		t->total_device_ms = std::max<double>(t->total_device_ms, d.total_ms);
		t->kernel_launches += d.launches;
		t->overflow_reads += d.overflow;
	}
	t->h2d_bytes = h->h2d_bytes; t->d2h_bytes = h->d2h_bytes;
	return 0;
}

int smem_gpu_set_param(smem_gpu_t *h, const char *name, int64_t v)
{
	if (!h || !name) return SMEM_GPU_E_ARG;
	if (!strcmp(name, "blocks_per_sm")) {
		for (int k : kSeedBounds) if (k == v) { h->blocks_per_sm = (int)v; return 0; }   // only the variants that were compiled in
		h->err = "blocks_per_sm: that launch-bounds variant is not compiled in (6, 8, 9; make ALL_BOUNDS=1 for the sweep set)";
		return SMEM_GPU_E_ARG;
	}
	if (!strcmp(name, "slot_cap")) { if (v < 1 || v > 4096) return SMEM_GPU_E_ARG; h->slot_cap = (int)v; return 0; }
	if (!strcmp(name, "force_wide")) { h->force_wide = v != 0; return 0; }
	if (!strcmp(name, "spare_sms")) { if (v < 0 || v > 64) return SMEM_GPU_E_ARG; h->spare_sms = (int)v; return 0; }
	if (!strcmp(name, "chain_lanes")) { h->chain_lanes = v != 0; return 0; }
	if (!strcmp(name, "turn_min_reads")) { if (v < 0) return SMEM_GPU_E_ARG; h->turn_min_reads = v; return 0; }
	if (!strcmp(name, "b_cap")) { if (v < 2 || v > 4096) return SMEM_GPU_E_ARG; h->b_cap = (int)v; return 0; }
	if (!strcmp(name, "l2_hot_min_intv")) { if (v < 0) return SMEM_GPU_E_ARG; h->hot_min_intv = v; return 0; }
	if (!strcmp(name, "l2_mode")) { if (v < 0 || v > 2) return SMEM_GPU_E_ARG; h->l2_mode = (int)v; return 0; }
	if (!strcmp(name, "repeat_filter")) { h->repeat_filter = v != 0; return 0; }
	if (!strcmp(name, "count_skips")) { h->count_skips = v != 0; return 0; }
	if (!strcmp(name, "spec_walk")) { h->spec_walk = v != 0; return 0; }
	if (!strcmp(name, "unique_walk")) { h->unique_walk = v != 0; return 0; }
	if (!strcmp(name, "unique_walk_min_run")) { if (v < 1 || v > 65535) return SMEM_GPU_E_ARG; h->uw_min_run = (int)v; return 0; }
	if (!strcmp(name, "unique_walk_min_left")) { if (v < 1 || v > 65535) return SMEM_GPU_E_ARG; h->uw_min_left = (int)v; return 0; }
	if (!strcmp(name, "unique_walk_isa_shift")) { if (v < 0 || v > 6) return SMEM_GPU_E_ARG; h->uw_isa_shift = (int)v; return 0; }
	if (!strcmp(name, "sa_from_tables")) { h->sa_from_tables = v != 0; return 0; }
	if (!strcmp(name, "tiny_path")) { h->tiny_path = v != 0; return 0; }
	if (!strcmp(name, "lanes_per_read")) { if (v < 1 || v > 4) return SMEM_GPU_E_ARG; h->lanes_per_read = (int)v; h->build_sectors = v != 2; return 0; }
	if (!strcmp(name, "build_sectors")) { h->build_sectors = v != 0; return 0; }          // takes effect at the next smem_gpu_upload_index
	if (!strcmp(name, "blocking_sync")) { g_blocking_sync = v != 0; return 0; }          // process-wide
	if (!strcmp(name, "acc_reset")) { for (auto &d : h->devs) { d.acc_stage_ms = d.acc_turn_ms = d.acc_run_ms = d.acc_fetch_ms = 0; d.acc_calls = d.acc_h2d = d.acc_d2h = 0; } return 0; }
	if (!strcmp(name, "probe_variant")) { if (v < 0 || (v > 15 && v != 20)) return SMEM_GPU_E_ARG; h->probe_variant = (int)v; return 0; }
	if (!strcmp(name, "l2_fetch_granularity")) {   // device-wide hint, cudaLimitMaxL2FetchGranularity (32, 64 or 128 bytes)
		if (v != 32 && v != 64 && v != 128) return SMEM_GPU_E_ARG;
		for (auto &d : h->devs) { cudaSetDevice(d.dev); if (cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)v) != cudaSuccess) return SMEM_GPU_E_CUDA; }
		return 0;
	}
	return SMEM_GPU_E_ARG;
}

int64_t smem_gpu_get_param(const smem_gpu_t *h, const char *name)
{
	if (!h || !name) return SMEM_GPU_E_ARG;
	if (!strcmp(name, "block_threads")) return h->block_threads;
	if (!strcmp(name, "blocks_per_sm")) return h->blocks_per_sm;
	if (!strcmp(name, "slot_cap")) return h->slot_cap;
	if (!strcmp(name, "b_cap")) return h->b_cap;
	if (!strcmp(name, "spare_sms")) return h->spare_sms;
	if (!strcmp(name, "chain_lanes")) return h->chain_lanes;
	if (!strcmp(name, "turn_min_reads")) return h->turn_min_reads;
	if (!strcmp(name, "l2_hot_min_intv")) return h->hot_min_intv;
	if (!strcmp(name, "l2_mode")) return h->l2_mode;
	if (!strcmp(name, "repeat_filter")) return h->repeat_filter;
	if (!strcmp(name, "count_skips")) return h->count_skips;
	if (!strcmp(name, "spec_walk")) return h->spec_walk;
	if (!strcmp(name, "unique_walk")) return h->unique_walk;
	if (!strcmp(name, "has_text_index")) return h->devs[0].d_fsa ? 1 : 0;
	if (!strcmp(name, "blocking_sync")) return g_blocking_sync ? 1 : 0;
	if (!strcmp(name, "has_repeat_filter")) return h->devs[0].d_rf ? 1 : 0;
	if (!strcmp(name, "rf_kmer")) return h->devs[0].rf_k;
	if (!strcmp(name, "rf_log2_bits")) return h->devs[0].rf_log2;
	if (!strcmp(name, "pass2_skipped")) { int64_t t = 0; for (auto &d : h->devs) t += d.pass2_skipped; return t; }
	if (!strcmp(name, "unique_walks")) { int64_t t = 0; for (auto &d : h->devs) t += d.uw_walks; return t; }
	if (!strcmp(name, "unique_walk_min_left")) return h->uw_min_left;
	if (!strcmp(name, "unique_walk_min_run")) return h->uw_min_run;
	if (!strcmp(name, "unique_walk_isa_shift")) return h->devs[0].d_fsa ? h->devs[0].uw_isa_shift : h->uw_isa_shift;
	if (!strcmp(name, "chain_kernels_us")) { float m = 0; for (auto &d : h->devs) m = std::max(m, d.chain_ms); return (int64_t)(m * 1000.0f); }
	if (!strcmp(name, "n_chains")) { int64_t t = 0; for (auto &d : h->devs) t += d.n_chains; return t; }
	if (!strcmp(name, "sm_count")) return h->devs[0].sm_count;
	if (!strcmp(name, "l2_fetch_granularity")) { size_t g = 0; cudaSetDevice(h->devs[0].dev); cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity); return (int64_t)g; }
	if (!strcmp(name, "n_devices")) return (int64_t)h->devs.size();
	// wall-time accumulators of the one-call forms (smem_gpu_collect*, trace, smem1), summed over the handle's lanes, in microseconds
	if (!strcmp(name, "acc_stage_us")) { double t = 0; for (auto &d : h->devs) t += d.acc_stage_ms; return (int64_t)(t * 1000); }
	if (!strcmp(name, "acc_turn_us")) { double t = 0; for (auto &d : h->devs) t += d.acc_turn_ms; return (int64_t)(t * 1000); }
	if (!strcmp(name, "acc_run_us")) { double t = 0; for (auto &d : h->devs) t += d.acc_run_ms; return (int64_t)(t * 1000); }
	if (!strcmp(name, "acc_fetch_us")) { double t = 0; for (auto &d : h->devs) t += d.acc_fetch_ms; return (int64_t)(t * 1000); }
	if (!strcmp(name, "acc_calls")) { int64_t t = 0; for (auto &d : h->devs) t += d.acc_calls; return t; }
	if (!strcmp(name, "acc_h2d_bytes")) { int64_t t = 0; for (auto &d : h->devs) t += d.acc_h2d; return t; }
	if (!strcmp(name, "acc_d2h_bytes")) { int64_t t = 0; for (auto &d : h->devs) t += d.acc_d2h; return t; }
	// HBM held by the optional accelerator tables of device 0 (repeat filter + unique-walk tables) and by the index itself
	if (!strcmp(name, "table_bytes")) { const DeviceCtx &d = h->devs[0]; return (int64_t)((d.d_rf ? (size_t)1 << (d.rf_log2 - 3) : 0) + d.uw_bytes); }
	if (!strcmp(name, "uw_table_bytes")) return (int64_t)h->devs[0].uw_bytes;
	if (!strcmp(name, "index_bytes")) return (int64_t)h->devs[0].index_bytes * (h->devs[0].d_sec ? 2 : 1);   // (64-byte blocks + the 32-byte sector form)
	return SMEM_GPU_E_ARG;
}

int smem_gpu_gather_roofline(smem_gpu_t *h, int block_bytes, uint64_t span_bytes, int chains_per_sm, int steps, double *gbps_out)
{
	if (!h || !gbps_out || (block_bytes != 32 && block_bytes != 64 && block_bytes != 128) || chains_per_sm < 256 || steps < 1) return SMEM_GPU_E_ARG;
	DeviceCtx &d = h->devs[0];
	if (!d.has_index) { h->err = "no index uploaded"; return SMEM_GPU_E_NOINDEX; }
	auto body = [&]() -> int {
		CK(cudaSetDevice(d.dev));
		if (span_bytes == 0 || span_bytes > d.index_bytes) span_bytes = d.index_bytes;
		const u64 units = span_bytes / block_bytes;
		const int grid = d.sm_count * (chains_per_sm / 256);
		u64 *sink = (u64 *)d.d_status;
		for (int rep = 0; rep < 2; ++rep) {   // first pass warms the TLB / instruction cache
			CK(cudaEventRecord(d.ev0, d.stream));
#define PROBE(B, V) gather_probe_kernel<B, V><<<grid, 256, 0, d.stream>>>(d.d_index, units, steps, sink)
#define PROBE_B(V) (block_bytes == 32 ? PROBE(32, V) : block_bytes == 64 ? PROBE(64, V) : PROBE(128, V))
#define COOP(L, W) gather_probe_coop_kernel<L, W><<<grid, 256, 0, d.stream>>>(d.d_index, span_bytes / (L * W), steps, sink)
			switch (h->probe_variant) {
			case 10: COOP(2, 32); break;    // 64 B units, 2 lanes x 256-bit
			case 11: COOP(4, 16); break;    // 64 B units, 4 lanes x 128-bit
			case 12: COOP(4, 32); break;    // 128 B units
			case 13: COOP(8, 16); break;    // 128 B units
			case 14: COOP(2, 16); break;    // 32 B units
			case 15: COOP(8, 32); break;    // 256 B units
			case 20: gather_probe_bulk_kernel<<<grid, 256, 0, d.stream>>>(d.d_index, span_bytes / 64, steps, sink); break;   // 64 B per lane by cp.async.bulk + mbarrier
			case 1: PROBE_B(1); break;
			case 2: PROBE_B(2); break;
			case 3: PROBE_B(3); break;
			case 4: PROBE_B(4); break;
			default: PROBE_B(0); break;
			}
#undef PROBE_B
#undef COOP
#undef PROBE
			CK(cudaGetLastError());
			CK(cudaEventRecord(d.ev1, d.stream));
			CK(stream_wait(d));
		}
		float ms = 0;
		CK(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
		double per_thread = block_bytes;
		switch (h->probe_variant) { case 10: case 12: case 15: per_thread = 32; break; case 11: case 13: case 14: per_thread = 16; break; case 20: per_thread = 64; break; default: break; }
		*gbps_out = (double)grid * 256.0 * steps * per_thread / (ms * 1e-3) / 1e9;
		return 0;
	};
	int rc = body();
	if (rc) h->err = d.err;
	return rc;
}

const char *smem_gpu_strerror(int code)
{
	switch (code) {
	case SMEM_GPU_OK: return "ok";
	case SMEM_GPU_E_ARG: return "bad argument";
	case SMEM_GPU_E_CUDA: return "CUDA runtime error";
	case SMEM_GPU_E_NOMEM: return "out of memory";
	case SMEM_GPU_E_NOINDEX: return "no index uploaded";
	case SMEM_GPU_E_CAPACITY: return "capacity exceeded";
	case SMEM_GPU_E_INTERNAL: return "device-side guard tripped";
	case SMEM_GPU_E_NODEVICE: return "no CUDA device";
	default: return "unknown error";
	}
}

const char *smem_gpu_last_error(const smem_gpu_t *h) { return h ? h->err.c_str() : "null handle"; }

} // extern "C"
