/* smem_gpu.h -- C ABI of the B200 SMEM-seeding service.
 *
 * This is the drop-in boundary that replaces the Intel AAL/MPF FPGA service of the reference
 * (HelloALINLB.cpp:254-482 bring-up, bwa.c:289-307 index upload, fastmap.c:320-429 manager thread,
 * bwt.c:444-774 bwt_smem1_batched pack/handshake/unpack).  Plain C: pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success or a negative SMEM_GPU_E_* code; nothing aborts, nothing
 *     falls back to the CPU (the reference's reject->CPU path, bwt.c:651-717, has no equivalent);
 *   - the caller owns all host arrays; the library owns device memory;
 *   - a handle may span 1..8 GPUs of one box: the index is replicated, reads are sharded
 *     contiguously, results land in caller order (no collective on the path);
 *   - a handle is used by one host thread at a time (the link-compatible adapter in
 *     host/bwt_smem1_batched_gpu.c serialises the reference's worker threads onto it).
 *
 * Read batch:   seq  uint8, one base per byte, 0..3 = A,C,G,T, >3 ambiguous (bwamem.c:1403-1406)
 *               offs int64[n+1]; read i = seq[offs[i] .. offs[i+1])
 * Result:       intv smem_intv_t[total] (== bwtintv_t, bwt.h:60-62), grouped by read in batch order,
 *               read_off int64[n+1] (CSR), each read's intervals in the order the reference's
 *               iterator yields them (per step: sorted by start, bwt.c:829 / bwamem.c:281-301).
 */
#ifndef SMEM_GPU_H
#define SMEM_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct smem_gpu smem_gpu_t;

/* == bwtintv_t (bwt.h:60-62): x[0] forward SA start, x[1] reverse-strand SA start, x[2] size,
 * info = query_begin<<32 | query_end (end exclusive). */
typedef struct { uint64_t x[3], info; } smem_intv_t;

/* The fields of bwt_t (bwt.h:46-58) that the path reads.  `bwt` is bwt_t::bwt: 64-byte blocks of
 * 4 x uint64 occ checkpoints + 8 x uint32 words of 16 two-bit symbols (bwtindex.c:128-150). */
typedef struct {
	uint64_t primary;
	uint64_t L2[5];
	uint64_t seq_len;
	uint64_t bwt_size;      /* in uint32 words */
	const uint32_t *bwt;
} smem_index_desc_t;

/* Seeding fields of mem_opt_t (bwamem.h:33-60); defaults mem_opt_init (bwamem.c:45-75):
 * min_seed_len 19, split_factor 1.5, split_width 10, start_width = MEM_F_NO_EXACT ? 2 : 1. */
typedef struct {
	int min_seed_len;
	double split_factor;
	int split_width;
	int start_width;
} smem_seed_opt_t;

enum {
	SMEM_GPU_OK = 0,
	SMEM_GPU_E_ARG = -1,        /* bad argument */
	SMEM_GPU_E_CUDA = -2,       /* CUDA runtime error, see smem_gpu_last_error */
	SMEM_GPU_E_NOMEM = -3,      /* host or device allocation failed */
	SMEM_GPU_E_NOINDEX = -4,    /* no index uploaded */
	SMEM_GPU_E_CAPACITY = -5,   /* batch larger than the handle was created for / output cap too small */
	SMEM_GPU_E_INTERNAL = -6,   /* device-side guard tripped */
	SMEM_GPU_E_NODEVICE = -7    /* no usable CUDA device */
};

/* Replaces HelloALINLBApp::run's service/buffer allocation (HelloALINLB.cpp:309-380).
 * device_ids may be NULL (= 0..n_devices-1). max_batch_reads / max_read_len size the device
 * buffers (per handle, split across devices).  A device id may be listed several times: each
 * occurrence is an independent pipeline lane (own stream, buffers and host thread) on that GPU, so
 * that within one smem_gpu_collect the H2D copy, the kernels and the D2H copy of different
 * shards overlap; lanes of one GPU share a single copy of the index. */
int smem_gpu_create(smem_gpu_t **out, int n_devices, const int *device_ids, int64_t max_batch_reads, int max_read_len);
int smem_gpu_destroy(smem_gpu_t *h);
/* Device memory of a handle grows with both limits: per lane read_cap * (max_read_len + ~2.6 KB) bytes of batch buffers plus
 * min(sm_count * 9, max_batch_reads / 64 + 1) * 64 * 3 * (max_read_len + 2) * 32 bytes of list scratch (for batches that
 * fill the GPU: 0.85 GB at 101 bp, 2.1 GB at 250 bp, 8.4 GB at 1024 bp; max_read_len <= 65535 is accepted, but beyond a few
 * thousand bases the scratch of a full-size batch exhausts HBM and create / resize return SMEM_GPU_E_NOMEM).  smem_gpu_resize re-sizes the batch buffers in place -- index, samples and accelerator
 * tables stay -- so a caller can start small and grow when a longer read or a larger batch arrives. */
int smem_gpu_resize(smem_gpu_t *h, int64_t max_batch_reads, int max_read_len);

/* Replaces the memcpy into SPL_BWT_ref + handshake `2` of bwa_idx_load_bwt (bwa.c:289-301):
 * copies bwt->bwt, primary, L2, seq_len to every GPU of the handle, once. */
int smem_gpu_upload_index(smem_gpu_t *h, const smem_index_desc_t *ix);
/* Same, but ix->bwt is a device pointer on CUDA device `src_device` (index built on the GPU). */
int smem_gpu_upload_index_device(smem_gpu_t *h, const smem_index_desc_t *ix, int src_device);

/* Several handles on the same GPU(s) -- one per host worker thread, the reference's `-t N` pattern -- can share
 * ONE copy of the index (and of the suffix-array samples): dst's contexts alias src's device buffers.  Both handles
 * must list the same devices; src must outlive dst and must not re-upload while dst is in use.  Calls on different
 * handles may run concurrently from different threads, so one thread's copies overlap another thread's kernels. */
int smem_gpu_share_index(smem_gpu_t *dst, const smem_gpu_t *src);

/* Repeat filter of the re-seeding pass -- optional accelerator table of smem_gpu_collect.
 * smem_next2 re-seeds from the middle x of a long, nearly unique SMEM with min_intv >= 2 (bwamem.c:272-278), and the
 * merge keeps an entry of that pass only if its length is >= max >> 1 (bwamem.c:288,297).  Such an entry is a pattern
 * through x that occurs at least twice in the indexed text, and it contains a kmer_len-mer window through x; so when
 * max >> 1 >= kmer_len and every kmer_len-mer window of the read through x occurs at most once in the text, the pass
 * cannot contribute and is skipped (about a third of all bwt_extend calls of a 101 bp read).  The table holds one bit per
 * hash value: set <=> some kmer_len-mer with that hash occurs more than once in T = forward + reverse complement; hash
 * collisions cost a skip, never a wrong one, and results never depend on the table.  Built on every GPU of the handle
 * from the 2-bit forward text: `pac` is bwaidx_t::pac / the .pac file of the reference (bntseq.c:_get_pac: 4 bases per byte,
 * first base in the top bits; the library appends the reverse complement itself, bntseq.c:268-273), l_pac = bntseq_t::l_pac;
 * src_device < 0: host pointer, else the CUDA device holding it.  kmer_len 0 = from the text
 * length (22 on a 3.1 Gbp index; 4^kmer_len ~ 1000 x the text length), log2_bits 0 = the smallest power of two
 * for which at most 1/256 of the bits are set (found by filling a table of 8 bits per text position and folding it down:
 * 32-128 MB and L2-friendly on a repeat-poor text, gigabytes on a repeat-rich one).  smem_gpu_smem1 / smem_gpu_trace return raw bwt_smem1 lists and never skip.
 * "repeat_filter" (smem_gpu_set_param) switches its use off and on; smem_gpu_share_index shares it.  Build it AFTER
 * smem_gpu_upload_index (2 * l_pac must equal the index's seq_len; uploading another index drops the filter). */
int smem_gpu_build_repeat_filter(smem_gpu_t *h, const uint8_t *pac, int64_t l_pac, int src_device, int kmer_len, int log2_bits);
/* Test hook: the bit table of device 0 (2^(log2_bits - 5) uint32 words; get_param "rf_kmer" / "rf_log2_bits"). */
int smem_gpu_get_repeat_filter(smem_gpu_t *h, uint32_t *out, int64_t out_words);

/* Unique-walk tables -- optional accelerator tables of smem_gpu_collect / smem_gpu_trace.
 * Once the forward sweep of bwt_smem1 (bwt.c:790-806) holds an interval of size 1, every further bwt_extend asks whether
 * the text continues like the read at the pattern's ONLY occurrence.  With the text (4 bits per base), the full suffix array and its
 * inverse in HBM (8 bytes per row each: 100 GB at 3.1 Gbp) that walk is one suffix-array lookup, one comparison of the
 * read with the text and one inverse lookup for the reverse-strand row (T = forward + reverse complement, so the reverse
 * complement of a pattern at t sits at seq_len - t - length); x[0] and the size do not change.  The tables are expanded on
 * the device from the suffix-array samples (smem_gpu_upload_sa first) and `pac` / l_pac (as for
 * smem_gpu_build_repeat_filter); results never depend on them ("unique_walk" switches their use off and on). */
int smem_gpu_build_text_index(smem_gpu_t *h, const uint8_t *pac, int64_t l_pac, int src_device);
/* Test hook: which = 0 full suffix array, 1 inverse; seq_len + 1 entries of device 0. */
int smem_gpu_get_text_index(smem_gpu_t *h, int which, uint64_t *out, int64_t n_out);

/* Whole-read seeding == the enumeration loop of mem_insert_seed (bwamem.c:453-460):
 * smem_next2 (bwamem.c:244-305: pass 1, 0.7.8 re-seed of the longest SMEM, ordered merge) to
 * exhaustion for every read.  step_out (nullable) receives, per interval, the index of the
 * smem_next2 call that produced it.  *total_out = number of intervals; if it exceeds intv_cap the
 * call returns SMEM_GPU_E_CAPACITY with read_off and *total_out valid; the CONTENT of intv_out is then unspecified (shards
 * that still fitted may have been copied).  The results stay resident: size the buffer from *total_out and call
 * smem_gpu_fetch -- there is no need to seed again. */
int smem_gpu_collect(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs,
                     const smem_seed_opt_t *opt, smem_intv_t *intv_out, int64_t intv_cap,
                     int64_t *read_off, uint16_t *step_out, int64_t *total_out);

/* One raw bwt_smem1 call per read (bwt.c:776-835) == one DO call of bwt_smem1_batched
 * (bwt.c:444): position x[i], minimum interval size min_intv[i] (clamped to >= 1);
 * ret[i] = end of the longest forward match (itr->start after pass 1). */
int smem_gpu_smem1(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs,
                   const int32_t *x, const int32_t *min_intv, smem_intv_t *intv_out, int64_t intv_cap,
                   int64_t *read_off, int32_t *ret, int64_t *total_out);

/* Seed -> reference position (the step right after the path; bwamem.c:420,474):
 * upload the suffix-array samples of bwt_t (sa_intv, n_sa, sa; bwt.c:79-101, sa[0] = -1) once, then
 * out[i] = bwt_sa(bwt, k[i]) (bwt.c:104-114) for a batch of suffix-array rows.  src_device < 0: `sa` is a host
 * pointer, else a device pointer on that CUDA device. */
int smem_gpu_upload_sa(smem_gpu_t *h, int sa_intv, uint64_t n_sa, const uint64_t *sa, int src_device);
int smem_gpu_sa(smem_gpu_t *h, int64_t n, const uint64_t *k, uint64_t *out);

/* The same walk as smem_gpu_collect, but every bwt_smem1 call keeps its raw list: for each read and each
 * smem_next2 step s the pass-1 list (tag 2s) and, if the iterator re-seeds, the pass-2 list (tag 2s+1), each in
 * ascending start order, with ret_out = bwt_smem1's return value of that call.  This is what the successive DO calls
 * of bwt_smem1_batched return for a batch (bwt.c:719-749), so an adapter can compute all rounds in one launch and
 * serve the reference's per-round calls from the result. */
int smem_gpu_trace(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs, const smem_seed_opt_t *opt,
                   smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off, uint16_t *tag_out, uint16_t *ret_out,
                   int64_t *total_out);

/* == mem_seed_t (bwamem.c:316-319) */
typedef struct { int64_t rbeg; int32_t qbeg, len; } smem_seed_t;

/* Intervals -> seeds, on the results of the last run that are still resident in HBM (smem_gpu_collect or
 * smem_gpu_run_collect): every interval with seed length >= min_seed_len and x[2] <= max_occ (the filter of
 * bwamem.c:412,467) is expanded into its x[2] seeds {rbeg = bwt_sa(x[0] + t), qbeg, len} (bwamem.c:420-424,
 * 474-476), in interval order, then t.  seed_off = int64[n_reads + 1], CSR per read.  Needs smem_gpu_upload_sa.
 * The reference's later per-seed tests (forward/reverse boundary, bwamem.c:426,478) are left to the caller. */
int smem_gpu_seeds(smem_gpu_t *h, int min_seed_len, int64_t max_occ, smem_seed_t *seeds_out, int64_t seeds_cap,
                   int64_t *seed_off, int64_t *total_out);

/* Seeds -> chains (the consumer of the seeds), on the seeds of the last smem_gpu_seeds call that are
 * still resident in HBM.  Per read: the seed loop of mem_insert_seed from the boundary test on (bwamem.c:478-496) -- a seed
 * bridging the forward/reverse boundary at l_pac is skipped, the closest chain at or below its reference position
 * (kb_intervalp on the kbtree of mem_chain_t keyed by pos, bwamem.c:483) absorbs it through test_and_merge
 * (bwamem.c:334-356) or it starts a new chain -- and the chains in ascending pos (__kb_traverse, bwamem.c:608-610); with
 * `filter` != 0 then mem_chain_flt (bwamem.c:629-700): chains ordered by mem_chain_weight (bwamem.c:502-521) with the tie
 * order of ks_introsort, shadowed light chains dropped.  == mem_chain (bwamem.c:593) [+ mem_chain_flt] per read.
 * smem_chain_t mirrors mem_chain_t (bwamem.c:321-325) without the pointer: seeds of chain c are
 * seeds_out[c.seed_first .. c.seed_first + c.n_seeds), in the chain's insertion order; chain_off = int64[n_reads + 1], CSR
 * per read.  Chains whose keys are EQUAL follow a kbtree leaf (first equal key found, new key right after it).
 * On SMEM_GPU_E_CAPACITY chain_off and the two totals are valid. */
typedef struct { int64_t pos; int64_t seed_first; int32_t n_seeds; int32_t weight; } smem_chain_t;
/* chaining fields of mem_opt_t (bwamem.h:33-60); defaults mem_opt_init (bwamem.c:45-75): w 100, max_chain_gap 10000,
 * min_seed_len 19, mask_level 0.5, chain_drop_ratio 0.5. */
typedef struct {
	int w, max_chain_gap, min_seed_len;
	float mask_level, chain_drop_ratio;
	int filter;
} smem_chain_opt_t;
int smem_gpu_chains(smem_gpu_t *h, const smem_chain_opt_t *opt, int64_t l_pac, smem_chain_t *chains_out, int64_t chains_cap,
                    smem_seed_t *seeds_out, int64_t seeds_cap, int64_t *chain_off, int64_t *n_chains_out, int64_t *n_seeds_out);

/* Split form of smem_gpu_collect, so that seeding can be timed with inputs resident in HBM:
 * stage (H2D) -> run (kernels only, may be repeated) -> fetch (D2H). */
int smem_gpu_stage_reads(smem_gpu_t *h, int64_t n_reads, const uint8_t *seq, const int64_t *offs);
int smem_gpu_run_collect(smem_gpu_t *h, const smem_seed_opt_t *opt, int64_t *total_out);
int smem_gpu_fetch(smem_gpu_t *h, smem_intv_t *intv_out, int64_t intv_cap, int64_t *read_off,
                   uint16_t *step_out, int64_t *total_out);

/* ---- Compact wire formats (what replaces the reference's 256-byte input record, bwt.c:574-596,
 * and its 64 + 32 n byte output record, afu_core.v:1931-1960 / bwt.c:719-749).  The interval lists are the same, bit for
 * bit; only their representation on the PCIe link and in the caller's buffers shrinks: 26 instead of 109 bytes per
 * 101 bp read going in, 16 instead of 32 bytes per interval and 4 instead of 8 bytes per read coming out.
 *
 * Reads: fixed-stride records of two bits per base, four bases per byte, FIRST base in the TOP bits (the reference's .pac
 * convention, bntseq.c:_set_pac); record i starts at seq2 + i * stride.  Every read has read_len bases unless lens != NULL.
 * A two-bit code cannot say "ambiguous" (bwamem.c:1403-1406 maps N to 4), so those bases travel as an exception list,
 * sorted by read; the two bits stored at such a position are ignored. */
typedef struct { uint32_t read; uint16_t pos; uint16_t reserved; } smem_amb_t;
typedef struct {
	int64_t n_reads;
	const uint8_t *seq2;
	int32_t stride;          /* bytes per record, >= ceil(longest read / 4) */
	int32_t read_len;        /* length of every read when lens == NULL */
	const uint16_t *lens;    /* optional: per-read lengths */
	const smem_amb_t *amb;   /* optional: ambiguous bases, ascending by read */
	int64_t n_amb;
} smem_reads2_t;

/* Result record: bits 0-32 x[0], 33-65 x[1], 66-98 x[2], 99-112 query begin, 113-126 query end (w[0] = low word).
 * Valid for indices with seq_len < 2^33 (8.5 Gbp of forward + reverse text: any genome up to 4.2 Gbp) and reads shorter
 * than 2^14 bases; the *_packed entry points return SMEM_GPU_E_ARG otherwise (use the 32-byte form). */
typedef struct { uint64_t w[2]; } smem_intv16_t;
static inline void smem_intv16_unpack(const smem_intv16_t *p, smem_intv_t *o)
{
	const uint64_t m33 = ((uint64_t)1 << 33) - 1;
	o->x[0] = p->w[0] & m33;
	o->x[1] = (p->w[0] >> 33) | ((p->w[1] & 3) << 31);
	o->x[2] = (p->w[1] >> 2) & m33;
	o->info = (((p->w[1] >> 35) & 0x3fff) << 32) | ((p->w[1] >> 49) & 0x3fff);
}

/* smem_gpu_collect with both formats: read_off = uint32[n_reads + 1] (CSR), intv_out in the same order as
 * smem_gpu_collect's.  On SMEM_GPU_E_CAPACITY read_off and *total_out are valid. */
int smem_gpu_collect_packed(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv16_t *intv_out,
                            int64_t intv_cap, uint32_t *read_off, int64_t *total_out);
/* Split form: compact reads in (then smem_gpu_run_collect, smem_gpu_seeds, smem_gpu_chains as usual) / 16-byte records out
 * (after any smem_gpu_run_collect or smem_gpu_collect*). */
int smem_gpu_stage_reads_packed(smem_gpu_t *h, const smem_reads2_t *reads);
int smem_gpu_fetch_packed(smem_gpu_t *h, smem_intv16_t *intv_out, int64_t intv_cap, uint32_t *read_off, int64_t *total_out);
/* 12-byte result record, for links where bytes are what limits (eight GPUs behind shared PCIe uplinks):
 * w[0] = low 32 bits of x[0], w[1] = low 32 bits of x[1], w[2] = bit 32 of x[0] | bit 32 of x[1] << 1 | query begin << 2 |
 * (query end - 1) << (2 + P) | field << (2 + 2 P), with P = *pos_bits_out = ceil(log2(max_read_len)) of the handle and
 * field = x[2] - 1 while that is below the all-ones value of the field's 30 - 2 P bits (16 bits = sizes up to 65535 for reads
 * of up to 128 bases).  An all-ones field says "look it up": such intervals are listed, in no particular order, in the
 * exception list {index of the record, x[2]}; decoding = unpack every record, then x[2] = exc.x2 at exc.index.  The same
 * limits as the 16-byte record, plus max_read_len <= 8192.  On SMEM_GPU_E_CAPACITY read_off, *total_out and *n_exc_out are
 * valid (size both buffers and call smem_gpu_fetch_packed12). */
typedef struct { uint32_t w[3]; } smem_intv12_t;
typedef struct { uint32_t index; uint32_t x2_lo; uint32_t x2_hi; } smem_x2exc_t;
/* returns 1 if x[2] has to come from the exception list (o->x[2] is then 0) */
static inline int smem_intv12_unpack(const smem_intv12_t *p, int pos_bits, smem_intv_t *o)
{
	const uint32_t w2 = p->w[2], pm = ((uint32_t)1 << pos_bits) - 1, esc = ((uint32_t)1 << (30 - 2 * pos_bits)) - 1;
	const uint32_t f = w2 >> (2 + 2 * pos_bits);
	o->x[0] = (uint64_t)p->w[0] | ((uint64_t)(w2 & 1) << 32);
	o->x[1] = (uint64_t)p->w[1] | ((uint64_t)((w2 >> 1) & 1) << 32);
	o->x[2] = f == esc ? 0 : (uint64_t)f + 1;
	o->info = ((uint64_t)((w2 >> 2) & pm) << 32) | (uint64_t)(((w2 >> (2 + pos_bits)) & pm) + 1);
	return f == esc;
}
int smem_gpu_collect_packed12(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv12_t *intv_out, int64_t intv_cap,
                              uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap, int64_t *n_exc_out, int32_t *pos_bits_out,
                              int64_t *total_out);
int smem_gpu_fetch_packed12(smem_gpu_t *h, smem_intv12_t *intv_out, int64_t intv_cap, uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap,
                            int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out);
/* The same record squeezed into 11 bytes, byte-packed (for reads of up to 512 bases): b[0..3] = low word of x[0], b[4..7] = low word
 * of x[1] (little endian), b[8..10] = the third word of smem_intv12_t with its size field cut to 22 - 2 P bits (8 bits = sizes up to
 * 255 for reads of up to 128 bases; larger ones go to the same exception list).  What it is for: the last 8 % of the link. */
typedef struct { uint8_t b[11]; } smem_intv11_t;
static inline int smem_intv11_unpack(const smem_intv11_t *p, int pos_bits, smem_intv_t *o)
{
	const uint32_t w0 = (uint32_t)p->b[0] | ((uint32_t)p->b[1] << 8) | ((uint32_t)p->b[2] << 16) | ((uint32_t)p->b[3] << 24);
	const uint32_t w1 = (uint32_t)p->b[4] | ((uint32_t)p->b[5] << 8) | ((uint32_t)p->b[6] << 16) | ((uint32_t)p->b[7] << 24);
	const uint32_t w2 = (uint32_t)p->b[8] | ((uint32_t)p->b[9] << 8) | ((uint32_t)p->b[10] << 16);
	const uint32_t pm = ((uint32_t)1 << pos_bits) - 1, esc = ((uint32_t)1 << (22 - 2 * pos_bits)) - 1;
	const uint32_t f = w2 >> (2 + 2 * pos_bits);
	o->x[0] = (uint64_t)w0 | ((uint64_t)(w2 & 1) << 32);
	o->x[1] = (uint64_t)w1 | ((uint64_t)((w2 >> 1) & 1) << 32);
	o->x[2] = f == esc ? 0 : (uint64_t)f + 1;
	o->info = ((uint64_t)((w2 >> 2) & pm) << 32) | (uint64_t)(((w2 >> (2 + pos_bits)) & pm) + 1);
	return f == esc;
}
int smem_gpu_collect_packed11(smem_gpu_t *h, const smem_reads2_t *reads, const smem_seed_opt_t *opt, smem_intv11_t *intv_out, int64_t intv_cap,
                              uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap, int64_t *n_exc_out, int32_t *pos_bits_out,
                              int64_t *total_out);
int smem_gpu_fetch_packed11(smem_gpu_t *h, smem_intv11_t *intv_out, int64_t intv_cap, uint32_t *read_off, smem_x2exc_t *exc_out, int64_t exc_cap,
                            int64_t *n_exc_out, int32_t *pos_bits_out, int64_t *total_out);
/* Host-side packer for callers that hold bwa's one-byte-per-base reads (n_threads host threads; lens and amb may be NULL if
 * the batch has one length / the caller knows there are no ambiguous bases -- then an ambiguous base is an error).
 * *n_amb_out = entries written (or needed, with SMEM_GPU_E_CAPACITY). */
int smem_gpu_pack_reads(int64_t n_reads, const uint8_t *seq, const int64_t *offs, int32_t stride, uint8_t *seq2, uint16_t *lens,
                        smem_amb_t *amb, int64_t amb_cap, int64_t *n_amb_out, int n_threads);

/* Pinned host memory for the caller's batches (replaces the MPF-VTP workspace buffers). */
int smem_gpu_host_alloc(void **ptr, size_t bytes);
int smem_gpu_host_free(void *ptr);

/* Measurement hooks (CUDA events on the library's own streams; max over the handle's devices). */
typedef struct {
	double seed_kernel_ms;     /* the SMEM kernel of the last run (first launch; the overflow re-run is not in it) */
	double total_device_ms;    /* pack pre-pass + seed kernel + scan + compaction (+ overflow re-run) */
	int64_t kernel_launches;   /* kernels launched by the last run, all devices */
	int64_t overflow_reads;    /* reads that needed the large-slot re-run */
	int64_t h2d_bytes, d2h_bytes;  /* bytes copied by the last stage / fetch */
} smem_gpu_timing_t;
int smem_gpu_last_timing(const smem_gpu_t *h, smem_gpu_timing_t *t);

/* Tuning knobs: "blocks_per_sm", "slot_cap", "b_cap" (prev/curr entries kept in shared
 * memory per read), "force_wide" (32-byte entries), "l2_fetch_granularity" (32|64|128, device-wide
 * cudaLimitMaxL2FetchGranularity hint), "probe_variant", "l2_hot_min_intv" (0 = off:
 * occ-block loads for intervals of size >= value carry an L2 evict_last hint), "spare_sms" / "chain_lanes" (scheduling of several pipeline lanes on one GPU), "repeat_filter" /
 * "spec_walk" / "unique_walk" (0 = do not use that shortcut of DESIGN.md section 10; results are the same),
 * "unique_walk_min_run" / "unique_walk_min_left" (a unique walk starts after that many extends of an interval of size 1
 * and with at least that many read bases left; 3 / 8), "count_skips" (debug counters "pass2_skipped", "unique_walks").
 * "lanes_per_read" -- the kernel variant (DESIGN.md sections 4 and 11; results are identical): 4 = a lane pair per read on the
 * 32-byte sector form of the index, the backward sweep split over the two lanes (default; indices with >= 2^32 occurrences of a
 * base take 2), 2 = a lane pair per read sharing 64-byte blocks (set BEFORE smem_gpu_upload_index to skip building the sector
 * form, + 0.5 byte per symbol of HBM), 3 = lane pairs on the sector form without the split, 1 = one lane per read on the sector
 * form; "sa_from_tables" (1 = seed positions from the full suffix array of the unique-walk tables when they are resident),
 * "tiny_path" (1 = batches of up to 2048 reads through smem_gpu_collect / _smem1 / _trace take the latency path).
 * "blocks_per_sm" accepts the launch-bounds variants that were compiled in (6, 8, 9).  Read-only: "table_bytes" /
 * "uw_table_bytes" / "index_bytes" (HBM of device 0), and the wall-time accumulators of the one-call forms, summed over the
 * handle's lanes: "acc_stage_us" (H2D), "acc_turn_us" (waiting for the GPU's kernel turn), "acc_run_us" (kernels),
 * "acc_fetch_us" (D2H), "acc_calls", "acc_h2d_bytes", "acc_d2h_bytes"; set "acc_reset" to clear them. */
int smem_gpu_set_param(smem_gpu_t *h, const char *name, int64_t value);
int64_t smem_gpu_get_param(const smem_gpu_t *h, const char *name);

/* Random-access roofline micro-benchmark: dependent chains of aligned
 * `block_bytes` (32 or 64) gathers, uniform over the first `span_bytes` of the uploaded index
 * (0 = all of it), `chains_per_sm` chains in flight per SM, `steps` gathers per chain.
 * *gbps_out = bytes gathered / CUDA-event time on device 0 of the handle. */
int smem_gpu_gather_roofline(smem_gpu_t *h, int block_bytes, uint64_t span_bytes, int chains_per_sm,
                             int steps, double *gbps_out);

const char *smem_gpu_strerror(int code);
const char *smem_gpu_last_error(const smem_gpu_t *h);   /* detail of the last failure on this handle */
int smem_gpu_device_count(void);

#ifdef __cplusplus
}
#endif
#endif
