/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's SMEM seeding path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The CUDA product path never calls it.  Parity status: PINNED -- every function here is
 * checked against the reference's own objects compiled from /root/reference/software
 * (oracle/_ref/libbwaref.so, tests/test_oracle_vs_ref.py) and against the golden fixtures
 * under tests/golden/ that were dumped from those objects (tests/golden/make_golden.py).
 */
#ifndef SMEM_ORACLE_H
#define SMEM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* The fields of bwt_t (bwt.h:46-58) the path reads. `bwt` is the packed array of bwt_t::bwt. */
typedef struct {
	uint64_t primary, L2[5], seq_len, bwt_size;
	const uint32_t *bwt;
} orc_index_t;

/* Seeding options of mem_opt_t (bwamem.h:33-60) used by mem_insert_seed (bwamem.c:453-458). */
typedef struct {
	int min_seed_len;
	double split_factor;
	int split_width;
	int start_width;
} orc_seed_opt_t;

/* Work counters behind the algorithmic-bytes figure of SURVEY.md section 8(d). */
typedef struct {
	uint64_t extends;      /* bwt_extend calls the reference algorithm executes */
	uint64_t blocks;       /* 64-byte occ blocks: 1 per extend if k',l' share a block else 2 */
	uint64_t smem1_calls;  /* bwt_smem1 invocations (pass 1 + re-seed) */
	uint64_t steps;        /* smem_next2 calls that returned a list */
	uint64_t intervals;    /* intervals emitted */
	uint64_t max_curr;     /* largest prev/curr population seen */
	uint64_t max_mem;      /* largest per-call mem population seen */
} orc_stats_t;

void orc_occ4(const orc_index_t *ix, uint64_t k, uint64_t cnt[4]);
void orc_extend(const orc_index_t *ix, const uint64_t ik3[3], int is_back, uint64_t ok12[12]);

/* Same flat conventions as oracle/ref_harness.c. Return total interval count (may exceed cap:
 * then only read_off is complete). */
int64_t orc_smem1(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs, const int32_t *x,
                  const int32_t *min_intv, uint64_t *intv, int64_t cap, int64_t *read_off, int32_t *ret);
int64_t orc_collect(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs, const orc_seed_opt_t *opt,
                    int nthreads, uint64_t *intv, int64_t cap, int64_t *read_off, uint16_t *step,
                    int32_t *n_steps, int32_t *last_start, orc_stats_t *stats);
/* Timing arm ("kind": "port"): wall seconds, checksum as in ref_time_collect. */
double orc_time_collect(const orc_index_t *ix, int64_t n, const uint8_t *seq, const int64_t *offs,
                        const orc_seed_opt_t *opt, int nthreads, uint64_t *checksum, int64_t *n_intervals);
/* Seed -> reference position: bwt_sa (bwt.c:104-114) over bwt_invPsi (bwt.c:71-77)
 * and bwt_occ (bwt.c:125-147).  sa = the samples of bwt_cal_sa (bwt.c:79-101), sa[0] = -1. */
uint64_t orc_occ(const orc_index_t *ix, uint64_t k, int c);
void orc_sa(const orc_index_t *ix, int sa_intv, const uint64_t *sa, int64_t n, const uint64_t *k, uint64_t *out);

/* Seeds and chains.  orc_seed_t == mem_seed_t (bwamem.c:316-319). */
typedef struct { int64_t rbeg; int32_t qbeg, len; } orc_seed_t;
/* chaining / chain-filter fields of mem_opt_t (bwamem.h:33-60; defaults w 100, max_chain_gap 10000, 0.5, 0.5) */
typedef struct { int w, max_chain_gap, min_seed_len; float mask_level, chain_drop_ratio; } orc_chain_opt_t;
int64_t orc_seeds(const orc_index_t *ix, int sa_intv, const uint64_t *sa, int64_t n, const uint64_t *intv, const int64_t *read_off,
                  int min_seed_len, int64_t max_occ, orc_seed_t *seeds, int64_t cap, int64_t *seed_off);
/* mem_chain's insertion loop (bwamem.c:478-496, test_and_merge :334-356) and, if do_flt, mem_chain_flt (bwamem.c:629-700).
 * chain_off int64[n+1]; chain int64[total][2] = {pos, n_seeds}; out_seeds grouped by chain.  Returns the chain total. */
int64_t orc_chains(int64_t n, const orc_seed_t *seeds, const int64_t *seed_off, int64_t l_pac, const orc_chain_opt_t *o, int do_flt,
                   int64_t *chain_off, int64_t *chain, int64_t chain_cap, orc_seed_t *out_seeds, int64_t seed_cap, int64_t *n_seeds_out);

/* Order-independent-across-reads checksum of a flat result (same definition as the timing arms). */
uint64_t orc_checksum(int64_t n, const uint64_t *intv, const int64_t *read_off);

#ifdef __cplusplus
}
#endif
#endif
