/* bwa_abi.h -- ABI mirrors of the reference structs that cross the drop-in boundary.
 *
 * The link-compatible adapter (bwa-mem-harp2_b200/host/bwt_smem1_batched_gpu.c) is compiled without the
 * reference's headers, so the few types it receives from the unchanged bwamem.c are re-declared here with
 * the exact layout of the reference (x86-64 SysV, bwa 0.7.8 as in TianheYu/bwa-mem-harp2):
 *   harp_bwtintv_t   == bwtintv_t   software/bwt.h:60-62
 *   harp_bwtintv_v   == bwtintv_v   software/bwt.h:64            (kvec: n, m, a; grown x2 from 2, kvec.h:75-81)
 *   harp_bwt_t       == bwt_t       software/bwt.h:46-58
 *   harp_smem_i      == smem_i      software/bwamem.h:17-24
 * tests/test_abi.py checks sizes/offsets against the reference headers when /root/reference is present.
 */
#ifndef BWA_ABI_H
#define BWA_ABI_H
#include <stdint.h>
#include <stddef.h>

typedef uint64_t harp_bwtint_t;

typedef struct { harp_bwtint_t x[3], info; } harp_bwtintv_t;
typedef struct { size_t n, m; harp_bwtintv_t *a; } harp_bwtintv_v;

typedef struct {
	harp_bwtint_t primary;
	harp_bwtint_t L2[5];
	harp_bwtint_t seq_len;
	harp_bwtint_t bwt_size;
	uint32_t *bwt;
	uint32_t cnt_table[256];
	int sa_intv;
	harp_bwtint_t n_sa;
	harp_bwtint_t *sa;
} harp_bwt_t;

typedef struct {
	const harp_bwt_t *bwt;
	const uint8_t *query;
	int start, len;
	harp_bwtintv_v *matches;
	harp_bwtintv_v *sub;
	harp_bwtintv_v *tmpvec[2];
} harp_smem_i;

/* bwt_batched_status values, software/bwamem.h:12-14 */
#define HARP_BWT_BATCHED_INIT 0
#define HARP_BWT_BATCHED_FREE 1
#define HARP_BWT_BATCHED_DO   2

#ifdef __cplusplus
extern "C" {
#endif

/* The symbol the unchanged bwamem.c calls (bwamem.c:172,204,1647,1657; defined in the reference at bwt.c:444).
 * Arguments are declared with the mirror types; the ABI is identical. */
void bwt_smem1_batched(harp_smem_i **itr, int *ori_start, int *max_i, int start_width, int is_middle,
                       int batch_size, const int *done, int bwt_batched_status);

/* Optional explicit control of the GPU service behind the adapter (otherwise created lazily on the first DO
 * call from itr[0]->bwt).  devices: comma list in SMEM_GPU_DEVICES (default "0"). Returns 0 or SMEM_GPU_E_*. */
int harp_gpu_service_start(const harp_bwt_t *bwt);
void harp_gpu_service_stop(void);
/* Counters for tests: GPU calls, reads sent, intervals received, lists served from the whole-batch cache.
 * Environment: SMEM_GPU_ADAPTER_CACHE=0 disables the cache; SMEM_GPU_SPLIT_LEN / SMEM_GPU_SPLIT_WIDTH tell it the
 * caller's re-seeding options (default 28 / 10 = -k 19 -r 1.5); a mismatch only costs speed, never correctness. */
void harp_gpu_service_stats(uint64_t out[4]);
/* ... plus: requests combined into those calls, nanoseconds inside the GPU calls, nanoseconds inside bwt_smem1_batched
 * (summed over the calling threads), reserved.  SMEM_GPU_ADAPTER_STATS=1 prints all of them at exit. */
void harp_gpu_service_stats8(uint64_t out[8]);
/* bwt_smem1_batched returns void (bwt.c:444), so a GPU failure cannot be handed to the caller: it is reported once on
 * stderr, the affected lists come back EMPTY (there is no CPU fallback), and the code of the last failure stays readable here
 * (0 = none; *msg = its text).  SMEM_GPU_ADAPTER_ABORT=1 makes a failure abort the process instead.
 * Other environment: SMEM_GPU_ADAPTER_HANDLES (service handles = concurrent launches, default 4), SMEM_GPU_MAX_BATCH (reads
 * per launch, default 65536), SMEM_GPU_MAX_READ_LEN (first size of the handles; they grow on demand). */
int harp_gpu_service_last_error(const char **msg);

#ifdef __cplusplus
}
#endif
#endif
