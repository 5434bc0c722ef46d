/* TEST / MEASUREMENT INFRASTRUCTURE ONLY -- not part of the shipped product path.
 *
 * bwa_ref_timed = the unmodified reference `bwa` with its own bwt_smem1_batched (software/bwt.c:444-774, renamed at
 * compile time with -Dbwt_smem1_batched=harp_ref_bwt_smem1_batched_unused, see Makefile) called through this wrapper,
 * which only adds up the wall time spent inside it: the reference's CPU seeding time inside `bwa mem`, the figure the
 * drop-in adapter's own timer (SMEM_GPU_ADAPTER_STATS, adapter_s) is compared with in tools/dropin_bench.py. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

void harp_ref_bwt_smem1_batched_unused(void **itr, int *ori_start, int *max_i, int start_width, int is_middle, int batch_size,
                                       const int *done, int status);
static uint64_t g_ns, g_calls;
static int g_hooked;

static void report(void) { fprintf(stderr, "[ref_timed] seed_s=%.3f calls=%llu\n", (double)g_ns * 1e-9, (unsigned long long)g_calls); }

void bwt_smem1_batched(void **itr, int *ori_start, int *max_i, int start_width, int is_middle, int batch_size, const int *done, int status)
{
	struct timespec a, b;
	if (!__atomic_exchange_n(&g_hooked, 1, __ATOMIC_RELAXED)) atexit(report);
	clock_gettime(CLOCK_MONOTONIC, &a);
	harp_ref_bwt_smem1_batched_unused(itr, ori_start, max_i, start_width, is_middle, batch_size, done, status);
	clock_gettime(CLOCK_MONOTONIC, &b);
	__atomic_fetch_add(&g_ns, (uint64_t)((b.tv_sec - a.tv_sec) * 1000000000ll + (b.tv_nsec - a.tv_nsec)), __ATOMIC_RELAXED);
	__atomic_fetch_add(&g_calls, 1, __ATOMIC_RELAXED);
}
