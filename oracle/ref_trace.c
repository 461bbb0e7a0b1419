/* oracle/ref_trace.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Call tracer for the UNMODIFIED reference program.  oracle/Makefile compiles every reference source
 * where it lies under /root/reference into oracle/_ref/GDiet_avx_{sr,lr}; four translation units get
 * one symbol renamed on the compiler command line (-Dname=gdref_real_name):
 *
 *     ksw2_extd2_avx.c : ksw_extd2_avx512      (GDiet-ShortReads/ksw2_extd2_avx.c:72)
 *     ksw2_dispatch.c  : exact_match_sse       (GDiet-ShortReads/ksw2_dispatch.c:55)
 *     align.c          : mm_update_extra       (GDiet-ShortReads/align.c:259)
 *     sketch.c         : mm_sketch2            (GDiet-ShortReads/sketch.c:2143)
 *
 * and this file supplies those four names: each forwards to the real function and, when the
 * environment variable GDREF_TRACE names a file, appends one binary record with the call's arguments
 * and results.  The records are what pins the seeding / voting / window stage of the device mapper
 * (SURVEY.md section 8 rows F1, F2): for every read, the exact (query, target, band) of every DP and
 * exact-match call the reference makes at GDiet-ShortReads/map.c:873-929 and the candidate window
 * (rid, qs, qe, rs, re, rev) it hands to mm_update_extra at map.c:954.  Use with -t 1 so that the
 * records of one read are contiguous and reads appear in input order.
 *
 * Record layout (little endian int32 unless noted): kind, then
 *   1 READ    len, bytes[len]                       one per mm_map_frag call (mm_sketch2 is its first call)
 *   2 EXACT   len, q[len], t[len], exact
 *   3 KSW     qlen, tlen, w, zdrop, end_bonus, flag, q, e, q2, e2, q[qlen], t[tlen], score, n_cigar, cigar[n_cigar]
 *   4 EXTRA   rid, score, qs, qe, rs, re, rev       state of the candidate BEFORE mm_update_extra edits it
 */
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>
#include "minimap.h"
#include "mmpriv.h"
#include "ksw2.h"

void gdref_real_ksw_extd2_avx512(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                                 const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                                 int end_bonus, int flag, ksw_extz_t *ez);
void gdref_real_exact_match_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                                const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag,
                                ksw_extz_t *ez, bool *exact_match, int *mismatch_cnt);
void gdref_real_mm_update_extra(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q,
                                int8_t e, int is_eqx, int log_gap);
mm_pattern_t gdref_real_mm_sketch2(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc,
                                   mm128_v *p, const char *Z, int W, const float max_seeds);

static FILE *g_fp;
static int g_init;
static pthread_mutex_t g_mu = PTHREAD_MUTEX_INITIALIZER;

static FILE *trace_fp(void)
{
	if (!g_init) {
		const char *fn = getenv("GDREF_TRACE");
		g_init = 1;
		if (fn && *fn) g_fp = fopen(fn, "wb");
	}
	return g_fp;
}
static void put32(FILE *fp, int32_t v) { fwrite(&v, 4, 1, fp); }

mm_pattern_t mm_sketch2(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p,
                        const char *Z, int W, const float max_seeds)
{
	FILE *fp;
	pthread_mutex_lock(&g_mu);
	if ((fp = trace_fp()) != 0) {
		put32(fp, 1), put32(fp, len);
		fwrite(str, 1, len, fp);
	}
	pthread_mutex_unlock(&g_mu);
	return gdref_real_mm_sketch2(km, str, len, w, k, rid, is_hpc, p, Z, W, max_seeds);
}

void exact_match_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                     const int8_t *mat, int8_t q, int8_t e, int w, int zdrop, int end_bonus, int flag, ksw_extz_t *ez,
                     bool *exact_match, int *mismatch_cnt)
{
	FILE *fp;
	gdref_real_exact_match_sse(km, qlen, query, tlen, target, m, mat, q, e, w, zdrop, end_bonus, flag, ez, exact_match,
	                           mismatch_cnt);
	pthread_mutex_lock(&g_mu);
	if ((fp = trace_fp()) != 0) {
		put32(fp, 2), put32(fp, qlen);
		fwrite(query, 1, qlen, fp), fwrite(target, 1, qlen, fp);
		put32(fp, *exact_match ? 1 : 0);
	}
	pthread_mutex_unlock(&g_mu);
}

void ksw_extd2_avx512(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                      const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus,
                      int flag, ksw_extz_t *ez)
{
	FILE *fp;
	gdref_real_ksw_extd2_avx512(km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, ez);
	pthread_mutex_lock(&g_mu);
	if ((fp = trace_fp()) != 0) {
		put32(fp, 3), put32(fp, qlen), put32(fp, tlen), put32(fp, w), put32(fp, zdrop), put32(fp, end_bonus);
		put32(fp, flag), put32(fp, q), put32(fp, e), put32(fp, q2), put32(fp, e2);
		fwrite(query, 1, qlen, fp), fwrite(target, 1, tlen, fp);
		put32(fp, ez->score), put32(fp, ez->n_cigar);
		if (ez->n_cigar > 0) fwrite(ez->cigar, 4, ez->n_cigar, fp);
	}
	pthread_mutex_unlock(&g_mu);
}

void mm_update_extra(mm_reg1_t *r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e,
                     int is_eqx, int log_gap)
{
	FILE *fp;
	pthread_mutex_lock(&g_mu);
	if ((fp = trace_fp()) != 0) {
		put32(fp, 4), put32(fp, r->rid), put32(fp, r->score), put32(fp, r->qs), put32(fp, r->qe);
		put32(fp, r->rs), put32(fp, r->re), put32(fp, r->rev);
		fflush(fp);
	}
	pthread_mutex_unlock(&g_mu);
	gdref_real_mm_update_extra(r, qseq, tseq, mat, q, e, is_eqx, log_gap);
}
