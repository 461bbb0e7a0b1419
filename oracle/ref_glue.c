/* oracle/ref_glue.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Thin flat-argument wrappers around the UNMODIFIED reference functions so that
 * Python (ctypes) tests and bench.py's cpu_baseline / --impl reference leg can
 * call them.  The reference sources are compiled where they lie under
 * /root/reference by oracle/Makefile into oracle/_ref/ (git-ignored); this file
 * only includes their public headers and supplies the three globals that the
 * reference defines in main.c / misc.c (GDiet-ShortReads/main.c:13-14,
 * GDiet-ShortReads/misc.c:5).
 *
 * Built twice:
 *   -DREF_HAVE_AVX512  -> _ref/libgdref_avx.so     (ksw_extd2_avx512 + AVX-512 sketch + SSE4.1 ksw)
 *   (none)             -> _ref/libgdref_scalar.so  (scalar sketch + SSE4.1 ksw)
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include "ksw2.h"
#include "minimap.h"
#include "mmpriv.h"
#include "kalloc.h"
#ifdef REF_HAVE_AVX512
#include "ksw2_extd2_avx.h"
#endif

/* globals the hot-path objects expect from main.c / misc.c */
int mm_dbg_flag = 0;
uint64_t km_size = 0;
int km_top = 0;

/* flat view of ksw_extz_t, shared with oracle/gd_oracle.h and include/gdiet_cuda.h */
typedef struct {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
} ref_extz_flat_t;

static void flatten(const ksw_extz_t *ez, ref_extz_flat_t *o)
{
	o->max = (int32_t)ez->max; o->zdropped = ez->zdropped;
	o->max_q = ez->max_q; o->max_t = ez->max_t;
	o->mqe = ez->mqe; o->mqe_t = ez->mqe_t; o->mte = ez->mte; o->mte_q = ez->mte_q;
	o->score = ez->score; o->n_cigar = ez->n_cigar; o->reach_end = ez->reach_end;
}

int ref_has_avx512(void)
{
#ifdef REF_HAVE_AVX512
	return 1;
#else
	return 0;
#endif
}

/* which: 0 = ksw_extd2_sse (SSE4.1 build), 1 = ksw_extd2_avx512.
 * Returns n_cigar (or -1 if the cigar did not fit cigar_cap). */
int ref_ksw_extd2(int which, void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m,
                  const int8_t *mat, int q, int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag,
                  ref_extz_flat_t *out, uint32_t *cigar, int cigar_cap)
{
	ksw_extz_t ez;
	int n;
	memset(&ez, 0, sizeof(ez));
#ifdef REF_HAVE_AVX512
	if (which == 1)
		ksw_extd2_avx512(km, qlen, query, tlen, target, (int8_t)m, mat, (int8_t)q, (int8_t)e, (int8_t)q2, (int8_t)e2,
		                 w, zdrop, end_bonus, flag, &ez);
	else
#endif
		ksw_extd2_sse(km, qlen, query, tlen, target, (int8_t)m, mat, (int8_t)q, (int8_t)e, (int8_t)q2, (int8_t)e2, w,
		              zdrop, end_bonus, flag, &ez);
	flatten(&ez, out);
	n = ez.n_cigar;
	if (cigar && n > 0) {
		if (n <= cigar_cap) memcpy(cigar, ez.cigar, (size_t)n * 4);
		else n = -1;
	}
	kfree(km, ez.cigar);
	return n;
}

int ref_exact_match(int qlen, const uint8_t *query, int tlen, const uint8_t *target)
{
	ksw_extz_t ez;
	bool em = false;
	int mm = 0;
	int8_t mat[25] = {0};
	memset(&ez, 0, sizeof(ez));
	exact_match_sse(0, qlen, query, tlen, target, 5, mat, 0, 0, 0, 0, 0, 0, &ez, &em, &mm);
	return em ? 1 : 0;
}

/* ---- batched, multi-threaded driver (the CPU baseline of bench.py) ---- */
typedef struct {
	int which, n, m, q, e, q2, e2, w, zdrop, end_bonus, flag;
	const int32_t *qlen, *tlen;
	const int64_t *qoff, *toff;
	const uint8_t *qbuf, *tbuf;
	const int8_t *mat;
	ref_extz_flat_t *out;
	uint32_t *cigar;   /* n * cigar_stride, may be NULL */
	int cigar_stride;
	volatile long next;
} ref_batch_t;

static void *batch_worker(void *arg)
{
	ref_batch_t *b = (ref_batch_t *)arg;
	void *km = km_init();
	for (;;) {
		long i = __sync_fetch_and_add(&b->next, 64), j, hi;
		if (i >= b->n) break;
		hi = i + 64 < b->n ? i + 64 : b->n;
		for (j = i; j < hi; ++j)
			ref_ksw_extd2(b->which, km, b->qlen[j], b->qbuf + b->qoff[j], b->tlen[j], b->tbuf + b->toff[j], b->m,
			              b->mat, b->q, b->e, b->q2, b->e2, b->w, b->zdrop, b->end_bonus, b->flag, &b->out[j],
			              b->cigar ? b->cigar + (size_t)j * b->cigar_stride : 0, b->cigar_stride);
	}
	km_destroy(km);
	return 0;
}

int ref_ksw_extd2_batch(int which, int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf,
                        const int32_t *tlen, const int64_t *toff, const uint8_t *tbuf, int m, const int8_t *mat, int q,
                        int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag, ref_extz_flat_t *out,
                        uint32_t *cigar, int cigar_stride, int n_threads)
{
	ref_batch_t b;
	pthread_t *tid;
	int i;
	if (n_threads < 1) n_threads = 1;
	b.which = which, b.n = n, b.m = m, b.q = q, b.e = e, b.q2 = q2, b.e2 = e2, b.w = w, b.zdrop = zdrop;
	b.end_bonus = end_bonus, b.flag = flag, b.qlen = qlen, b.tlen = tlen, b.qoff = qoff, b.toff = toff;
	b.qbuf = qbuf, b.tbuf = tbuf, b.mat = mat, b.out = out, b.cigar = cigar, b.cigar_stride = cigar_stride;
	b.next = 0;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, batch_worker, &b);
	for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
	free(tid);
	return 0;
}

/* ---- sketching ---- */
static long copy_out(mm128_v *v, uint64_t *out_xy, long cap)
{
	long n = (long)v->n;
	if (n <= cap && n > 0) memcpy(out_xy, v->a, (size_t)n * 16);
	free(v->a);
	return n;
}

/* mm_sketch: returns the number of minimizers (written as x,y pairs if it fits cap) */
long ref_mm_sketch(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, uint64_t *out_xy,
                   long cap)
{
	mm128_v v = {0, 0, 0};
	mm_sketch(0, str, len, w, k, rid, 0, &v, Z, W);
	return copy_out(&v, out_xy, cap);
}

/* mm_sketch3: *ret receives the function's return value */
long ref_mm_sketch3(const char *str, unsigned len, int w, int k, uint32_t rid, const char *Z, int W, int shift,
                    uint32_t max_nb_seeds, uint64_t *out_xy, long cap, uint32_t *ret)
{
	mm128_v v = {0, 0, 0};
	*ret = mm_sketch3(0, str, len, w, k, rid, 0, &v, Z, W, shift, max_nb_seeds);
	return copy_out(&v, out_xy, cap);
}

/* mm_sketch2: counts[W] receives shift_seeds_number */
long ref_mm_sketch2(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, float max_seeds,
                    uint64_t *out_xy, long cap, uint32_t *counts)
{
	mm128_v v = {0, 0, 0};
	mm_pattern_t p = mm_sketch2(0, str, len, w, k, rid, 0, &v, Z, W, max_seeds);
	memcpy(counts, p.shift_seeds_number, sizeof(uint32_t) * (size_t)W);
	free(p.shift_seeds_number);
	return copy_out(&v, out_xy, cap);
}

/* batched mm_sketch over many sequences, multi-threaded (CPU baseline for the sketch leg) */
typedef struct {
	int n, w, k, W;
	const char *Z;
	const int64_t *off;
	const int32_t *len;
	const char *buf;
	int64_t *counts;
	volatile long next;
} ref_skb_t;

static void *sk_worker(void *arg)
{
	ref_skb_t *b = (ref_skb_t *)arg;
	mm128_v v = {0, 0, 0};
	for (;;) {
		long i = __sync_fetch_and_add(&b->next, 16), j, hi;
		if (i >= b->n) break;
		hi = i + 16 < b->n ? i + 16 : b->n;
		for (j = i; j < hi; ++j) {
			v.n = 0;
			mm_sketch(0, b->buf + b->off[j], b->len[j], b->w, b->k, (uint32_t)j, 0, &v, b->Z, b->W);
			b->counts[j] = (int64_t)v.n;
		}
	}
	free(v.a);
	return 0;
}

int ref_mm_sketch_batch(int n, const int64_t *off, const int32_t *len, const char *buf, int w, int k, const char *Z,
                        int W, int64_t *counts, int n_threads)
{
	ref_skb_t b;
	pthread_t *tid;
	int i;
	if (n_threads < 1) n_threads = 1;
	b.n = n, b.w = w, b.k = k, b.W = W, b.Z = Z, b.off = off, b.len = len, b.buf = buf, b.counts = counts, b.next = 0;
	tid = (pthread_t *)malloc(sizeof(pthread_t) * n_threads);
	for (i = 0; i < n_threads; ++i) pthread_create(&tid[i], 0, sk_worker, &b);
	for (i = 0; i < n_threads; ++i) pthread_join(tid[i], 0);
	free(tid);
	return 0;
}
