/* oracle/gd_oracle.c -- TEST INFRASTRUCTURE ONLY (see gd_oracle.h / README.md).
 *
 * Scalar restatement of
 *   - the ksw2 dual-affine banded extension DP   (GDiet-ShortReads/ksw2_extd2_sse.c:27-401,
 *     score row of GDiet-ShortReads/ksw2_extd2_avx.c:187-208,312-313, helpers ksw2.h:100-188)
 *   - sparsified minimizer sketching             (GDiet-ShortReads/sketch.c:20-34,1577-2225)
 * written one cell / one position at a time, with none of the reference's SIMD structure.
 * What IS kept, because it is observable in the output, is the reference's 16-cell range
 * rounding and its persistent per-target-column int8 state.
 */
#include "gd_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* DP                                                                                          */
/* ------------------------------------------------------------------------------------------ */

#define F_SCORE_ONLY 0x01
#define F_RIGHT 0x02
#define F_GENERIC_SC 0x04
#define F_APPROX_MAX 0x08
#define F_APPROX_DROP 0x10
#define F_EXTZ_ONLY 0x40
#define F_REV_CIGAR 0x80

static inline int8_t w8(int v) { return (int8_t)(uint8_t)(v & 0xff); } /* int8 wrap-around */

static void extz_reset(gdo_extz_t *ez) /* ksw2.h:165-170 */
{
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0;
	ez->score = ez->mqe = ez->mte = GDO_NEG_INF;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
}

/* ksw2.h:172-188 with is_rot=1. ez->max is a 31-bit unsigned field there; H > max >= 0 keeps it so. */
static int zdrop_test(gdo_extz_t *ez, int32_t H, int r, int t, int zdrop, int e)
{
	if (H > ez->max) {
		ez->max = H & 0x7fffffff, ez->max_t = t, ez->max_q = r - t;
	} else if (t >= ez->max_t && r - t >= ez->max_q) {
		int tl = t - ez->max_t, ql = (r - t) - ez->max_q, l = tl > ql ? tl - ql : ql - tl;
		if (zdrop >= 0 && ez->max - H > zdrop + l * e) {
			ez->zdropped = 1;
			return 1;
		}
	}
	return 0;
}

typedef struct {
	uint32_t *a;
	int n, cap, overflow;
} cig_t;

static void cig_push(cig_t *c, uint32_t op, int len) /* ksw2.h:100-111 */
{
	if (c->n == 0 || op != (c->a[c->n - 1] & 0xf)) {
		if (c->n == c->cap) {
			c->overflow = 1;
			return;
		}
		c->a[c->n++] = (uint32_t)len << 4 | op;
	} else c->a[c->n - 1] += (uint32_t)len << 4;
}

/* ksw2.h:115-163 for the rotated matrix: p row r holds cells off[r]..off_end[r] */
static void traceback(const uint8_t *p, const int *off, const int *off_end, size_t row_bytes, int i0, int j0,
                      int keep_reversed, cig_t *c)
{
	int i = i0, j = j0, state = 0, k;
	while (i >= 0 && j >= 0) {
		int r = i + j, force = -1;
		uint32_t cell;
		if (i < off[r]) force = 2;
		if (i > off_end[r]) force = 1;
		cell = force < 0 ? p[(size_t)r * row_bytes + (size_t)(i - off[r])] : 0;
		if (state == 0) state = cell & 7;
		else if (!((cell >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = cell & 7;
		if (force >= 0) state = force;
		if (state == 0) cig_push(c, 0, 1), --i, --j;           /* M */
		else if (state == 1 || state == 3) cig_push(c, 2, 1), --i; /* D */
		else cig_push(c, 1, 1), --j;                           /* I */
	}
	if (i >= 0) cig_push(c, 2, i + 1);
	if (j >= 0) cig_push(c, 1, j + 1);
	if (!keep_reversed)
		for (k = 0; k < c->n >> 1; ++k) {
			uint32_t t = c->a[k];
			c->a[k] = c->a[c->n - 1 - k], c->a[c->n - 1 - k] = t;
		}
}

static int8_t init_gap_delta(int r, int q, int e, int e2, int long_thres, int long_diff)
{ /* first-column / first-row difference value, ksw2_extd2_sse.c:158,162 */
	return w8(r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2);
}

int gdo_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int q_,
                  int e_, int q2_, int e2_, int w, int zdrop, int end_bonus, int flag, int score_rule,
                  gdo_extz_t *ez, uint32_t *cigar, int cigar_cap)
{
	int8_t q = (int8_t)q_, e = (int8_t)e_, q2 = (int8_t)q2_, e2 = (int8_t)e2_;
	const int with_cigar = !(flag & F_SCORE_ONLY), approx = !!(flag & F_APPROX_MAX), right = !!(flag & F_RIGHT);
	int T16, ncol16, nrows, r, t, long_thres, long_diff, min_sc, last_st = -1, last_en = -1;
	/* score_rule: 0 = SSE build, 1 = AVX-512 build (the parity target), 3 = AVX-512 score row on the SSE build's
	 * 16-aligned rows (test aid: tells which pairs the lead-in cells change).  score_rule == 1 is the AVX-512 build: besides its xor-table score row, it works on 64-cell vectors whose first one
	 * starts at st0 rounded down to 64 (ksw2_extd2_avx.c:242,383); see the core loop below */
	const int lead = score_rule == 1 ? 64 : 16;
	int8_t *u, *v, *x, *y, *x2, *y2, *s, sc_mch, sc_mis, sc_N, qe, qe2, pmat[16];
	int32_t *H = 0, H0 = 0;
	int H0_t = 0;
	/* quirk kept on purpose: the reference evaluates `int qe = q + e` BEFORE it orders the two gap
	 * pieces (ksw2_extd2_sse.c:68,78) and uses that value to seed H (lines 358,382) */
	const int qe_seed = (int8_t)q_ + (int8_t)e_;
	uint8_t *p = 0;
	int *off = 0, *off_end = 0;
	cig_t cg;

	extz_reset(ez);
	if (m <= 1 || qlen <= 0 || tlen <= 0) return 0;
	if (q2 + e2 < q + e) { /* order the two gap pieces, ksw2_extd2_sse.c:78 */
		int8_t tq = q, te = e;
		q = q2, e = e2, q2 = tq, e2 = te;
	}
	qe = w8(q + e), qe2 = w8(q2 + e2);
	sc_mch = mat[0], sc_mis = mat[1];
	sc_N = mat[m * m - 1] == 0 ? w8(-e2) : mat[m * m - 1];
	memset(pmat, 0, 16);
	pmat[0] = sc_mch, pmat[1] = pmat[2] = pmat[3] = sc_mis;
	for (t = 4; t <= 12; ++t) pmat[t] = sc_N;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	T16 = (tlen + lead - 1) / lead * lead; /* state arrays: tlen_ vectors (ksw2_extd2_sse.c:91, ksw2_extd2_avx.c:144) */
	ncol16 = qlen < tlen ? qlen : tlen;
	ncol16 = ((ncol16 < w + 1 ? ncol16 : w + 1) + lead - 1) / lead + 1; /* ksw2_extd2_sse.c:92-93, ksw2_extd2_avx.c:145-146 */
	ncol16 *= lead;
	for (t = 1, min_sc = mat[1]; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
	if (-min_sc > 2 * (q + e)) return 0;

	long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	long_diff = long_thres * (e - e2) - (q2 - q) - e2;

	nrows = qlen + tlen - 1;
	u = (int8_t *)malloc((size_t)T16 * 7);
	v = u + T16, x = v + T16, y = x + T16, x2 = y + T16, y2 = x2 + T16, s = y2 + T16;
	memset(u, w8(-q - e), (size_t)T16 * 4);
	memset(x2, w8(-q2 - e2), (size_t)T16 * 2);
	memset(s, 0, (size_t)T16);
	if (!approx) {
		H = (int32_t *)malloc((size_t)T16 * 4);
		for (t = 0; t < T16; ++t) H[t] = GDO_NEG_INF;
	}
	if (with_cigar) {
		p = (uint8_t *)calloc((size_t)nrows * ncol16, 1);
		off = (int *)malloc(sizeof(int) * 2 * nrows);
		off_end = off + nrows;
	}

	for (r = 0; r < nrows; ++r) {
		int st0 = 0, en0 = tlen - 1, st, en, ext_end, stv;
		int8_t x1, x21, v1, bx1, bx21, bv1;
		if (st0 < r - qlen + 1) st0 = r - qlen + 1;
		if (en0 > r) en0 = r;
		if (st0 < (r - w + 1) >> 1) st0 = (r - w + 1) >> 1;
		if (en0 > (r + w) >> 1) en0 = (r + w) >> 1;
		if (st0 > en0) {
			ez->zdropped = 1;
			break;
		}
		st = st0 / 16 * 16, en = (en0 + 16) / 16 * 16 - 1; /* the range the state arrays are updated on */
		/* left boundary of the row */
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) x1 = x[st - 1], x21 = x2[st - 1], v1 = v[st - 1];
			else x1 = w8(-q - e), x21 = w8(-q2 - e2), v1 = w8(-q - e);
		} else {
			x1 = w8(-q - e), x21 = w8(-q2 - e2);
			v1 = init_gap_delta(r, q, e, e2, long_thres, long_diff);
		}
		if (en >= r) {
			y[r] = w8(-q - e), y2[r] = w8(-q2 - e2);
			u[r] = init_gap_delta(r, q, e, e2, long_thres, long_diff);
		}
		/* score row: rewritten from st0 in whole 16-cell steps (may run past en0), everything else stale */
		if (!(flag & F_GENERIC_SC)) {
			ext_end = st0 + (en0 - st0 + 16) / 16 * 16; /* exclusive */
			for (t = st0; t < ext_end && t < T16; ++t) {
				int qi = r - t; /* query index of cell (r,t); outside [0,qlen) reads zero padding */
				uint8_t tc = t < tlen ? target[t] : 0, qc = (qi >= 0 && qi < qlen) ? query[qi] : 0;
				if (!(score_rule & 1)) {
					s[t] = (tc == m - 1 || qc == m - 1) ? sc_N : (tc == qc ? sc_mch : sc_mis);
				} else {
					uint8_t idx = tc ^ (qc == 4 ? 8 : qc);
					s[t] = (idx & 0x80) ? 0 : pmat[idx & 15];
				}
			}
		} else {
			for (t = st0; t <= en0; ++t) {
				uint8_t qc = query[r - t];
				s[t] = mat[target[t] * m + qc];
			}
		}
		/* core update over the rounded range.  The AVX-512 build starts its first 64-cell vector at stv = st0/64*64 and
		 * stores every lane of it (ksw2_extd2_avx.c:242,383,442-476,497-553): the cells of [stv, st) are updated like
		 * any other cell -- from the stale scores and the state left in those columns -- their backtrack bytes are
		 * written, and off[r] = stv, so ksw_backtrack reads them where the SSE build forces an insertion
		 * (ksw2.h:136).  The boundary x1/v1/x21 is blended into the lane of column st (mskc_ar, :36,99); lane 0 of
		 * that vector receives byte 15 of its own 128-bit lane from the in-lane byte shuffle (index[0] = 15, :88-93),
		 * i.e. the previous row's x/v/x2 of column stv+15.  Nothing a later row reads as a true cell depends on these
		 * columns (st never decreases and last_st stays 16-aligned, :884). */
		stv = st0 / lead * lead;
		bx1 = x1, bv1 = v1, bx21 = x21;
		if (stv < st) x1 = x[stv + 15], v1 = v[stv + 15], x21 = x2[stv + 15];
		if (with_cigar) off[r] = stv, off_end[r] = en;
		for (t = stv; t <= en; ++t) {
			int8_t z, a, b, a2, b2, u_old, tq, tq2, nx1, nv1, nx21;
			uint8_t d = 0;
			if (t == st) x1 = bx1, v1 = bv1, x21 = bx21;
			z = s[t], a = w8(x1 + v1), b = w8(y[t] + u[t]), a2 = w8(x21 + v1), b2 = w8(y2[t] + u[t]);
			u_old = u[t];
			nx1 = x[t], nv1 = v[t], nx21 = x2[t]; /* become the left neighbours of column t+1 */
			if (!right) {
				if (a > z) d = 1, z = a;
				if (b > z) d = 2, z = b;
				if (a2 > z) d = 3, z = a2;
				if (b2 > z) d = 4, z = b2;
			} else {
				if (a >= z) d = 1, z = a;
				if (b >= z) d = 2, z = b;
				if (a2 >= z) d = 3, z = a2;
				if (b2 >= z) d = 4, z = b2;
			}
			if (z > sc_mch) z = sc_mch;
			u[t] = w8(z - v1);
			v[t] = w8(z - u_old);
			tq = w8(z - q), tq2 = w8(z - q2);
			a = w8(a - tq), b = w8(b - tq), a2 = w8(a2 - tq2), b2 = w8(b2 - tq2);
			if (!right) {
				if (a > 0) d |= 0x08; else a = 0;
				if (b > 0) d |= 0x10; else b = 0;
				if (a2 > 0) d |= 0x20; else a2 = 0;
				if (b2 > 0) d |= 0x40; else b2 = 0;
			} else {
				if (a >= 0) d |= 0x08; else a = 0;
				if (b >= 0) d |= 0x10; else b = 0;
				if (a2 >= 0) d |= 0x20; else a2 = 0;
				if (b2 >= 0) d |= 0x40; else b2 = 0;
			}
			x[t] = w8(a - qe), y[t] = w8(b - qe), x2[t] = w8(a2 - qe2), y2[t] = w8(b2 - qe2);
			if (with_cigar) p[(size_t)r * ncol16 + (t - stv)] = d;
			x1 = nx1, v1 = nv1, x21 = nx21;
		}
		/* score tracking */
		if (!approx) { /* ksw2_extd2_sse.c:323-366 */
			int32_t max_H, max_t;
			if (r > 0) {
				int32_t best[4], best_t[4];
				int en1 = st0 + (en0 - st0) / 4 * 4, i;
				max_H = H[en0] = en0 > 0 ? H[en0 - 1] + u[en0] : H[en0] + v[en0];
				max_t = en0;
				for (i = 0; i < 4; ++i) best[i] = max_H, best_t[i] = max_t;
				for (t = st0; t < en1; ++t) { /* four interleaved running maxima, strict > */
					i = (t - st0) & 3;
					H[t] += v[t];
					if (H[t] > best[i]) best[i] = H[t], best_t[i] = t;
				}
				for (i = 0; i < 4; ++i)
					if (max_H < best[i]) max_H = best[i], max_t = best_t[i];
				for (t = en1; t < en0; ++t) {
					H[t] += v[t];
					if (H[t] > max_H) max_H = H[t], max_t = t;
				}
			} else H[0] = v[0] - qe_seed, max_H = H[0], max_t = 0;
			if (en0 == tlen - 1 && H[en0] > ez->mte) ez->mte = H[en0], ez->mte_q = r - en;
			if (r - st0 == qlen - 1 && H[st0] > ez->mqe) ez->mqe = H[st0], ez->mqe_t = st0;
			if (zdrop_test(ez, max_H, r, max_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H[tlen - 1];
		} else { /* ksw2_extd2_sse.c:367-383 */
			if (r > 0) {
				int in0 = H0_t >= st0 && H0_t <= en0, in1 = H0_t + 1 >= st0 && H0_t + 1 <= en0;
				if (in0 && in1) {
					int32_t d0 = v[H0_t], d1 = u[H0_t + 1];
					if (d0 > d1) H0 += d0;
					else H0 += d1, ++H0_t;
				} else if (in0) {
					H0 += v[H0_t];
				} else {
					++H0_t, H0 += u[H0_t];
				}
			} else H0 = v[0] - qe_seed, H0_t = 0;
			if ((flag & F_APPROX_DROP) && zdrop_test(ez, H0, r, H0_t, zdrop, e2)) break;
			if (r == qlen + tlen - 2 && en0 == tlen - 1) ez->score = H0;
		}
		last_st = st, last_en = en;
	}
	free(u);
	free(H);
	cg.a = cigar, cg.n = 0, cg.cap = cigar ? cigar_cap : 0, cg.overflow = 0;
	if (with_cigar) { /* ksw2_extd2_sse.c:389-400 */
		int rev = !!(flag & F_REV_CIGAR);
		if (!ez->zdropped && !(flag & F_EXTZ_ONLY)) {
			traceback(p, off, off_end, ncol16, tlen - 1, qlen - 1, rev, &cg);
		} else if (!ez->zdropped && (flag & F_EXTZ_ONLY) && ez->mqe + end_bonus > ez->max) {
			ez->reach_end = 1;
			traceback(p, off, off_end, ncol16, ez->mqe_t, qlen - 1, rev, &cg);
		} else if (ez->max_t >= 0 && ez->max_q >= 0) {
			traceback(p, off, off_end, ncol16, ez->max_t, ez->max_q, rev, &cg);
		}
		free(p);
		free(off);
	}
	ez->n_cigar = cg.n;
	return cg.overflow ? -1 : cg.n;
}

int64_t gdo_band_cells(int qlen, int tlen, int w)
{
	int64_t cells = 0;
	int r;
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	for (r = 0; r < qlen + tlen - 1; ++r) {
		int st0 = 0, en0 = tlen - 1;
		if (st0 < r - qlen + 1) st0 = r - qlen + 1;
		if (en0 > r) en0 = r;
		if (st0 < (r - w + 1) >> 1) st0 = (r - w + 1) >> 1;
		if (en0 > (r + w) >> 1) en0 = (r + w) >> 1;
		if (st0 > en0) break;
		cells += en0 - st0 + 1;
	}
	return cells;
}

int gdo_exact_match(int qlen, const uint8_t *query, int tlen, const uint8_t *target)
{ /* exact_match_sse.c:27-88: the first qlen bytes are compared (call sites pass qlen == tlen) */
	if (qlen <= 0 || tlen <= 0) return 0;
	return memcmp(query, target, (size_t)qlen) == 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Sketching                                                                                   */
/* ------------------------------------------------------------------------------------------ */

uint64_t gdo_hash64(uint64_t key, uint64_t mask) /* sketch.c:25-34 */
{
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

static int nt4(unsigned char c) /* sketch.c:11-18 */
{
	switch (c) {
	case 'A': case 'a': case 0: return 0;
	case 'C': case 'c': case 1: return 1;
	case 'G': case 'g': case 2: return 2;
	case 'T': case 't': case 'U': case 'u': case 3: return 3;
	default: return 4;
	}
}

/* Position-parallel model: X[i], Y[i], run[i] for every sparsified position, then
 * emit i iff it is the minimum of at least one full window that contains it. */
long gdo_sketch_core(const char *str, unsigned len_crop, int w, int k, uint32_t rid, const char *Z, int W,
                     unsigned shift, uint64_t cap, uint64_t *out_xy, long out_cap, uint64_t *last_y)
{
	int ones = 0, ones_loc[64], g;
	unsigned diet_len, rem, i;
	uint64_t mask = (1ULL << 2 * k) - 1, fw = 0, rv = 0, *X, *Y;
	uint32_t *run, l = 0;
	long n = 0;
	for (g = 0; g < W && ones < 64; ++g)
		if (Z[g] == '1') ones_loc[ones++] = g;
	if (ones == 0 || len_crop < shift) return 0;
	diet_len = ((len_crop - shift) / (unsigned)W) * (unsigned)ones;
	rem = (len_crop - shift) % (unsigned)W;
	for (i = 0; i < rem; ++i)
		if (Z[i] == '1') ++diet_len;
	if (diet_len == 0) return 0;
	X = (uint64_t *)malloc(sizeof(uint64_t) * 2 * diet_len);
	Y = X + diet_len;
	run = (uint32_t *)malloc(sizeof(uint32_t) * diet_len);
	for (i = 0; i < diet_len; ++i) {
		unsigned real = (i / (unsigned)ones) * (unsigned)W + (unsigned)ones_loc[i % (unsigned)ones] + shift;
		int c = nt4((unsigned char)str[real]);
		X[i] = UINT64_MAX, Y[i] = UINT64_MAX;
		if (c < 4) {
			fw = (fw << 2 | (uint64_t)c) & mask;
			rv = (rv >> 2) | (uint64_t)(3 ^ c) << (2 * (k - 1));
			++l;
			if (l >= (uint32_t)k && fw != rv) {
				int z = fw < rv ? 0 : 1;
				X[i] = gdo_hash64(z ? rv : fw, mask) << 8 | (uint64_t)k;
				Y[i] = (uint64_t)rid << 32 | (uint64_t)(uint32_t)real << 1 | (uint64_t)z;
			}
		} else l = 0;
		run[i] = l;
	}
	for (i = 0; i < diet_len; ++i) {
		unsigned e, hi = i + (unsigned)w - 1 < diet_len - 1 ? i + (unsigned)w - 1 : diet_len - 1;
		int emit = 0;
		if (X[i] == UINT64_MAX) continue;
		for (e = i; e <= hi && !emit; ++e) {
			unsigned j;
			uint64_t mn = UINT64_MAX;
			if (run[e] < (uint32_t)(w + k - 1)) continue; /* window ending at e is not full */
			for (j = e + 1 - (unsigned)w; j <= e; ++j) mn = X[j] < mn ? X[j] : mn;
			emit = mn == X[i];
		}
		if (emit) {
			if (n < out_cap) out_xy[2 * n] = X[i], out_xy[2 * n + 1] = Y[i];
			if (last_y) *last_y = Y[i];
			++n;
			if (cap && (uint64_t)n == cap) break;
		}
	}
	free(X);
	free(run);
	return n;
}

long gdo_mm_sketch(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, uint64_t *out_xy,
                   long cap)
{ /* sketch.c:156 / 1577 */
	return gdo_sketch_core(str, (unsigned)len, w, k, rid, Z, W, 0, 0, out_xy, cap, 0);
}

long gdo_mm_sketch3(const char *str, unsigned len, int w, int k, uint32_t rid, const char *Z, int W, int shift,
                    uint32_t max_nb_seeds, uint64_t *out_xy, long cap, uint32_t *ret)
{ /* sketch.c:1078 / 1908: truncate at max_nb_seeds, return the position (y>>1) of the capping entry */
	uint64_t last = 0;
	long n = gdo_sketch_core(str, len, w, k, rid, Z, W, shift < 0 ? 0u : (unsigned)shift, max_nb_seeds, out_xy, cap,
	                         &last);
	*ret = (n > 0 && (uint64_t)n == (uint64_t)max_nb_seeds) ? (uint32_t)(last >> 1) : len;
	return n;
}

long gdo_mm_sketch2(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, float max_seeds,
                    uint64_t *out_xy, long cap, uint32_t *counts)
{ /* sketch.c:2143-2225 */
	unsigned len_crop;
	uint32_t cap_seeds;
	long total = 0;
	int shift;
	if (max_seeds < 1) len_crop = (unsigned)((float)max_seeds * len), cap_seeds = UINT32_MAX;
	else len_crop = (unsigned)len, cap_seeds = (uint32_t)max_seeds;
	for (shift = 0; shift < W; ++shift) {
		long room = cap - total > 0 ? cap - total : 0;
		long n = gdo_sketch_core(str, len_crop, w, k, rid, Z, W, (unsigned)shift, cap_seeds, out_xy + 2 * total, room, 0);
		counts[shift] = (uint32_t)n;
		total += n;
		if (cap_seeds == UINT32_MAX) len_crop = (unsigned)len, cap_seeds = (uint32_t)n;
	}
	return total;
}
