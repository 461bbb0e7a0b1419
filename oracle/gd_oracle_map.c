/* oracle/gd_oracle_map.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C restatement of the short-read mapper's stages either side of the two hot kernels
 * (SURVEY.md section 8 rows F1 "index lookup" and F2 "seed-hit sort + location voting"), i.e. what
 * GDiet-ShortReads/map.c:mm_map_frag does between mm_sketch2 and the ksw_extd2 call:
 *
 *   index          GDiet-ShortReads/index.c:84-100,216-304  (bucket/khash replaced by one sorted key array:
 *                  the lookup result -- count + positions sorted by y -- is the same set)
 *   shift          GDiet-ShortReads/seed.c:166-194 (mm_get_shift)
 *   seed filters   GDiet-ShortReads/seed.c:5-29 (mm_seed_mz_flt), :36-62,67-113,143-164
 *   hits + sort    GDiet-ShortReads/map.c:261-356
 *   voting         GDiet-ShortReads/map.c:447-584
 *   windows        GDiet-ShortReads/map.c:586-840
 *   exact match/DP GDiet-ShortReads/map.c:859-929
 *
 * Parity status: PINNED by tests/test_oracle_map_vs_ref.py against the call trace of the unmodified
 * reference program (oracle/ref_trace.c, oracle/_ref/GDiet_avx_sr) and by tests/golden/map_*.npz.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "gd_oracle.h"
#include "gd_oracle_map.h"

static int nt4c(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': case 0: return 0;
	case 'C': case 'c': case 1: return 1;
	case 'G': case 'g': case 2: return 2;
	case 'T': case 't': case 'U': case 'u': case 3: return 3;
	default: return 4;
	}
}

/* ------------------------------------------------------------------------------------------ */
/* index                                                                                        */
/* ------------------------------------------------------------------------------------------ */
static int cmp_xy(const void *a, const void *b)
{
	const uint64_t *p = (const uint64_t *)a, *q = (const uint64_t *)b;
	uint64_t kx = p[0] >> 8, ky = q[0] >> 8;
	if (kx != ky) return kx < ky ? -1 : 1;
	return p[1] < q[1] ? -1 : p[1] > q[1];
}

gdo_index_t *gdo_index_build(int n_seq, const char *buf, const int64_t *off, const int32_t *len, int w, int k,
                             const char *Z, int W)
{
	gdo_index_t *mi = (gdo_index_t *)calloc(1, sizeof(gdo_index_t));
	int64_t total = 0, n = 0, cap = 0, i, j;
	uint64_t *xy = 0;
	mi->n_seq = n_seq, mi->w = w, mi->k = k;
	mi->len = (uint32_t *)calloc(n_seq, 4), mi->offset = (uint64_t *)calloc(n_seq + 1, 8);
	for (i = 0; i < n_seq; ++i) mi->len[i] = len[i], mi->offset[i] = total, total += len[i];
	mi->offset[n_seq] = total;
	mi->codes = (uint8_t *)malloc(total + 1);
	for (i = 0; i < n_seq; ++i) /* index.c:351-356: the reference keeps nt4 codes, 4 bits each */
		for (j = 0; j < len[i]; ++j) mi->codes[mi->offset[i] + j] = nt4c((unsigned char)buf[off[i] + j]);
	for (i = 0; i < n_seq; ++i) { /* index.c:365-379 */
		long m;
		if (len[i] <= 0) continue;
		if (cap - n < len[i] + 16) cap = n + len[i] + 16, xy = (uint64_t *)realloc(xy, cap * 16);
		m = gdo_mm_sketch(buf + off[i], len[i], w, k, (uint32_t)i, Z, W, xy + 2 * n, cap - n);
		n += m;
	}
	qsort(xy, n, 16, cmp_xy); /* index.c:225 + :255: by minimizer, then by position */
	mi->n_pos = n;
	mi->pos = (uint64_t *)malloc((n + 1) * 8);
	mi->key = (uint64_t *)malloc((n + 1) * 8), mi->start = (uint64_t *)malloc((n + 1) * 8);
	mi->cnt = (uint32_t *)malloc((n + 1) * 4);
	for (i = 0; i < n; ++i) {
		mi->pos[i] = xy[2 * i + 1];
		if (i == 0 || xy[2 * i] >> 8 != xy[2 * i - 2] >> 8)
			mi->key[mi->n_keys] = xy[2 * i] >> 8, mi->start[mi->n_keys] = i, mi->cnt[mi->n_keys++] = 0;
		++mi->cnt[mi->n_keys - 1];
	}
	free(xy);
	return mi;
}

void gdo_index_destroy(gdo_index_t *mi)
{
	if (!mi) return;
	free(mi->len), free(mi->offset), free(mi->codes), free(mi->pos), free(mi->key), free(mi->start), free(mi->cnt);
	free(mi);
}

/* index.c:84-100 */
const uint64_t *gdo_index_get(const gdo_index_t *mi, uint64_t minier, int *n)
{
	int64_t lo = 0, hi = mi->n_keys - 1;
	*n = 0;
	while (lo <= hi) {
		int64_t mid = (lo + hi) >> 1;
		if (mi->key[mid] == minier) {
			*n = (int)mi->cnt[mid];
			return mi->pos + mi->start[mid];
		}
		if (mi->key[mid] < minier) lo = mid + 1;
		else hi = mid - 1;
	}
	return 0;
}

static int cmp_u32(const void *a, const void *b)
{
	uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
	return x < y ? -1 : x > y;
}

/* index.c:182-201 (mm_idx_cal_max_occ) */
int32_t gdo_index_cal_max_occ(const gdo_index_t *mi, float f)
{
	uint32_t *a, thres;
	if (f <= 0.) return INT32_MAX;
	if (mi->n_keys == 0) return 1;
	a = (uint32_t *)malloc(mi->n_keys * 4);
	memcpy(a, mi->cnt, mi->n_keys * 4);
	qsort(a, mi->n_keys, 4, cmp_u32);
	thres = a[(uint32_t)((1. - f) * mi->n_keys)] + 1;
	free(a);
	return (int32_t)thres;
}

/* ------------------------------------------------------------------------------------------ */
/* seeds                                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t n, q_pos, q_span, flt;
	const uint64_t *cr;
} seed_t;

typedef struct {
	uint64_t target;
	uint32_t query;
} loc_t;

static int cmp_loc(const void *a, const void *b)
{
	const loc_t *p = (const loc_t *)a, *q = (const loc_t *)b;
	if (p->target != q->target) return p->target < q->target ? -1 : 1;
	return p->query < q->query ? -1 : p->query > q->query; /* tie order does not reach the vote */
}

/* seed.c:5-29 */
static size_t mz_flt(uint64_t *xy, size_t n, int32_t q_occ_max, float q_occ_frac)
{
	size_t i, j, m;
	if (n <= (size_t)q_occ_max || q_occ_frac <= 0.0f || q_occ_max <= 0) return n;
	{ /* count on a copy so that zeroing does not disturb later counts */
		uint64_t *x = (uint64_t *)malloc(n * 8);
		for (i = 0; i < n; ++i) x[i] = xy[2 * i];
		for (i = 0; i < n; ++i) {
			int32_t cnt = 0;
			for (j = 0; j < n; ++j) cnt += x[j] == x[i];
			if (cnt > q_occ_max && cnt > n * q_occ_frac) xy[2 * i] = 0;
		}
		free(x);
	}
	for (i = m = 0; i < n; ++i)
		if (xy[2 * i] != 0) xy[2 * m] = xy[2 * i], xy[2 * m + 1] = xy[2 * i + 1], ++m;
	return m;
}

/* seed.c:67-113: in every streak of high-occurrence seeds keep the max_high_occ least frequent ones */
static void seed_select(int32_t n, seed_t *a, int len, int max_occ, int max_max_occ, int dist)
{
	int32_t i, last0, m;
	uint64_t b[128];
	if (n == 0 || n == 1) return;
	for (i = m = 0; i < n; ++i)
		if ((int32_t)a[i].n > max_occ) ++m;
	if (m == 0) return;
	for (i = 0, last0 = -1; i <= n; ++i) {
		if (i == n || (int32_t)a[i].n <= max_occ) {
			if (i - last0 > 1) {
				int32_t ps = last0 < 0 ? 0 : (int32_t)(a[last0].q_pos >> 1);
				int32_t pe = i == n ? len : (int32_t)(a[i].q_pos >> 1);
				int32_t j, k, st = last0 + 1, en = i;
				int32_t max_high_occ = (int32_t)((double)(pe - ps) / dist + .499);
				if (max_high_occ > 0) {
					if (max_high_occ > 128) max_high_occ = 128;
					for (j = st, k = 0; j < en && k < max_high_occ; ++j, ++k) b[k] = (uint64_t)a[j].n << 32 | j;
					for (; j < en; ++j) { /* the heap top of seed.c:94-99 is the maximum of b[] */
						int32_t t, top = 0;
						for (t = 1; t < k; ++t)
							if (b[t] > b[top]) top = t;
						if ((int32_t)a[j].n < (int32_t)(b[top] >> 32)) b[top] = (uint64_t)a[j].n << 32 | j;
					}
					for (j = 0; j < k; ++j) a[(uint32_t)b[j]].flt = 1;
				}
				for (j = st; j < en; ++j) a[j].flt ^= 1;
				for (j = st; j < en; ++j)
					if ((int32_t)a[j].n > max_max_occ) a[j].flt = 1;
			}
			last0 = i;
		}
	}
}

/* ------------------------------------------------------------------------------------------ */
/* voting, map.c:447-584                                                                        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	uint32_t chrom_id;
	int32_t target_loc;
	uint32_t first_query_loc, last_query_loc, str, score;
} vt_t;

static void vt_emit(vt_t *pot, unsigned *out_len, vt_t *recovery, uint64_t target_loc, uint32_t fq,
                    uint32_t lq, unsigned counter, int str, int32_t tmp_ext, unsigned thr, unsigned max_loc,
                    unsigned rec_thr)
{
	vt_t v;
	unsigned k;
	v.chrom_id = (uint32_t)(target_loc >> 32);
	v.target_loc = (int32_t)(target_loc & 0xffffffffu) + (str ? 0 : -tmp_ext);
	v.first_query_loc = fq, v.last_query_loc = lq, v.str = str, v.score = counter;
	if (counter > thr) {
		if (*out_len == max_loc) {
			if (pot[*out_len - 1].score >= counter) return;
		} else
			++*out_len;
		pot[*out_len - 1] = v;
		for (k = *out_len - 1; k > 0; k--) {
			if (pot[k].score > pot[k - 1].score) {
				vt_t t = pot[k];
				pot[k] = pot[k - 1], pot[k - 1] = t;
			} else
				break;
		}
	} else if (*out_len == 0 && counter > rec_thr && counter > recovery->score)
		*recovery = v;
}

static void vote(const loc_t *loc, unsigned len, int str, vt_t *pot, unsigned *nb, unsigned dist, int32_t tmp_ext,
                 vt_t *recovery, unsigned thr, unsigned max_loc, unsigned rec_thr)
{
	unsigned i, counter = 1, out_len = *nb;
	uint64_t target_loc;
	uint32_t fq, lq;
	if (len == 0) return;
	target_loc = loc[0].target, fq = lq = loc[0].query;
	for (i = 1; i < len; i++) {
		loc_t cur = loc[i];
		if (cur.target - target_loc <= dist) {
			counter++;
			if (cur.query < fq) target_loc = cur.target, fq = cur.query;
			if (cur.query > lq) lq = cur.query;
		} else {
			vt_emit(pot, &out_len, recovery, target_loc, fq, lq, counter, str, tmp_ext, thr, max_loc, rec_thr);
			target_loc = cur.target, fq = lq = cur.query, counter = 1;
		}
	}
	vt_emit(pot, &out_len, recovery, target_loc, fq, lq, counter, str, tmp_ext, thr, max_loc, rec_thr);
	*nb = out_len;
}

/* ------------------------------------------------------------------------------------------ */
/* one read, map.c:586-979 up to (not including) mm_update_extra                                */
/* ------------------------------------------------------------------------------------------ */
int gdo_sr_map_read(const gdo_index_t *mi, const char *seq, int qlen, const gdo_sr_opt_t *o, gdo_sr_cand_t *out,
                    int out_cap, uint32_t *cigar, int cigar_cap, gdo_sr_dbg_t *dbg)
{
	const int k = mi->k, w = mi->w;
	unsigned qlen_sum = (unsigned)qlen;
	unsigned bw = o->bw;
	long cap = qlen + 16, n3;
	uint64_t *mv = (uint64_t *)malloc((size_t)cap * 16 * (o->W + 1));
	uint32_t *counts = (uint32_t *)calloc(o->W + 1, 4);
	unsigned shift = 0, max_hits = 0, s, i;
	uint32_t max_nb_seeds = o->frag_mode ? (o->max_frag_len == 0 ? 800u : (uint32_t)o->max_frag_len) : UINT32_MAX;
	uint32_t tmp_ext = 0;
	size_t n_mv, n_m0, n_m;
	seed_t *m;
	int64_t n_a = 0;
	loc_t *a_for, *a_rev;
	unsigned n_for = 0, n_rev = 0, nb = 0, vt_threshold, vt_rec_threshold, n_out = 0;
	int cig_used = 0;
	vt_t *pot, recovery;
	uint8_t *qs_for, *qs_rev, *ts;
	gdo_sr_dbg_t d;
	memset(&d, 0, sizeof(d));
	memset(&recovery, 0, sizeof(recovery));
	if (qlen <= 0) { free(mv), free(counts); if (dbg) *dbg = d; return 0; }

	/* pattern alignment, map.c:609-615 + seed.c:166-194 */
	gdo_mm_sketch2(seq, qlen, w, k, 0, o->Z, o->W, o->max_seeds, mv, cap * (o->W + 1), counts);
	{
		uint64_t *p = mv;
		for (s = 0; s < (unsigned)o->W; ++s) {
			unsigned cur = 0, j;
			for (j = 0; j < counts[s]; ++j) {
				int t;
				gdo_index_get(mi, p[2 * j] >> 8, &t);
				cur += t;
			}
			if (cur > max_hits) shift = s, max_hits = cur;
			p += 2 * counts[s];
		}
	}
	d.shift = shift;

	/* seeding, map.c:634-648 */
	n3 = gdo_mm_sketch3(seq, (unsigned)qlen, w, k, 0, o->Z, o->W, (int)shift, max_nb_seeds, mv, cap, &tmp_ext);
	n_mv = (size_t)n3;
	d.tmp_extracted_len = tmp_ext;
	if (o->q_occ_frac > 0.0f) n_mv = mz_flt(mv, n_mv, o->mid_occ, o->q_occ_frac);
	d.n_mv = (uint32_t)n_mv;
	m = (seed_t *)calloc(n_mv + 1, sizeof(seed_t));
	for (i = 0, n_m0 = 0; i < n_mv; ++i) { /* seed.c:36-62 */
		int t;
		const uint64_t *cr = gdo_index_get(mi, mv[2 * i] >> 8, &t);
		if (t == 0) continue;
		m[n_m0].q_pos = (uint32_t)mv[2 * i + 1], m[n_m0].q_span = mv[2 * i] & 0xff, m[n_m0].cr = cr, m[n_m0].n = t;
		m[n_m0++].flt = 0;
	}
	if (o->occ_dist > 0 && o->max_max_occ > o->mid_occ) /* seed.c:143-164 */
		seed_select((int32_t)n_m0, m, qlen, o->mid_occ, o->max_max_occ, o->occ_dist);
	else
		for (i = 0; i < n_m0; ++i)
			if ((int32_t)m[i].n > o->mid_occ) m[i].flt = 1;
	for (i = 0, n_m = 0; i < n_m0; ++i)
		if (!m[i].flt) n_a += m[i].n, m[n_m++] = m[i];
	a_for = (loc_t *)malloc((n_a + 1) * sizeof(loc_t)), a_rev = (loc_t *)malloc((n_a + 1) * sizeof(loc_t));
	for (i = 0; i < n_m; ++i) { /* map.c:284-311 */
		uint32_t kk;
		for (kk = 0; kk < m[i].n; ++kk) {
			uint64_t r = m[i].cr[kk];
			uint32_t qpos = m[i].q_pos >> 1, loc = (uint32_t)r >> 1;
			unsigned str = (r & 1) ^ (m[i].q_pos & 1);
			if (str ? o->for_only : o->rev_only) continue; /* map.c:121-127 (skip_seed) */
			if (str) a_rev[n_rev].target = (r >> 32) << 32 | (uint32_t)(loc + qpos), a_rev[n_rev++].query = qpos;
			else a_for[n_for].target = (r >> 32) << 32 | (uint32_t)(loc + tmp_ext - qpos), a_for[n_for++].query = qpos;
		}
	}
	free(m);
	qsort(a_for, n_for, sizeof(loc_t), cmp_loc), qsort(a_rev, n_rev, sizeof(loc_t), cmp_loc);
	d.n_a_for = n_for, d.n_a_rev = n_rev;

	if (o->bw_max > 0) { /* map.c:624-631: the band and vote distance of THIS read */
		bw = (unsigned)(float)(qlen * o->bw_frac);
		if (o->bw_min > bw) bw = o->bw_min;
		else if (o->bw_max < bw) bw = o->bw_max;
	}
	/* voting, map.c:665-699 */
	{
		int frag = o->frag_mode && tmp_ext < qlen_sum;
		vt_threshold = frag ? (float)max_nb_seeds * o->min_cnt : (float)n_mv * o->min_cnt;
		vt_rec_threshold = frag ? (float)max_nb_seeds * o->rec_threshold_frac : (float)n_mv * o->rec_threshold_frac;
		if (vt_threshold == 0) vt_threshold = 1;
	}
	d.vt_threshold = vt_threshold;
	pot = (vt_t *)calloc(o->af_max_loc + 1, sizeof(vt_t));
	vote(a_for, n_for, 0, pot, &nb, bw, (int32_t)tmp_ext, &recovery, vt_threshold, o->af_max_loc, vt_rec_threshold);
	vote(a_rev, n_rev, 1, pot, &nb, bw, (int32_t)tmp_ext, &recovery, vt_threshold, o->af_max_loc, vt_rec_threshold);
	free(a_for), free(a_rev), free(mv), free(counts);
	if (nb == 0) {
		if (recovery.score == 0) { free(pot); if (dbg) *dbg = d; return 0; }
		nb = 1, pot[0] = recovery;
	}
	d.nb_potentials = nb;

	/* windows + DP, map.c:737-929 */
	qs_for = (uint8_t *)malloc(qlen_sum), qs_rev = (uint8_t *)malloc(qlen_sum), ts = (uint8_t *)calloc(qlen_sum + 1, 1);
	for (i = 0; i < qlen_sum; ++i) qs_for[i] = nt4c((unsigned char)seq[i]), qs_rev[qlen_sum - i - 1] = qs_for[i] ^ 3;
	for (i = 0; i < nb && (int)n_out < out_cap; ++i) {
		int str = pot[i].str;
		unsigned target_id = pot[i].chrom_id;
		uint32_t start_offset, end_offset, len, j;
		int32_t target_start, target_end, tlen = (int32_t)mi->len[target_id];
		const uint8_t *qs;
		gdo_sr_cand_t *c = &out[n_out];
		int8_t mat[25];
		gdo_extz_t ez;
		if (str) pot[i].target_loc -= (k - 1);
		target_start = target_end = pot[i].target_loc;
		if (qlen_sum > 300) {
			if (pot[i].first_query_loc == pot[i].last_query_loc) continue;
			start_offset = pot[i].first_query_loc - (k - 1);
			end_offset = pot[i].last_query_loc;
			if (str) {
				target_end -= start_offset, target_start -= end_offset;
				if (target_start < 0) end_offset += target_start, target_start = 0;
				qs = &qs_rev[qlen_sum - 1 - end_offset];
			} else {
				target_start += start_offset, target_end += end_offset;
				if (target_end + 1 > tlen) end_offset = tlen - 1 - target_start + start_offset, target_end = tlen - 1;
				qs = &qs_for[start_offset];
			}
		} else {
			if (str) {
				if (target_end > tlen - 1) start_offset = target_end - (tlen - 1), target_end = tlen - 1;
				else start_offset = 0;
				if ((unsigned)target_end < qlen_sum - start_offset - 1) end_offset = start_offset + target_end, target_start = 0;
				else end_offset = qlen_sum - 1, target_start = target_end - (end_offset - start_offset);
				qs = &qs_rev[qlen_sum - 1 - end_offset];
			} else {
				if (target_start < 0) start_offset = -target_start, target_start = 0;
				else start_offset = 0;
				if ((unsigned)(tlen - target_start) < qlen_sum - start_offset)
					end_offset = tlen - 1 - target_start + start_offset, target_end = tlen - 1;
				else end_offset = qlen_sum - 1, target_end = target_start + (end_offset - start_offset);
				qs = &qs_for[start_offset];
			}
		}
		len = end_offset - start_offset + 1;
		for (j = 0; j < len; ++j) { /* index.c:157-166 */
			int64_t p = (int64_t)target_start + j;
			if (p < tlen && p <= target_end) ts[j] = mi->codes[mi->offset[target_id] + p];
		}
		{
			int g = o->a, bb = o->b < 0 ? o->b : -o->b, x, y;
			for (x = 0; x < 5; ++x)
				for (y = 0; y < 5; ++y) mat[x * 5 + y] = (x == 4 || y == 4) ? 0 : (x == y ? g : bb);
		}
		memset(c, 0, sizeof(*c));
		c->rid = target_id, c->rs = target_start, c->re = target_end + 1, c->qs = start_offset, c->qe = end_offset + 1;
		c->rev = str, c->votes = pot[i].score, c->first_q = pot[i].first_query_loc, c->last_q = pot[i].last_query_loc;
		c->cigar_off = cig_used;
		if (qlen_sum < 300 && gdo_exact_match(len, qs, len, ts)) { /* map.c:873-915 */
			c->exact = 1, c->score = qlen_sum * o->a, c->n_cigar = 1;
			if (cig_used < cigar_cap) cigar[cig_used] = len << 4 | 0;
			cig_used += 1;
		} else {
			gdo_ksw_extd2(len, qs, len, ts, 5, mat, o->q, o->e, o->q2, o->e2, (int)bw, o->zdrop, o->end_bonus, 0x08, 1,
			                       &ez, cigar + cig_used, cigar_cap - cig_used);
			c->score = ez.score, c->n_cigar = ez.n_cigar;
			if (ez.n_cigar > 0) cig_used += ez.n_cigar;
		}
		++n_out;
	}
	free(qs_for), free(qs_rev), free(ts), free(pot);
	if (dbg) *dbg = d;
	return (int)n_out;
}

/* ========================================================================================== */
/* Long reads: GDiet-LongReads/map.c:mm_map_frag (:1273-1853) up to (not including)           */
/* mm_update_extra / concatenate_cigars                                                        */
/* ========================================================================================== */
typedef struct { /* vt_t, LR/map.c:1032-1045 */
	uint32_t chrom_id;
	int32_t first_target_loc, last_target_loc;
	uint32_t first_query_loc, last_query_loc;
	unsigned score;
	int next; /* index into seqs[] instead of a pointer; -1 = NULL */
	unsigned str, concat;
} lvt_t;

static int cmp_loc_merge(const void *a, const void *b)
{ /* the order LR/map.c:merge_sort (:176-255) leaves: ascending target; on ties the later unit (= larger query
     position, the seed list is position-ordered) first, because merge_locations takes the second run on a tie */
	const loc_t *p = (const loc_t *)a, *q = (const loc_t *)b;
	if (p->target != q->target) return p->target < q->target ? -1 : 1;
	return p->query > q->query ? -1 : p->query < q->query;
}

static inline uint64_t lr_loc(int str, uint64_t target, uint32_t query, int32_t tmp_ext)
{
	return str ? (target - query) : target - (uint64_t)(int64_t)(tmp_ext - (int32_t)query);
}

static void lr_emit(lvt_t *seqs, unsigned *out_len, unsigned max_loc, uint64_t ft, uint64_t lt, uint32_t fq, uint32_t lq,
                    unsigned counter, int str, int *skipped)
{
	unsigned k;
	*skipped = 0;
	if (*out_len == max_loc) {
		if (seqs[*out_len - 1].score >= counter) { *skipped = 1; return; }
	} else
		++*out_len;
	seqs[*out_len - 1].chrom_id = (uint32_t)(ft >> 32), seqs[*out_len - 1].first_target_loc = (int32_t)(uint32_t)ft;
	seqs[*out_len - 1].last_target_loc = (int32_t)(uint32_t)lt, seqs[*out_len - 1].first_query_loc = fq;
	seqs[*out_len - 1].last_query_loc = lq, seqs[*out_len - 1].str = str, seqs[*out_len - 1].score = counter;
	seqs[*out_len - 1].next = -1, seqs[*out_len - 1].concat = 0;
	for (k = *out_len - 1; k > 0; k--) {
		if (seqs[k].score > seqs[k - 1].score) {
			lvt_t t = seqs[k];
			seqs[k] = seqs[k - 1], seqs[k - 1] = t;
		} else
			break;
	}
}

/* LR/map.c:1052-1182 */
static void lr_vote(const loc_t *loc, unsigned len, int str, lvt_t *seqs, unsigned *nb, uint32_t dist, int32_t tmp_ext,
                    unsigned max_loc, uint32_t cov_thr)
{
	unsigned i, counter = 1, out_len = *nb;
	uint64_t ft, lt, ref_loc;
	uint32_t fq, lq;
	int skipped;
	if (len == 0) return;
	ft = lt = lr_loc(str, loc[0].target, loc[0].query, tmp_ext), fq = lq = loc[0].query, ref_loc = loc[0].target;
	for (i = 1; i < len; i++) {
		loc_t cur = loc[i];
		if (cur.target - ref_loc <= dist) {
			uint64_t l = lr_loc(str, cur.target, cur.query, tmp_ext);
			counter++;
			if (cur.query < fq) fq = cur.query, ref_loc = cur.target;
			if (cur.query > lq) lq = cur.query;
			if (l > lt) lt = l;
			if (l < ft) ft = l;
		} else {
			if (lq - fq > cov_thr) lr_emit(seqs, &out_len, max_loc, ft, lt, fq, lq, counter, str, &skipped);
			ft = lt = lr_loc(str, cur.target, cur.query, tmp_ext), fq = lq = cur.query, ref_loc = cur.target, counter = 1;
		}
	}
	if (lq - fq > cov_thr) lr_emit(seqs, &out_len, max_loc, ft, lt, fq, lq, counter, str, &skipped);
	*nb = out_len;
}

/* LR/map.c:1184-1271 */
static void lr_vote_2(const loc_t *loc, unsigned len, int str, lvt_t *vt, unsigned dist, int32_t tmp_ext, uint32_t min, uint32_t max)
{
	unsigned i, counter = 1;
	uint64_t ft, lt, ref_loc;
	uint32_t fq, lq;
	lvt_t best = *vt;
	if (len == 0) return;
	ft = lt = lr_loc(str, loc[0].target, loc[0].query, tmp_ext), fq = lq = loc[0].query, ref_loc = loc[0].target;
	for (i = 1; i <= len; i++) {
		if (i < len && loc[i].target - ref_loc <= dist) {
			loc_t cur = loc[i];
			if (cur.query < max && cur.query > min) {
				uint64_t l = lr_loc(str, cur.target, cur.query, tmp_ext);
				counter++;
				if (cur.query < fq) fq = cur.query, ref_loc = cur.target;
				if (cur.query > lq) lq = cur.query;
				if (l > lt) lt = l;
				if (l < ft) ft = l;
			}
		} else {
			if (counter > best.score && lq < max && fq > min) {
				best.chrom_id = (uint32_t)(ft >> 32), best.first_target_loc = (int32_t)(uint32_t)ft;
				best.last_target_loc = (int32_t)(uint32_t)lt, best.first_query_loc = fq, best.last_query_loc = lq;
				best.str = str, best.score = counter, best.next = -1, best.concat = 0;
			}
			if (i < len) {
				loc_t cur = loc[i];
				ft = lt = lr_loc(str, cur.target, cur.query, tmp_ext), fq = lq = cur.query, ref_loc = cur.target, counter = 1;
			}
		}
	}
	*vt = best;
}

static void lr_second_round(const gdo_lr_opt_t *o, int k, const loc_t *a_for, unsigned n_for, const loc_t *a_rev, unsigned n_rev,
                            int32_t tmp_ext, uint32_t min, uint32_t max, lvt_t *seqs, unsigned *nb)
{ /* LR/map.c:1402-1445 */
	lvt_t v2;
	const unsigned bw = o->bw;
	memset(&v2, 0, sizeof(v2));
	v2.next = -1;
	lr_vote_2(a_for, n_for, 0, &v2, o->vt_dis, tmp_ext, min, max);
	lr_vote_2(a_rev, n_rev, 1, &v2, o->vt_dis, tmp_ext, min, max);
	v2.first_query_loc -= (k - 1), v2.first_target_loc -= (k - 1);
	if ((float)v2.score > o->vt_df2 * (float)(v2.last_target_loc - v2.first_target_loc)) {
		if (v2.last_query_loc - v2.first_query_loc + 0.5 * bw < v2.last_target_loc - v2.first_target_loc)
			v2.last_target_loc = v2.first_target_loc + v2.last_query_loc - v2.first_query_loc + 0.5 * bw;
		seqs[(*nb)++] = v2;
	}
}

int gdo_lr_map_read(const gdo_index_t *mi, const char *seq, int qlen, const gdo_lr_opt_t *o, gdo_sr_cand_t *out, int out_cap,
                    uint32_t *cigar, int cigar_cap, gdo_sr_dbg_t *dbg)
{
	const int k = mi->k, w = mi->w;
	const unsigned qlen_sum = (unsigned)qlen, bw = o->bw;
	long cap = qlen + 16, n3;
	uint64_t *mv = (uint64_t *)malloc((size_t)cap * 16 * (o->W + 1));
	uint32_t *counts = (uint32_t *)calloc(o->W + 1, 4);
	unsigned shift = 0, max_hits = 0, s, i, j, nb = 0, n_for = 0, n_rev = 0, n_out = 0;
	uint32_t max_nb_seeds = o->frag_mode ? (o->max_frag_len == 0 ? 800u : (uint32_t)o->max_frag_len) : UINT32_MAX;
	uint32_t tmp_ext = 0, cov_thr, qrstart, qrend;
	size_t n_mv, n_m0, n_m;
	seed_t *m;
	int64_t n_a = 0;
	loc_t *a_for, *a_rev;
	lvt_t *seqs;
	int cig_used = 0;
	uint8_t *qs_for, *qs_rev, *ts;
	gdo_sr_dbg_t d;
	memset(&d, 0, sizeof(d));
	if (qlen <= 0) { free(mv), free(counts); if (dbg) *dbg = d; return 0; }
	/* pattern alignment + seeding: identical to the short-read tree (LR/map.c:1293-1330) */
	gdo_mm_sketch2(seq, qlen, w, k, 0, o->Z, o->W, o->max_seeds, mv, cap * (o->W + 1), counts);
	{
		uint64_t *p = mv;
		for (s = 0; s < (unsigned)o->W; ++s) {
			unsigned cur = 0;
			for (j = 0; j < counts[s]; ++j) {
				int t;
				gdo_index_get(mi, p[2 * j] >> 8, &t);
				cur += t;
			}
			if (cur > max_hits) shift = s, max_hits = cur;
			p += 2 * counts[s];
		}
	}
	d.shift = shift;
	n3 = gdo_mm_sketch3(seq, (unsigned)qlen, w, k, 0, o->Z, o->W, (int)shift, max_nb_seeds, mv, cap, &tmp_ext);
	n_mv = (size_t)n3;
	d.tmp_extracted_len = tmp_ext;
	if (o->q_occ_frac > 0.0f) n_mv = mz_flt(mv, n_mv, o->mid_occ, o->q_occ_frac);
	d.n_mv = (uint32_t)n_mv;
	m = (seed_t *)calloc(n_mv + 1, sizeof(seed_t));
	for (i = 0, n_m0 = 0; i < n_mv; ++i) {
		int t;
		const uint64_t *cr = gdo_index_get(mi, mv[2 * i] >> 8, &t);
		if (t == 0) continue;
		m[n_m0].q_pos = (uint32_t)mv[2 * i + 1], m[n_m0].q_span = mv[2 * i] & 0xff, m[n_m0].cr = cr, m[n_m0].n = t;
		m[n_m0++].flt = 0;
	}
	if (o->occ_dist > 0 && o->max_max_occ > o->mid_occ) seed_select((int32_t)n_m0, m, qlen, o->mid_occ, o->max_max_occ, o->occ_dist);
	else
		for (i = 0; i < n_m0; ++i)
			if ((int32_t)m[i].n > o->mid_occ) m[i].flt = 1;
	for (i = 0, n_m = 0; i < n_m0; ++i)
		if (!m[i].flt) n_a += m[i].n, m[n_m++] = m[i];
	a_for = (loc_t *)malloc((n_a + 1) * sizeof(loc_t)), a_rev = (loc_t *)malloc((n_a + 1) * sizeof(loc_t));
	for (i = 0; i < n_m; ++i) {
		uint32_t kk;
		for (kk = 0; kk < m[i].n; ++kk) {
			uint64_t r = m[i].cr[kk];
			uint32_t qpos = m[i].q_pos >> 1, loc = (uint32_t)r >> 1;
			unsigned str = (r & 1) ^ (m[i].q_pos & 1);
			if (str ? o->for_only : o->rev_only) continue;
			if (str) a_rev[n_rev].target = (r >> 32) << 32 | (uint32_t)(loc + qpos), a_rev[n_rev++].query = qpos;
			else a_for[n_for].target = (r >> 32) << 32 | (uint32_t)(loc + tmp_ext - qpos), a_for[n_for++].query = qpos;
		}
	}
	free(m), free(mv), free(counts);
	qsort(a_for, n_for, sizeof(loc_t), cmp_loc_merge), qsort(a_rev, n_rev, sizeof(loc_t), cmp_loc_merge);
	d.n_a_for = n_for, d.n_a_rev = n_rev;

	/* first voting round + density filter, LR/map.c:1342-1369 */
	cov_thr = (float)qlen_sum * o->vt_cov;
	d.vt_threshold = cov_thr;
	seqs = (lvt_t *)calloc(o->vt_nb_loc + 3, sizeof(lvt_t));
	lr_vote(a_for, n_for, 0, seqs, &nb, o->vt_dis, (int32_t)tmp_ext, o->vt_nb_loc, cov_thr);
	lr_vote(a_rev, n_rev, 1, seqs, &nb, o->vt_dis, (int32_t)tmp_ext, o->vt_nb_loc, cov_thr);
	if (nb > 0) {
		unsigned df = 0;
		for (i = 0; i < nb; i++)
			if ((float)seqs[i].score > o->vt_df1 * (float)(seqs[i].last_target_loc - seqs[i].first_target_loc)) seqs[i] = seqs[df], df++;
		nb = df;
	}
	if (nb == 0) { free(a_for), free(a_rev), free(seqs); if (dbg) *dbg = d; return 0; }
	/* score filter, k-mer start, band guard, covered query range, LR/map.c:1371-1400 */
	{
		const unsigned filtering_threshold = (float)seqs[0].score * o->vt_f;
		qrstart = qlen_sum, qrend = 0;
		for (i = 0; i < nb; i++) {
			if (seqs[i].score < filtering_threshold) { nb = i; break; }
			seqs[i].first_query_loc -= (k - 1), seqs[i].first_target_loc -= (k - 1);
			seqs[i].next = -1, seqs[i].concat = 0;
			if (seqs[i].last_query_loc - seqs[i].first_query_loc + 0.5 * bw < seqs[i].last_target_loc - seqs[i].first_target_loc)
				seqs[i].last_target_loc = seqs[i].first_target_loc + seqs[i].last_query_loc - seqs[i].first_query_loc + 0.5 * bw;
			if (seqs[i].first_query_loc < qrstart) qrstart = seqs[i].first_query_loc;
			if (seqs[i].last_query_loc > qrend) qrend = seqs[i].last_query_loc;
		}
	}
	/* second voting round on the uncovered ends, LR/map.c:1402-1445 */
	if (qrstart > cov_thr) lr_second_round(o, k, a_for, n_for, a_rev, n_rev, (int32_t)tmp_ext, 0, qrstart, seqs, &nb);
	if (qlen_sum - qrend > cov_thr) lr_second_round(o, k, a_for, n_for, a_rev, n_rev, (int32_t)tmp_ext, qrend, qlen_sum, seqs, &nb);
	free(a_for), free(a_rev);
	d.nb_potentials = nb;

	/* which candidates continue each other, LR/map.c:1467-1590 */
	{
		const unsigned max_max_gap = o->max_max_gap, max_min_gap = o->max_min_gap;
		for (i = 0; i < nb; i++) {
			lvt_t *s1 = &seqs[i];
			for (j = 0; j < nb; j++) {
				lvt_t *s2 = &seqs[j];
				int take = 0, better = 0;
				if (j == i || s2->concat != 0 || s1->str != s2->str || s1->chrom_id != s2->chrom_id) continue;
				if (s1->str) {
					if (s2->last_query_loc < s1->first_query_loc && s1->last_target_loc > s2->first_target_loc &&
					    s1->first_target_loc < s2->first_target_loc) {
						take = s2->last_query_loc + max_max_gap > s1->first_query_loc;
						better = s1->next >= 0 && s2->last_query_loc > seqs[s1->next].last_query_loc;
					} else if (s2->last_query_loc < s1->first_query_loc && s1->last_target_loc < s2->first_target_loc) {
						take = (s2->last_query_loc + max_min_gap > s1->first_query_loc ||
						        s1->last_target_loc + max_min_gap > (unsigned)s2->first_target_loc) &&
						       s2->last_query_loc + max_max_gap > s1->first_query_loc &&
						       s1->last_target_loc + max_max_gap > (unsigned)s2->first_target_loc;
						better = s1->next >= 0 && s2->last_query_loc > seqs[s1->next].last_query_loc;
					} else if (s2->last_query_loc > s1->first_query_loc && s1->last_target_loc < s2->first_target_loc &&
					           s2->last_query_loc < s1->last_query_loc && s2->first_query_loc < s1->first_query_loc) {
						take = s1->last_target_loc + max_max_gap > (unsigned)s2->first_target_loc;
						better = s1->next >= 0 && s2->last_query_loc < seqs[s1->next].last_query_loc;
					}
				} else {
					if (s1->last_query_loc < s2->first_query_loc && s1->last_target_loc > s2->first_target_loc &&
					    s1->first_target_loc < s2->first_target_loc) {
						take = s1->last_query_loc + max_max_gap > s2->first_query_loc;
						better = s1->next >= 0 && s2->first_query_loc < seqs[s1->next].first_query_loc;
					} else if (s1->last_query_loc < s2->first_query_loc && s1->last_target_loc < s2->first_target_loc) {
						take = (s1->last_query_loc + max_min_gap > s2->first_query_loc ||
						        s1->last_target_loc + max_min_gap > (unsigned)s2->first_target_loc) &&
						       s1->last_target_loc + max_max_gap > (unsigned)s2->first_target_loc &&
						       s1->last_query_loc + max_max_gap > s2->first_query_loc;
						better = s1->next >= 0 && s2->first_query_loc < seqs[s1->next].first_query_loc;
					} else if (s1->last_query_loc > s2->first_query_loc && s1->last_target_loc < s2->first_target_loc &&
					           s1->first_query_loc < s2->first_query_loc && s1->last_query_loc < s2->last_query_loc) {
						take = s1->last_target_loc + max_max_gap > (unsigned)s2->first_target_loc;
						better = s1->next >= 0 && s2->first_query_loc < seqs[s1->next].first_query_loc;
					}
				}
				if (take && (s1->next < 0 || better)) s1->next = (int)j;
			}
			if (s1->next >= 0) {
				lvt_t *s2 = &seqs[s1->next];
				s2->concat = 1;
				if (s1->str) {
					if (s2->last_query_loc < s1->first_query_loc && s1->last_target_loc < s2->first_target_loc) {
						const uint32_t diffq = s1->first_query_loc - s2->last_query_loc, difft = s2->first_target_loc - s1->last_target_loc;
						const uint32_t mn = difft > diffq ? diffq : difft;
						s2->last_query_loc += mn, s1->last_target_loc += mn, s1->first_query_loc -= mn, s2->first_target_loc -= mn;
					}
				} else {
					if (s1->last_query_loc < s2->first_query_loc && s1->last_target_loc < s2->first_target_loc) {
						const uint32_t diffq = s2->first_query_loc - s1->last_query_loc, difft = s2->first_target_loc - s1->last_target_loc;
						const uint32_t mn = difft > diffq ? diffq : difft;
						s1->last_query_loc += mn, s1->last_target_loc += mn, s2->first_query_loc -= mn, s2->first_target_loc -= mn;
					}
				}
				if (s2->last_target_loc < s1->last_target_loc) s1->last_target_loc = s2->last_target_loc - 1;
			}
		}
	}

	/* windows + DP, LR/map.c:1654-1805 */
	qs_for = (uint8_t *)malloc(qlen_sum), qs_rev = (uint8_t *)malloc(qlen_sum);
	for (i = 0; i < qlen_sum; ++i) qs_for[i] = nt4c((unsigned char)seq[i]), qs_rev[qlen_sum - i - 1] = qs_for[i] ^ 3;
	for (i = 0; i < nb && (int)n_out < out_cap; ++i) {
		const int str = seqs[i].str;
		const unsigned target_id = seqs[i].chrom_id;
		uint32_t target_start = seqs[i].first_target_loc, target_end = seqs[i].last_target_loc, query_start, query_end, ql, tl, jj;
		const uint8_t *qseq;
		gdo_sr_cand_t *c = &out[n_out];
		int8_t mat[25];
		gdo_extz_t ez;
		const int32_t chrom_len = (int32_t)mi->len[target_id];
		if (str) query_end = qlen_sum - 1 - seqs[i].first_query_loc, query_start = qlen_sum - 1 - seqs[i].last_query_loc;
		else query_start = seqs[i].first_query_loc, query_end = seqs[i].last_query_loc;
		if (!(qlen_sum > 300)) {
			if (target_start < query_start) query_start -= target_start, target_start = 0;
			else target_start -= query_start, query_start = 0;
			if ((uint32_t)chrom_len + query_end < qlen_sum + target_end) query_end += chrom_len - target_end - 1, target_end = chrom_len - 1;
			else target_end += qlen_sum - query_end - 1, query_end = qlen_sum - 1;
		}
		qseq = str ? &qs_rev[query_start] : &qs_for[query_start];
		ql = query_end - query_start + 1, tl = target_end - target_start + 1;
		if (str) {
			uint32_t tmp = qlen_sum - 1 - query_start;
			query_start = qlen_sum - 1 - query_end, query_end = tmp;
		}
		ts = (uint8_t *)calloc((size_t)tl + 1, 1);
		for (jj = 0; jj < tl; ++jj) { /* mm_idx_getseq: clipped at the contig end, index.c:157-166 */
			int64_t p = (int64_t)target_start + jj;
			if (p < chrom_len) ts[jj] = mi->codes[mi->offset[target_id] + p];
		}
		{
			int g = o->a, bb = o->b < 0 ? o->b : -o->b, x, y;
			for (x = 0; x < 5; ++x)
				for (y = 0; y < 5; ++y) mat[x * 5 + y] = (x == 4 || y == 4) ? 0 : (x == y ? g : bb);
		}
		memset(c, 0, sizeof(*c));
		c->rid = target_id, c->rs = target_start, c->re = target_end + 1, c->qs = query_start, c->qe = query_end + 1, c->rev = str;
		c->votes = seqs[i].score, c->first_q = seqs[i].first_query_loc, c->last_q = seqs[i].last_query_loc;
		c->reserved[0] = seqs[i].next, c->reserved[1] = seqs[i].concat;
		c->cigar_off = cig_used;
		if (qlen_sum < 300 && ql == tl && gdo_exact_match(ql, qseq, tl, ts)) {
			c->exact = 1, c->score = qlen_sum * o->a, c->n_cigar = 1;
			if (cig_used < cigar_cap) cigar[cig_used] = ql << 4;
			cig_used += 1;
		} else {
			gdo_ksw_extd2(ql, qseq, tl, ts, 5, mat, o->q, o->e, o->q2, o->e2, bw, o->zdrop, o->end_bonus, 0x08, 1, &ez,
			              cigar + cig_used, cigar_cap - cig_used);
			c->score = ez.score, c->n_cigar = ez.n_cigar;
			if (ez.n_cigar > 0) cig_used += ez.n_cigar;
		}
		free(ts);
		++n_out;
	}
	free(qs_for), free(qs_rev), free(seqs);
	if (dbg) *dbg = d;
	return (int)n_out;
}
