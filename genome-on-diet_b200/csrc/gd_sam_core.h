// gd_sam_core.h -- the short-read host stage for ONE read, written so that the same source runs
// as a CUDA device function (one thread per read, gd_sam.cu) and as plain host code (tests/emu/emu_sam.cpp, which
// compares it with the threaded host implementation host/gd_sr_post.cpp on the CPU):
//
//   mm_update_extra + mm_fix_cigar   GDiet-ShortReads/align.c:93-172,259-318   (linear gap cost: the sr preset, MM_F_SR)
//   candidate filter + ordering      map.c:956-978
//   mm_set_sam_params                hit.c:494-557
//   mm_write_sam3, one segment       format.c:302-338,349-360,412-603
//
// Text goes through a Sink that either counts or writes, so the record length can be asked for without a buffer.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include "../../include/gdiet_cuda.h"

#ifdef __CUDACC__
#define GD_SAM_HD __host__ __device__ __forceinline__
#else
#define GD_SAM_HD static inline
#endif

namespace gdsam {

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3 };
enum { MAX_REGS = 32 }; // candidates per read: --AF_max_loc <= 32 (gd_sr_map_batch refuses more)

struct Reg { // the mm_reg1_t / mm_extra_t fields this path touches (minimap.h:105-131)
	int32_t rid, score, qs, qe, rs, re, rev;
	int32_t id, parent, mapq, sam_pri, mlen, blen;
	int32_t dp_score, dp_max, n_ambi;
	uint32_t *cig; // the candidate's CIGAR in the (mutable) pool: mm_fix_cigar edits it in place
	uint32_t n_cig;
};

struct Sink {
	char *p;  // NULL: count only
	size_t n;
};
GD_SAM_HD void put_c(Sink &s, char c)
{
	if (s.p) s.p[s.n] = c;
	++s.n;
}
GD_SAM_HD void put_s(Sink &s, const char *z)
{
	while (*z) put_c(s, *z++);
}
GD_SAM_HD void put_n(Sink &s, const char *z, size_t l)
{
	if (s.p)
		for (size_t i = 0; i < l; ++i) s.p[s.n + i] = z[i];
	s.n += l;
}
GD_SAM_HD void put_int(Sink &s, long v)
{
	char b[24];
	int i = 24;
	unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
	do b[--i] = (char)('0' + u % 10), u /= 10;
	while (u);
	if (v < 0) b[--i] = '-';
	put_n(s, b + i, (size_t)(24 - i));
}

GD_SAM_HD unsigned char comp_char(unsigned char c)
{ // seq_comp_table, bseq.c:11-28: IUPAC complement, case preserved, everything else unchanged
	const unsigned char up = (unsigned char)(c & ~32u), lo = (unsigned char)(c & 32u);
	if (c >= 128 || up < 'A' || up > 'Z') return c;
	unsigned char r;
	switch (up) {
	case 'A': r = 'T'; break;
	case 'C': r = 'G'; break;
	case 'G': r = 'C'; break;
	case 'T': r = 'A'; break;
	case 'U': r = 'A'; break;
	case 'B': r = 'V'; break;
	case 'V': r = 'B'; break;
	case 'D': r = 'H'; break;
	case 'H': r = 'D'; break;
	case 'K': r = 'M'; break;
	case 'M': r = 'K'; break;
	case 'R': r = 'Y'; break;
	case 'Y': r = 'R'; break;
	default: return c;
	}
	return (unsigned char)(r | lo);
}

GD_SAM_HD void put_seq(Sink &s, const char *seq, int l, int rev, int comp)
{ // sam_write_sq, format.c:349-360
	if (s.p) {
		char *d = s.p + s.n;
		if (!rev)
			for (int i = 0; i < l; ++i) d[i] = seq[i];
		else if (comp)
			for (int i = 0; i < l; ++i) d[i] = (char)comp_char((unsigned char)seq[l - 1 - i]);
		else
			for (int i = 0; i < l; ++i) d[i] = seq[l - 1 - i];
	}
	s.n += (size_t)l;
}

// printf("%.4f", x) for 0 <= x <= 1, exactly: the double is m * 2^e; m * 10^4 is shifted right by -e bits with
// round-half-to-even on the exact remainder -- what glibc's correctly rounded conversion prints.
GD_SAM_HD void put_fixed4(Sink &s, double x)
{
	union {
		double d;
		uint64_t u;
	} z;
	z.d = x;
	const int be = (int)((z.u >> 52) & 0x7ff);
	uint64_t m = z.u & ((1ull << 52) - 1);
	int e;
	if (be == 0) e = -1074; // subnormal
	else m |= 1ull << 52, e = be - 1075;
	unsigned __int128 t = (unsigned __int128)m * 10000u;
	uint64_t q;
	if (e >= 0) q = (uint64_t)(t << e); // (x <= 1 has e < 0; kept for completeness)
	else {
		const int sh = -e;
		if (sh >= 120) q = 0; // below 10^4 * 2^-120: rounds to zero
		else {
			const unsigned __int128 one = 1;
			const unsigned __int128 rem = t & ((one << sh) - 1), half = one << (sh - 1);
			q = (uint64_t)(t >> sh);
			if (rem > half || (rem == half && (q & 1))) ++q;
		}
	}
	put_int(s, (long)(q / 10000));
	put_c(s, '.');
	const uint32_t f = (uint32_t)(q % 10000);
	put_c(s, (char)('0' + f / 1000)), put_c(s, (char)('0' + f / 100 % 10)), put_c(s, (char)('0' + f / 10 % 10)), put_c(s, (char)('0' + f % 10));
}

// align.c:93-172
GD_SAM_HD void fix_cigar(Reg &r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
{
	uint32_t *c = r.cig;
	int32_t toff = 0, qoff = 0;
	bool shrink = false;
	*qshift = *tshift = 0;
	if (r.n_cig <= 1) return;
	const uint32_t n = r.n_cig;
	for (uint32_t k = 0; k < n; ++k) { // indel left alignment
		const uint32_t op = c[k] & 0xf, len = c[k] >> 4;
		if (len == 0) shrink = true;
		if (op == OP_M) toff += len, qoff += len;
		else if (op == OP_I || op == OP_D) {
			if (k > 0 && k < n - 1 && (c[k - 1] & 0xf) == 0 && (c[k + 1] & 0xf) == 0) {
				const int prev_len = (int)(c[k - 1] >> 4);
				const uint8_t *sq = op == OP_I ? qseq : tseq;
				const int o = op == OP_I ? qoff : toff;
				int l = 0;
				while (l < prev_len && sq[o - 1 - l] == sq[o + (int)len - 1 - l]) ++l;
				if (l > 0) c[k - 1] -= (uint32_t)l << 4, c[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
				if (l == prev_len) shrink = true;
			}
			if (op == OP_I) qoff += len;
			else toff += len;
		} else if (op == OP_N) toff += len;
	}
	for (uint32_t k = 0; k + 2 < n; ++k) { // runs like 5I6D7I become one I and one D
		if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
			uint32_t l, sum[3] = {0, 0, 0};
			for (l = k; l < n; ++l) {
				const uint32_t op = c[l] & 0xf;
				if (op == OP_I || op == OP_D || c[l] >> 4 == 0) sum[op] += c[l] >> 4;
				else break;
			}
			if (sum[1] > 0 && sum[2] > 0 && l - k > 2) {
				c[k] = sum[1] << 4 | OP_I, c[k + 1] = sum[2] << 4 | OP_D;
				for (k += 2; k < l; ++k) c[k] &= 0xf;
				shrink = true;
			}
			k = l;
		}
	}
	if (shrink) {
		uint32_t l = 0, m = n;
		for (uint32_t k = 0; k < m; ++k) // squeeze out zero-length operations
			if (c[k] >> 4 != 0) c[l++] = c[k];
		m = l, l = 0;
		for (uint32_t k = 0; k < m; ++k) // merge equal neighbours
			if (k == m - 1 || (c[k] & 0xf) != (c[k + 1] & 0xf)) c[l++] = c[k];
			else c[k + 1] += c[k] >> 4 << 4;
		r.n_cig = l;
	}
	if ((c[0] & 0xf) == OP_I || (c[0] & 0xf) == OP_D) { // drop a leading I or D
		const int32_t l = (int32_t)(c[0] >> 4);
		if ((c[0] & 0xf) == OP_I) {
			if (r.rev) r.qe -= l;
			else r.qs += l;
			*qshift = l;
		} else r.rs += l, *tshift = l;
		++r.cig, --r.n_cig;
	}
}

// align.c:259-318 with the linear gap cost of MM_F_SR (is_eqx = 0): every term is a small integer
GD_SAM_HD void update_extra(Reg &r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int q, int e)
{
	int qshift, tshift;
	int32_t toff = 0, qoff = 0, s = 0, mx = 0;
	fix_cigar(r, qseq, tseq, &qshift, &tshift);
	qseq += qshift, tseq += tshift;
	r.blen = r.mlen = 0;
	for (uint32_t k = 0; k < r.n_cig; ++k) {
		const uint32_t op = r.cig[k] & 0xf, len = r.cig[k] >> 4;
		if (op == OP_M) {
			int n_ambi = 0, n_diff = 0;
			for (uint32_t l = 0; l < len; ++l) {
				const int cq = qseq[qoff + l], ct = tseq[toff + l];
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				// the reference indexes its 25-entry matrix with ct*5+cq even for cq == 7 (reverse-strand N): inside the
				// array that is the entry of (ct+1, 2); beyond it the read is undefined -> 0, as in gd_sr_post.cpp
				const int mi = ct * 5 + cq;
				s += mi < 25 ? mat[mi] : 0;
				if (s < 0) s = 0;
				else mx = mx > s ? mx : s;
			}
			r.blen += len - n_ambi, r.mlen += len - (n_ambi + n_diff), r.n_ambi += n_ambi;
			toff += len, qoff += len;
		} else if (op == OP_I || op == OP_D) {
			int n_ambi = 0;
			const uint8_t *sq = op == OP_I ? qseq + qoff : tseq + toff;
			for (uint32_t l = 0; l < len; ++l)
				if (sq[l] > 3) ++n_ambi;
			r.blen += len - n_ambi, r.n_ambi += n_ambi;
			s -= q + e;
			if (s < 0) s = 0;
			if (op == OP_I) qoff += len;
			else toff += len;
		} else if (op == OP_N) toff += len;
	}
	r.dp_max = mx; // (int32_t)(mx + .499) of an integer
}

GD_SAM_HD void swap_reg(Reg &a, Reg &b)
{
	const Reg t = a;
	a = b, b = t;
}

// hit.c:494-557
GD_SAM_HD void set_sam_params(Reg *regs, int n_regs, unsigned qlen, unsigned match_score, unsigned max_nb_sec)
{
	const int supp_threshold = (int)(0.8 * (double)(float)(regs[0].qe - regs[0].qs));
	unsigned nb_sec = 0;
	int dp_max2 = 0;
	regs[0].sam_pri = 1, regs[0].parent = regs[0].id;
	for (int i = 1; i < n_regs; i++) {
		regs[i].sam_pri = 0;
		if (regs[i].qe - regs[i].qs > supp_threshold) nb_sec++, regs[i].mapq = 0, regs[i].parent = regs[i].id + 1, dp_max2 = regs[i].score;
		else regs[i].mapq = 60, regs[i].parent = regs[i].id;
	}
	for (int i = 1; i < n_regs - 1; i++) { // supplementaries in front of secondaries, secondaries by score
		if (regs[i].parent != regs[i].id) {
			for (int j = i + 1; j < n_regs; j++) {
				if (regs[j].parent == regs[j].id) {
					swap_reg(regs[i], regs[j]);
					break;
				} else if (regs[i].score < regs[j].score) swap_reg(regs[i], regs[j]);
			}
		}
	}
	if (max_nb_sec < nb_sec) nb_sec = max_nb_sec;
	uint32_t mapq;
	if (nb_sec > 9) mapq = 0;
	else if (nb_sec > 6) mapq = 1;
	else if (nb_sec > 4) mapq = 2;
	else if (nb_sec == 3) mapq = 3; // (the reference tests == 3 twice: 4 secondaries fall through to 60)
	else if (nb_sec == 2) mapq = 5;
	else if (nb_sec == 1) {
		const int dp_max = regs[0].score;
		// float arithmetic, one rounding per operation as the host's SSE code does (no fused multiply-add)
#ifdef __CUDA_ARCH__
		const float identity = __fdiv_rn((float)regs[0].mlen, (float)regs[0].blen);
		const float num = __fmul_rn(__fmul_rn(54.0f, identity), (float)(dp_max - dp_max2));
		const float v = __fadd_rn(__fdiv_rn(num, (float)(qlen * match_score - (unsigned)dp_max2)), 5.0f);
#else
		const float identity = (float)regs[0].mlen / regs[0].blen;
		const float v = 54 * identity * (dp_max - dp_max2) / (qlen * match_score - dp_max2) + 5;
#endif
		mapq = (uint32_t)v;
	} else mapq = 60;
	regs[0].mapq = (int32_t)(mapq & 0xff); // 8-bit field
}

GD_SAM_HD void put_tags(Sink &s, const Reg &r)
{ // write_tags, format.c:302-338 (inv = 0, cnt = 0, subsc = 0, split = 0 on this path)
	put_s(s, "\tNM:i:"), put_int(s, r.blen - r.mlen + r.n_ambi);
	put_s(s, "\tms:i:"), put_int(s, r.dp_max);
	put_s(s, "\tAS:i:"), put_int(s, r.dp_score);
	put_s(s, "\tnn:i:"), put_int(s, r.n_ambi);
	put_s(s, r.id == r.parent ? "\ttp:A:P" : "\ttp:A:S");
	put_s(s, "\tcm:i:0\ts1:i:"), put_int(s, r.score);
	if (r.parent == r.id) put_s(s, "\ts2:i:0");
	int32_t n_gap = 0, n_gapo = 0; // mm_event_identity, align.c:949-966
	for (uint32_t k = 0; k < r.n_cig; ++k)
		if ((r.cig[k] & 0xf) == OP_I || (r.cig[k] & 0xf) == OP_D) ++n_gapo, n_gap += (int32_t)(r.cig[k] >> 4);
	const int32_t den = r.blen + r.n_ambi - n_gap + n_gapo;
	const double div = 1.0 - (double)r.mlen / den;
	put_s(s, "\tde:f:");
	if (div == 0.0) put_c(s, '0');
	else put_fixed4(s, div);
}

struct ReadIn {
	const char *name, *seq, *qual; // qual may be NULL
	int qlen, n_cand;
	const gd_sr_cand_t *cand;      // the read's candidates, in the order of gd_sr_map_batch
	uint32_t *cigar;               // CIGAR pool (cand[j].cigar_off indexes it); edited in place
	const uint8_t *qcodes, *tcodes; // code strings of candidate j at + j * stride (map.c:737-757; what the DP read)
	int64_t stride;
};
struct RefNames {
	const char *blob;     // NUL-terminated contig names back to back
	const int32_t *off;   // name of contig r at blob + off[r]
};

// One read: every SAM line it produces (or the flag-4 line).  Returns through s.n the number of bytes.
GD_SAM_HD void one_read(const ReadIn &R, const gd_sr_post_opt_t &o, const RefNames &N, Sink &out)
{
	const int qlen = R.qlen;
	int8_t mat[25]; // map.c:861-865
	const int g = o.a, bb = o.b < 0 ? o.b : -o.b;
	for (int x = 0; x < 5; ++x)
		for (int y = 0; y < 5; ++y) mat[x * 5 + y] = (int8_t)((x == 4 || y == 4) ? 0 : (x == y ? g : bb));
	Reg regs[MAX_REGS];
	int n_regs = 0;
	for (int j = 0; j < R.n_cand && j < MAX_REGS; ++j) { // map.c:932-978
		const gd_sr_cand_t &c = R.cand[j];
		Reg r;
		r.rid = c.rid, r.score = c.score, r.qs = c.qs, r.qe = c.qe, r.rs = c.rs, r.re = c.re, r.rev = c.rev;
		r.id = r.parent = r.mapq = r.sam_pri = r.mlen = r.blen = r.dp_max = r.n_ambi = 0;
		r.dp_score = c.score;
		r.cig = 0, r.n_cig = 0;
		if (c.n_cigar > 0) r.cig = R.cigar + c.cigar_off, r.n_cig = (uint32_t)c.n_cigar;
		update_extra(r, R.qcodes + (int64_t)j * R.stride, R.tcodes + (int64_t)j * R.stride, mat, o.q, o.e);
		const uint32_t clip0 = r.rev ? (uint32_t)(qlen - r.qe) : (uint32_t)r.qs, clip1 = r.rev ? (uint32_t)r.qs : (uint32_t)(qlen - r.qe);
		if (!(clip0 < (uint32_t)qlen && clip1 < (uint32_t)qlen) || r.dp_score < o.min_dp_max) continue;
		regs[n_regs++] = r;
		for (int k = n_regs - 1; k > 0 && regs[k].score > regs[k - 1].score; --k) swap_reg(regs[k], regs[k - 1]);
	}
	if (n_regs > 0) set_sam_params(regs, n_regs, (unsigned)qlen, (unsigned)o.a, o.no_print_2nd ? 0u : (unsigned)o.best_n);
	// ---- format.c:412-603 with n_seg == 1
	if (n_regs == 0) {
		if (o.sam_hit_only) return;
		put_s(out, R.name), put_s(out, "\t4\t*\t0\t0\t*\t*\t0\t0\t");
		put_n(out, R.seq, (size_t)qlen), put_c(out, '\t');
		if (R.qual) put_n(out, R.qual, (size_t)qlen);
		else put_c(out, '*');
		put_s(out, "\trl:i:0\n");
		return;
	}
	for (int j = 0; j < n_regs; ++j) {
		const Reg &r = regs[j];
		if (o.no_print_2nd && r.id != r.parent) continue; // map.c:1236
		int flag = 0;
		if (r.rev) flag |= 0x10;
		if (r.parent != r.id) flag |= 0x100;
		else if (!r.sam_pri) flag |= 0x800;
		put_s(out, R.name), put_c(out, '\t'), put_int(out, flag), put_c(out, '\t'), put_s(out, N.blob + N.off[r.rid]), put_c(out, '\t');
		put_int(out, r.rs + 1), put_c(out, '\t'), put_int(out, r.mapq), put_c(out, '\t');
		const uint32_t clip0 = r.rev ? (uint32_t)(qlen - r.qe) : (uint32_t)r.qs, clip1 = r.rev ? (uint32_t)r.qs : (uint32_t)(qlen - r.qe);
		const char clip_char = ((flag & 0x800) && !o.softclip) ? 'H' : 'S';
		if (clip0) put_int(out, clip0), put_c(out, clip_char);
		for (uint32_t k = 0; k < r.n_cig; ++k) put_int(out, r.cig[k] >> 4), put_c(out, "MIDNSHP=XB"[r.cig[k] & 0xf]);
		if (clip1) put_int(out, clip1), put_c(out, clip_char);
		put_s(out, "\t*\t0\t0\t");
		if ((flag & 0x900) == 0 || o.softclip) {
			put_seq(out, R.seq, qlen, r.rev, r.rev), put_c(out, '\t');
			if (R.qual) put_seq(out, R.qual, qlen, r.rev, 0);
			else put_c(out, '*');
		} else if (flag & 0x100) put_s(out, "*\t*");
		else {
			put_seq(out, R.seq + r.qs, r.qe - r.qs, r.rev, r.rev), put_c(out, '\t');
			if (R.qual) put_seq(out, R.qual + r.qs, r.qe - r.qs, r.rev, 0);
			else put_c(out, '*');
		}
		put_tags(out, r);
		if (r.parent == r.id && n_regs > 1) { // SA tag, format.c:563-592
			int n_sa = 0;
			for (int k = 0; k < n_regs; ++k)
				if (k != j && regs[k].parent == regs[k].id) ++n_sa;
			if (n_sa > 0) {
				put_s(out, "\tSA:Z:");
				for (int k = 0; k < n_regs; ++k) {
					const Reg &q = regs[k];
					if (k == j || q.parent != q.id) continue;
					int l_M, l_I = 0, l_D = 0;
					if (q.qe - q.qs < q.re - q.rs) l_M = q.qe - q.qs, l_D = (q.re - q.rs) - l_M;
					else l_M = q.re - q.rs, l_I = (q.qe - q.qs) - l_M;
					const int clip5 = q.rev ? qlen - q.qe : q.qs, clip3 = q.rev ? q.qs : qlen - q.qe;
					put_s(out, N.blob + N.off[q.rid]), put_c(out, ','), put_int(out, q.rs + 1), put_c(out, ','), put_c(out, "+-"[q.rev]), put_c(out, ',');
					if (clip5) put_int(out, clip5), put_c(out, 'S');
					if (l_M) put_int(out, l_M), put_c(out, 'M');
					if (l_I) put_int(out, l_I), put_c(out, 'I');
					if (l_D) put_int(out, l_D), put_c(out, 'D');
					if (clip3) put_int(out, clip3), put_c(out, 'S');
					put_c(out, ','), put_int(out, q.mapq), put_c(out, ','), put_int(out, q.blen - q.mlen + q.n_ambi), put_c(out, ';');
				}
			}
		}
		put_s(out, "\trl:i:0\n");
	}
}

// a safe upper bound of the text one read can produce (sizes the per-read slots before the text is written)
GD_SAM_HD size_t text_bound(size_t name_len, int qlen, int n_cand, int64_t n_cigar_total, size_t max_ref_name)
{
	if (n_cand == 0) return name_len + 2 * (size_t)qlen + 64;
	const size_t per = name_len + 2 * (size_t)qlen + 320 + max_ref_name;
	return (size_t)n_cand * per + 12 * (size_t)n_cigar_total + (size_t)n_cand * (size_t)n_cand * (max_ref_name + 96);
}

} // namespace gdsam
