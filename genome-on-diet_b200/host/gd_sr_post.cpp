// gd_sr_post.cpp -- host side after the DP: what GDiet-ShortReads/map.c:932-984 and
// format.c:412-603 do with the ksw_extz_t of every candidate, for a batch of reads and on all host cores:
//
//   mm_update_extra + mm_fix_cigar   align.c:93-172,259-318   indel left-alignment, blen / mlen / n_ambi / dp_max
//   candidate filter + ordering      map.c:956-978            clip / min_dp_max filter, insertion by score
//   mm_set_sam_params                hit.c:494-557            primary / secondary / supplementary, mapq
//   mm_write_sam3 (single segment)   format.c:349-410,412-603 one SAM line per reported location (or flag 4)
//
// This is HOST code of the product (plain C++, no CUDA): the reference keeps these steps on the CPU too.  Input is
// exactly what gd_sr_map_batch returns plus the read / reference text the host already holds.
#include <math.h>
#include <sys/mman.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>
#include "../../include/gdiet_cuda.h"

namespace {

enum { OP_M = 0, OP_I = 1, OP_D = 2, OP_N = 3 };

struct Reg { // the mm_reg1_t / mm_extra_t fields this path touches (minimap.h:105-131)
	int32_t rid = 0, score = 0, qs = 0, qe = 0, rs = 0, re = 0, rev = 0;
	int32_t id = 0, parent = 0, mapq = 0, sam_pri = 0, mlen = 0, blen = 0;
	int32_t dp_score = 0, dp_max = 0, n_ambi = 0;
	uint32_t *cig = nullptr; // view into the calling thread's scratch (mm_extra_t::cigar)
	uint32_t n_cig = 0;
};

struct CigView { // range-for over a Reg's CIGAR
	const uint32_t *b, *e;
	const uint32_t *begin() const { return b; }
	const uint32_t *end() const { return e; }
};
inline CigView cigar_of(const Reg &r) { return CigView{r.cig, r.cig + r.n_cig}; }

struct Nt4Table { // seq_nt4_table, sketch.c:11-18
	uint8_t t[256];
	Nt4Table()
	{
		for (int i = 0; i < 256; ++i) t[i] = 4;
		t[0] = 0, t[1] = 1, t[2] = 2, t[3] = 3;
		t['A'] = t['a'] = 0, t['C'] = t['c'] = 1, t['G'] = t['g'] = 2, t['T'] = t['t'] = t['U'] = t['u'] = 3;
	}
};
const Nt4Table g_nt4;
inline int nt4(unsigned char c) { return g_nt4.t[c]; }

struct CompTable { // seq_comp_table, bseq.c:11-28: IUPAC complement, case preserved, everything else unchanged
	unsigned char t[256];
	CompTable()
	{
		for (int i = 0; i < 256; ++i) t[i] = (unsigned char)i;
		const char *a = "ACBDKRTU", *b = "TGVHMYAA";
		for (int i = 0; a[i]; ++i) {
			t[(int)a[i]] = b[i], t[(int)a[i] + 32] = b[i] + 32;
			if (a[i] != 'T' && a[i] != 'U') t[(int)b[i]] = a[i], t[(int)b[i] + 32] = a[i] + 32;
		}
	}
};
const CompTable g_comp;

inline float mg_log2(float x)
{ // mmpriv.h:146-157
	union {
		float f;
		uint32_t i;
	} z = {x};
	float log_2 = (float)((z.i >> 23) & 255) - 128;
	z.i &= ~(255u << 23);
	z.i += 127u << 23;
	log_2 += (-0.34484843f * z.f + 2.02466578f) * z.f - 0.67487759f;
	return log_2;
}

// align.c:93-172
void fix_cigar(Reg &r, const uint8_t *qseq, const uint8_t *tseq, int *qshift, int *tshift)
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
				const uint8_t *s = op == OP_I ? qseq : tseq;
				const int o = op == OP_I ? qoff : toff;
				int l = 0;
				while (l < prev_len && s[o - 1 - l] == s[o + (int)len - 1 - l]) ++l;
				if (l > 0) c[k - 1] -= (uint32_t)l << 4, c[k + 1] += (uint32_t)l << 4, qoff -= l, toff -= l;
				if (l == prev_len) shrink = true;
			}
			if (op == OP_I) qoff += len;
			else toff += len;
		} else if (op == OP_N) toff += len;
	}
	for (uint32_t k = 0; k + 2 < n; ++k) { // runs like 5I6D7I become one I and one D
		if ((c[k] & 0xf) > 0 && (c[k] & 0xf) + (c[k + 1] & 0xf) == 3) {
			uint32_t l, s[3] = {0, 0, 0};
			for (l = k; l < n; ++l) {
				const uint32_t op = c[l] & 0xf;
				if (op == OP_I || op == OP_D || c[l] >> 4 == 0) s[op] += c[l] >> 4;
				else break;
			}
			if (s[1] > 0 && s[2] > 0 && l - k > 2) {
				c[k] = s[1] << 4 | OP_I, c[k + 1] = s[2] << 4 | OP_D;
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

// align.c:259-318 (is_eqx = 0).  Acc = double is the reference's arithmetic; with the linear gap cost of the sr
// preset every term is a small integer, for which Acc = int32_t gives the same values without the floating-point
// dependency chain.
template <class Acc> void update_extra_t(Reg &r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e, int log_gap)
{
	int qshift, tshift;
	int32_t toff = 0, qoff = 0;
	Acc s = 0, mx = 0;
	fix_cigar(r, qseq, tseq, &qshift, &tshift);
	qseq += qshift, tseq += tshift;
	r.blen = r.mlen = 0;
	for (uint32_t cg : cigar_of(r)) {
		const uint32_t op = cg & 0xf, len = cg >> 4;
		if (op == OP_M) {
			int n_ambi = 0, n_diff = 0;
			// Blocks of 8 identical unambiguous bases: with one positive match score on the whole diagonal the running sum
			// only grows inside such a block, so adding 8 matches at once leaves s and the maximum exactly as the
			// per-base steps would (integer accumulator only).
			const bool fast = sizeof(Acc) == sizeof(int32_t) && mat[0] > 0 && mat[6] == mat[0] && mat[12] == mat[0] && mat[18] == mat[0];
			uint32_t l = 0;
			while (l < len) {
				if (fast && l + 8 <= len) {
					uint64_t a, b;
					memcpy(&a, qseq + qoff + l, 8), memcpy(&b, tseq + toff + l, 8);
					if (a == b && !(a & 0xfcfcfcfcfcfcfcfcull)) {
						s += (Acc)(8 * mat[0]);
						mx = mx > s ? mx : s;
						l += 8;
						continue;
					}
				}
				const int cq = qseq[qoff + l], ct = tseq[toff + l];
				if (ct > 3 || cq > 3) ++n_ambi;
				else if (ct != cq) ++n_diff;
				// the reference indexes its 25-entry matrix with ct*5+cq even for cq == 7 (reverse-strand N);
				// inside the array that is the entry of (ct+1, 2); beyond it the read is undefined -> 0 here
				const int mi = ct * 5 + cq;
				s += mi < 25 ? mat[mi] : 0;
				if (s < 0) s = 0;
				else mx = mx > s ? mx : s;
				++l;
			}
			r.blen += len - n_ambi, r.mlen += len - (n_ambi + n_diff), r.n_ambi += n_ambi;
			toff += len, qoff += len;
		} else if (op == OP_I || op == OP_D) {
			int n_ambi = 0;
			const uint8_t *sq = op == OP_I ? qseq + qoff : tseq + toff;
			for (uint32_t l = 0; l < len; ++l)
				if (sq[l] > 3) ++n_ambi;
			r.blen += len - n_ambi, r.n_ambi += n_ambi;
			if (log_gap) s -= (Acc)(q + (double)e * mg_log2(1.0f + len));
			else s -= q + e;
			if (s < 0) s = 0;
			if (op == OP_I) qoff += len;
			else toff += len;
		} else if (op == OP_N) toff += len;
	}
	r.dp_max = (int32_t)(mx + .499);
}
void update_extra(Reg &r, const uint8_t *qseq, const uint8_t *tseq, const int8_t *mat, int8_t q, int8_t e, int log_gap)
{
	if (log_gap) update_extra_t<double>(r, qseq, tseq, mat, q, e, 1);
	else update_extra_t<int32_t>(r, qseq, tseq, mat, q, e, 0);
}

// hit.c:494-557
void set_sam_params(Reg *regs, int n_regs, unsigned qlen, unsigned match_score, unsigned max_nb_sec)
{
	const int supp_threshold = (int)(0.8 * (float)(regs[0].qe - regs[0].qs));
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
					std::swap(regs[i], regs[j]);
					break;
				} else if (regs[i].score < regs[j].score) std::swap(regs[i], regs[j]);
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
		const float identity = (float)regs[0].mlen / regs[0].blen;
		mapq = (uint32_t)(54 * identity * (dp_max - dp_max2) / (qlen * match_score - dp_max2) + 5);
	} else mapq = 60;
	regs[0].mapq = (int32_t)(mapq & 0xff); // 8-bit field
}

struct Out { // append-only text buffer; need() once per record, then unchecked pointer bumps
	char *b = nullptr;
	size_t n = 0, cap = 0;
	~Out() { free(b); }
	void need(size_t k)
	{
		if (n + k <= cap) return;
		cap = std::max(cap * 2, n + k + 4096);
		static const bool huge = !(getenv("GDIET_SAM_HUGEPAGES") && atoi(getenv("GDIET_SAM_HUGEPAGES")) == 0);
		if (huge && cap >= ((size_t)8 << 20)) {
			// Large text buffers: 2 MB aligned and marked for transparent huge pages.  Every worker thread first-touches
			// hundreds of MB of fresh memory per batch; with 4 KB pages that is one page fault per 9 records, all of them
			// taking the process's address-space lock (GDIET_SAM_HUGEPAGES=0 switches this off; tools/sam_stage_bench.py).
			cap = (cap + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
			void *nb = nullptr;
			if (posix_memalign(&nb, (size_t)2 << 20, cap) != 0) abort();
			madvise(nb, cap, MADV_HUGEPAGE);
			if (n) memcpy(nb, b, n);
			free(b);
			b = (char *)nb;
			return;
		}
		b = (char *)realloc(b, cap);
		if (!b) abort();
	}
	size_t size() const { return n; }
	const char *data() const { return b; }
	Out &operator+=(char c) { return b[n++] = c, *this; }
	Out &operator+=(const char *s)
	{
		while (*s) b[n++] = *s++;
		return *this;
	}
	void append(const char *s, size_t l) { memcpy(b + n, s, l), n += l; }
};

inline void put_int(Out &s, long v)
{
	char b[24];
	int i = 24;
	unsigned long u = v < 0 ? 0ul - (unsigned long)v : (unsigned long)v;
	do b[--i] = (char)('0' + u % 10), u /= 10;
	while (u);
	if (v < 0) b[--i] = '-';
	s.append(b + i, (size_t)(24 - i));
}

void put_seq(Out &s, const char *seq, int l, int rev, int comp)
{ // sam_write_sq, format.c:349-360
	if (!rev) {
		s.append(seq, (size_t)l);
		return;
	}
	char *d = s.b + s.n;
	if (comp)
		for (int i = 0; i < l; ++i) {
			const int c = seq[l - 1 - i];
			d[i] = (char)((c >= 0 && c < 128) ? g_comp.t[c] : c);
		}
	else
		for (int i = 0; i < l; ++i) d[i] = seq[l - 1 - i];
	s.n += (size_t)l;
}

void put_tags(Out &s, const Reg &r)
{ // write_tags, format.c:302-338 (inv = 0, cnt = 0, subsc = 0, split = 0 on this path)
	s += "\tNM:i:", put_int(s, r.blen - r.mlen + r.n_ambi);
	s += "\tms:i:", put_int(s, r.dp_max);
	s += "\tAS:i:", put_int(s, r.dp_score);
	s += "\tnn:i:", put_int(s, r.n_ambi);
	s += r.id == r.parent ? "\ttp:A:P" : "\ttp:A:S";
	s += "\tcm:i:0\ts1:i:", put_int(s, r.score);
	if (r.parent == r.id) s += "\ts2:i:0";
	int32_t n_gap = 0, n_gapo = 0; // mm_event_identity, align.c:949-966
	for (uint32_t cg : cigar_of(r))
		if ((cg & 0xf) == OP_I || (cg & 0xf) == OP_D) ++n_gapo, n_gap += (int32_t)(cg >> 4);
	// "%.4f" of 1 - mlen/den: the value depends on two small integers, and neighbouring reads repeat them, so the
	// printf result is kept in a small per-thread direct-mapped table
	const int32_t den = r.blen + r.n_ambi - n_gap + n_gapo;
	struct DeMemo {
		int32_t mlen, den;
		char txt[16];
	};
	static thread_local DeMemo memo[256];
	DeMemo &m = memo[((uint32_t)r.mlen * 31u + (uint32_t)den) & 255u];
	if (m.mlen != r.mlen || m.den != den || m.txt[0] == 0) {
		const double div = 1.0 - (double)r.mlen / den;
		m.mlen = r.mlen, m.den = den;
		if (div == 0.0) m.txt[0] = '0', m.txt[1] = 0;
		else snprintf(m.txt, 16, "%.4f", div);
	}
	s += "\tde:f:", s += (const char *)m.txt;
}

struct Ctx {
	int n;
	const char *const *names;
	const int64_t *off;
	const int32_t *len;
	const char *seq, *qual;
	const int64_t *cand_off;
	const gd_sr_cand_t *cand;
	const uint32_t *cigar;
	int n_seq;
	const char *const *seq_names;
	const int64_t *ref_off;
	const int32_t *ref_len;
	const char *ref;
	const gd_sr_post_opt_t *o;
	bool long_read;
};

// ---- concatenate_cigars, LR/map.c:41-640: two chained candidates of one read become one record --------------------
// rs is continued by re (same strand and contig).  Where the two alignments overlap -- on the query, else on the target --
// every cut position is scored from the two CIGARs (prefix score of the first + remaining score of the second) and the
// best cut joins them, with one I / D for whatever the other coordinate still lacks.  The reference's own evaluation is
// kept as it is: the search compares al_start[p] + al_start[p] with the initial al_start[0] + al_end[0] (map.c:267,493).
// qseq: the whole read in the strand's orientation (qs_for / qs_rev, map.c:1626-1643).  Returns 1 = not merged.
struct GapCost {
	uint32_t o1, e1, o2, e2;
	uint32_t cost(uint32_t len) const { return std::min(o1 + len * e1, o2 + len * e2); }
	void pick(uint32_t len, uint32_t &o, uint32_t &e) const
	{
		if (o1 + len * e1 < o2 + len * e2) o = o1, e = e1;
		else o = o2, e = e2;
	}
};

int concat_cigars(Reg &rs, const Reg &re, const uint8_t *qseq, int str, uint32_t read_len, const Ctx &C, uint32_t sc_mch, uint32_t sc_mis,
                  const GapCost &G, std::vector<uint32_t> &store, std::vector<uint8_t> &tseq, std::vector<int> &al_s, std::vector<int> &al_e)
{
	const uint32_t tstart = (uint32_t)rs.rs, tend = (uint32_t)re.re, tstart_junc = (uint32_t)re.rs, tend_junc = (uint32_t)rs.re;
	const uint32_t qstart = str ? read_len - rs.qe : rs.qs, qend = str ? read_len - re.qs : re.qe;
	const uint32_t qstart_junc = str ? read_len - re.qe : re.qs, qend_junc = str ? read_len - rs.qs : rs.qe;
	if (tend_junc <= tstart_junc && qend_junc <= qstart_junc) return 1; // they do not touch on either coordinate
	if (tend_junc >= tend || tstart >= tstart_junc) return 1;           // one inside the other
	if (qend_junc >= qend || qstart >= qstart_junc) return 1;
	auto fetch = [&](int rid, uint32_t st, uint32_t en) { // mm_idx_getseq (index.c:157-166) into tseq
		const uint32_t L = (uint32_t)C.ref_len[rid];
		if (en > L) en = L;
		tseq.assign((size_t)(en > st ? en - st : 0) + 8, 0);
		const unsigned char *tp = (const unsigned char *)C.ref + C.ref_off[rid] + st;
		for (uint32_t j = 0; st + j < en; ++j) tseq[j] = g_nt4.t[tp[j]];
	};
	uint32_t juncq, junct, cigar_pos;
	int score;
	const bool on_query = qend_junc > qstart_junc;
	const uint32_t juncture_len = on_query ? qend_junc - qstart_junc : tend_junc - tstart_junc;
	al_s.assign((size_t)juncture_len + 1, 0), al_e.assign((size_t)juncture_len + 1, 0);
	// ---- prefix scores of the first alignment over the overlap: the coordinate that overlaps is `a` (query positions for
	// on_query, target positions otherwise), the op that consumes only `a` is I (on_query) or D
	{
		fetch(rs.rid, tstart, tend_junc);
		const uint32_t a_junc = on_query ? qstart_junc : tstart_junc - tstart; // first overlap position in the walk's `a` units
		const uint32_t OP_A = on_query ? OP_I : OP_D, OP_B = on_query ? OP_D : OP_I;
		int al = 0;
		uint32_t toff = 0, qoff = qstart;
		for (uint32_t i = 0; i < rs.n_cig; ++i) {
			const uint32_t op = rs.cig[i] & 0xf, len = rs.cig[i] >> 4;
			const uint32_t a = on_query ? qoff : toff;
			if (op == OP_M) {
				for (uint32_t j = 0; j < len; ++j) {
					if (a + j >= a_junc) al_s[a + j - a_junc] = al;
					if (qseq[qoff + j] == tseq[toff + j]) al += (int)sc_mch;
					else al -= (int)sc_mis;
				}
				qoff += len, toff += len;
			} else if (op == OP_A) {
				uint32_t o, e;
				G.pick(len, o, e);
				if (a + len <= a_junc) al -= (int)G.cost(len);
				else if (a < a_junc) {
					al -= (int)(o + e * (a_junc - a));
					for (uint32_t j = 0; j < a + len - a_junc; ++j) al_s[j] = al, al -= (int)e;
				} else {
					al_s[a - a_junc] = al;
					al -= (int)(o + e);
					for (uint32_t j = 1; j < len; ++j) al_s[a + j - a_junc] = al, al -= (int)e;
				}
				if (on_query) qoff += len;
				else toff += len;
			} else if (op == OP_B) {
				al -= (int)G.cost(len);
				if (on_query) toff += len;
				else qoff += len;
			} else if (op == OP_N) toff += len;
		}
	}
	// ---- remaining scores of the second alignment over the overlap
	{
		fetch(re.rid, tstart_junc, tend);
		const uint32_t OP_A = on_query ? OP_I : OP_D, OP_B = on_query ? OP_D : OP_I;
		const uint32_t a_end = on_query ? qend_junc : tend_junc - tstart_junc; // end of the overlap in the walk's `a` units
		const uint32_t a_base = on_query ? qstart_junc : 0;
		int al = on_query ? re.score : 0;
		uint32_t toff = 0, qoff = qstart_junc;
		for (uint32_t i = 0; i < re.n_cig && (on_query ? qoff : toff) <= a_end; ++i) {
			const uint32_t op = re.cig[i] & 0xf, len = re.cig[i] >> 4;
			const uint32_t a = on_query ? qoff : toff;
			if (op == OP_M) {
				for (uint32_t j = 0; j < len && a + j < a_end; ++j) {
					if (qseq[qoff + j] == tseq[toff + j]) al -= (int)sc_mch;
					else al += (int)sc_mis;
					al_e[a + j - a_base] = al;
				}
				qoff += len, toff += len;
			} else if (op == OP_A) {
				uint32_t o, e;
				G.pick(len, o, e);
				al += (int)o;
				for (uint32_t j = 0; j < len && a + j < a_end; ++j) al += (int)e, al_e[a + j - a_base] = al;
				if (on_query) qoff += len;
				else toff += len;
			} else if (op == OP_B) {
				al += (int)G.cost(len);
				if (on_query) toff += len;
				else qoff += len;
			} else if (op == OP_N) toff += len;
		}
	}
	// ---- the cut (LR/map.c:262-274 / :488-499, evaluated exactly as written there)
	{
		int max_score = al_s[0] + al_e[0];
		uint32_t best = 0;
		for (uint32_t p = 1; p < juncture_len; ++p) {
			const int total = al_s[p] + al_s[p];
			if (total > max_score) max_score = total, best = p;
		}
		score = max_score;
		if (on_query) juncq = best + qstart_junc, junct = 0;
		else junct = best + tstart_junc, juncq = 0;
	}
	// ---- the first CIGAR up to the cut (LR/map.c:284-320 / :513-547); the result may be longer than either input
	const size_t base = store.size();
	store.resize(base + rs.n_cig + re.n_cig + 2);
	uint32_t *out = store.data() + base;
	memcpy(out, rs.cig, (size_t)rs.n_cig * 4);
	{
		uint32_t qoff = qstart, toffs = (uint32_t)rs.rs, i;
		for (i = 0; i < rs.n_cig; ++i) {
			const uint32_t op = out[i] & 0xf, len = out[i] >> 4;
			if (op == OP_M) {
				const bool hit = on_query ? qoff + len >= juncq : toffs + len >= junct;
				if (hit) {
					const uint32_t new_len = on_query ? juncq - qoff : junct - toffs;
					out[i] = OP_M | (new_len << 4);
					qoff += new_len, toffs += new_len;
					++i;
					break;
				}
				qoff += len, toffs += len;
			} else if (op == OP_I) {
				if (on_query && qoff + len >= juncq) { // move the cut in front of the insertion
					juncq = qoff;
					break;
				}
				qoff += len;
			} else if (op == OP_D) {
				if (!on_query && toffs + len >= junct) {
					junct = toffs;
					break;
				}
				toffs += len;
			} else if (op == OP_N) toffs += len;
		}
		if (on_query) junct = toffs;
		else juncq = qoff;
		cigar_pos = i;
	}
	// ---- the second CIGAR from the cut on, with the joining I / D (LR/map.c:551-616)
	{
		uint32_t toffe = (uint32_t)re.rs, qoffend = qstart_junc, i = cigar_pos;
		bool crossed = false;
		for (uint32_t j = 0; j < re.n_cig; ++j) {
			const uint32_t op = re.cig[j] & 0xf, len = re.cig[j] >> 4;
			if (op == OP_M || op == OP_I || op == OP_D || op == OP_N) {
				if (crossed) out[i++] = re.cig[j];
				if (op == OP_M) qoffend += len, toffe += len;
				else if (op == OP_I) qoffend += len;
				else toffe += len;
			}
			if (!crossed && qoffend >= juncq && toffe >= junct) {
				const uint32_t tar_len = toffe - junct, que_len = qoffend - juncq;
				if (que_len > tar_len) {
					const uint32_t l = que_len - tar_len;
					score -= (int)G.cost(l);
					out[i++] = OP_I | (l << 4);
					if (tar_len != 0) out[i++] = OP_M | (tar_len << 4);
				} else if (que_len < tar_len) {
					const uint32_t l = tar_len - que_len;
					score -= (int)G.cost(l);
					out[i++] = OP_D | (l << 4);
					if (que_len != 0) out[i++] = OP_M | (que_len << 4);
				} else out[i++] = OP_M | (tar_len << 4);
				crossed = true;
			}
		}
		rs.cig = out, rs.n_cig = i;
	}
	rs.dp_score = score, rs.score = score;
	if (str) rs.qs = re.qs;
	else rs.qe = re.qe;
	rs.re = re.re;
	return 0;
}

struct Scratch { // per-thread, reused across reads
	std::vector<Reg> regs, all;
	std::vector<int> valid, next;
	std::vector<uint32_t> merged; // CIGARs produced by concat_cigars
	std::vector<uint8_t> strand[2], tseq;
	std::vector<int> al_s, al_e;
	bool stitch = false; // (kept for the ABI: no read is left to the host any more)
	std::vector<uint8_t> qs, ts;
	std::vector<uint32_t> cig;
};

void one_read(const Ctx &C, int i, Out &out, Scratch &T)
{
	const gd_sr_post_opt_t &o = *C.o;
	const bool lr = C.long_read;
	T.stitch = false;
	const int qlen = C.len[i];
	const char *rd = C.seq + C.off[i], *ql = C.qual ? C.qual + C.off[i] : nullptr;
	const char *name = C.names[i];
	int8_t mat[25]; // map.c:861-865
	const int g = o.a, bb = o.b < 0 ? o.b : -o.b;
	for (int x = 0; x < 5; ++x)
		for (int y = 0; y < 5; ++y) mat[x * 5 + y] = (int8_t)((x == 4 || y == 4) ? 0 : (x == y ? g : bb));
	std::vector<Reg> &regs = T.regs;
	std::vector<uint8_t> &qs = T.qs, &ts = T.ts;
	regs.clear();
	T.all.clear(), T.valid.clear();
	size_t ncig = 0;
	for (int64_t ci = C.cand_off[i]; ci < C.cand_off[i + 1]; ++ci) ncig += C.cand[ci].n_cigar > 0 ? (size_t)C.cand[ci].n_cigar : 0;
	if (T.cig.size() < ncig + 1) T.cig.resize(2 * ncig + 64);
	uint32_t *cig_next = T.cig.data();
	for (int64_t ci = C.cand_off[i]; ci < C.cand_off[i + 1]; ++ci) { // map.c:932-978
		const gd_sr_cand_t &c = C.cand[ci];
		Reg r = Reg();
		if (lr && c.score == -0x40000000) { // LR/map.c:1812: the candidate is dropped (valid = 0)
			T.all.push_back(r), T.valid.push_back(0);
			continue;
		}
		r.rid = c.rid, r.score = c.score, r.qs = c.qs, r.qe = c.qe, r.rs = c.rs, r.re = c.re, r.rev = c.rev;
		r.dp_score = c.score;
		if (c.n_cigar > 0) {
			memcpy(cig_next, C.cigar + c.cigar_off, (size_t)c.n_cigar * 4);
			r.cig = cig_next, r.n_cig = (uint32_t)c.n_cigar, cig_next += c.n_cigar;
		}
		const int n = c.qe - c.qs, tl = c.re - c.rs;
		qs.resize((size_t)n + 1), ts.resize((size_t)tl + 1);
		{ // map.c:737-757
			uint8_t *qd = qs.data(), *td = ts.data();
			const unsigned char *src = (const unsigned char *)rd, *tp = (const unsigned char *)C.ref + C.ref_off[c.rid] + c.rs;
			const int t_in = std::min(tl, C.ref_len ? C.ref_len[c.rid] - c.rs : tl); // mm_idx_getseq clips at the contig end
			if (c.rev)
				for (int j = 0; j < n; ++j) qd[j] = g_nt4.t[src[c.qe - 1 - j]] ^ 3;
			else
				for (int j = 0; j < n; ++j) qd[j] = g_nt4.t[src[c.qs + j]];
			for (int j = 0; j < t_in; ++j) td[j] = g_nt4.t[tp[j]];
			for (int j = std::max(t_in, 0); j < tl; ++j) td[j] = 0;
		}
		update_extra(r, qs.data(), ts.data(), mat, (int8_t)o.q, (int8_t)o.e, !o.is_sr);
		const uint32_t clip0 = r.rev ? (uint32_t)(qlen - r.qe) : (uint32_t)r.qs, clip1 = r.rev ? (uint32_t)r.qs : (uint32_t)(qlen - r.qe);
		if (lr) { // LR/map.c:1843-1851: only the clip test here; chaining and the score filter follow
			T.all.push_back(r), T.valid.push_back((clip0 < (uint32_t)qlen && clip1 < (uint32_t)qlen) ? 1 : 0);
			continue;
		}
		if (!(clip0 < (uint32_t)qlen && clip1 < (uint32_t)qlen) || r.dp_score < o.min_dp_max) continue;
		regs.push_back(r);
		for (size_t k = regs.size() - 1; k > 0 && regs[k].score > regs[k - 1].score; --k) std::swap(regs[k], regs[k - 1]);
	}
	if (lr) {
		const int64_t c0 = C.cand_off[i];
		const int nc = (int)T.all.size();
		// LR/map.c:1855-1874: a valid candidate continued by a valid one absorbs it (concatenate_cigars), repeatedly
		T.next.assign((size_t)nc, -1);
		bool any_chain = false;
		for (int j = 0; j < nc; ++j) {
			const int nx = C.cand[c0 + j].reserved[0];
			T.next[j] = (nx >= 0 && nx < nc) ? nx : -1;
			any_chain |= T.valid[j] && T.next[j] >= 0 && T.valid[T.next[j]];
		}
		if (any_chain) {
			for (int st = 0; st < 2; ++st) T.strand[st].resize((size_t)qlen + 8);
			for (int j = 0; j < qlen; ++j) { // qs_for / qs_rev, LR/map.c:1626-1643
				const uint8_t c = g_nt4.t[(unsigned char)rd[j]];
				T.strand[0][j] = c, T.strand[1][qlen - 1 - j] = c ^ 3;
			}
			size_t room = 0;
			for (int j = 0; j < nc; ++j) room += T.all[j].n_cig + 2;
			T.merged.clear();
			T.merged.reserve(room * (size_t)nc + 64); // views into it must stay valid: never reallocated below
			const GapCost G = {(uint32_t)o.q, (uint32_t)o.e, (uint32_t)o.q2, (uint32_t)o.e2};
			for (int j = 0; j < nc; ++j) {
				while (T.valid[j] && T.next[j] >= 0 && T.valid[T.next[j]]) {
					const int nx = T.next[j];
					if (concat_cigars(T.all[j], T.all[nx], T.strand[T.all[j].rev ? 1 : 0].data(), T.all[j].rev, (uint32_t)qlen, C, (uint32_t)o.a,
					                  (uint32_t)o.b, G, T.merged, T.tseq, T.al_s, T.al_e) == 0)
						T.valid[nx] = 0, T.next[j] = T.next[nx];
					else T.next[j] = -1;
				}
			}
		}
		for (int j = 0; j < nc; ++j) { // LR/map.c:1876-1910
			if (!T.valid[j] || T.all[j].dp_score < o.min_dp_max) continue;
			regs.push_back(T.all[j]);
			for (size_t k = regs.size() - 1; k > 0 && regs[k].score > regs[k - 1].score; --k) std::swap(regs[k], regs[k - 1]);
		}
	}
	if (!regs.empty()) set_sam_params(regs.data(), (int)regs.size(), (unsigned)qlen, (unsigned)o.a, o.no_print_2nd ? 0u : (unsigned)o.best_n);
	// ---- format.c:412-603 with n_seg == 1
	const size_t name_len = strlen(name);
	if (regs.empty()) {
		if (o.sam_hit_only) return;
		out.need(name_len + 2 * (size_t)qlen + 64);
		out += name, out += "\t4\t*\t0\t0\t*\t*\t0\t0\t";
		out.append(rd, (size_t)qlen), out += '\t';
		if (ql) out.append(ql, (size_t)qlen);
		else out += '*';
		out += "\trl:i:0\n";
		return;
	}
	for (size_t j = 0; j < regs.size(); ++j) {
		const Reg &r = regs[j];
		if (o.no_print_2nd && r.id != r.parent) continue; // map.c:1236
		out.need(name_len + 2 * (size_t)qlen + 12 * (size_t)r.n_cig + 512 + strlen(C.seq_names[r.rid]));
		int flag = 0;
		if (r.rev) flag |= 0x10;
		if (r.parent != r.id) flag |= 0x100;
		else if (!r.sam_pri) flag |= 0x800;
		out += name, out += '\t', put_int(out, flag), out += '\t', out += C.seq_names[r.rid], out += '\t', put_int(out, r.rs + 1);
		out += '\t', put_int(out, r.mapq), out += '\t';
		const uint32_t clip0 = r.rev ? (uint32_t)(qlen - r.qe) : (uint32_t)r.qs, clip1 = r.rev ? (uint32_t)r.qs : (uint32_t)(qlen - r.qe);
		const char clip_char = ((flag & 0x800) && !o.softclip) ? 'H' : 'S';
		if (clip0) put_int(out, clip0), out += clip_char;
		for (uint32_t cg : cigar_of(r)) put_int(out, cg >> 4), out += "MIDNSHP=XB"[cg & 0xf];
		if (clip1) put_int(out, clip1), out += clip_char;
		out += "\t*\t0\t0\t";
		if ((flag & 0x900) == 0 || o.softclip) {
			put_seq(out, rd, qlen, r.rev, r.rev), out += '\t';
			if (ql) put_seq(out, ql, qlen, r.rev, 0);
			else out += '*';
		} else if (flag & 0x100) out += "*\t*";
		else {
			put_seq(out, rd + r.qs, r.qe - r.qs, r.rev, r.rev), out += '\t';
			if (ql) put_seq(out, ql + r.qs, r.qe - r.qs, r.rev, 0);
			else out += '*';
		}
		put_tags(out, r);
		if (r.parent == r.id && regs.size() > 1) { // SA tag, format.c:563-592
			int n_sa = 0;
			for (size_t k = 0; k < regs.size(); ++k)
				if (k != j && regs[k].parent == regs[k].id) ++n_sa;
			if (n_sa > 0) {
				out += "\tSA:Z:";
				for (size_t k = 0; k < regs.size(); ++k) {
					const Reg &q = regs[k];
					if (k == j || q.parent != q.id) continue;
					out.need(strlen(C.seq_names[q.rid]) + 160);
					int l_M, l_I = 0, l_D = 0;
					if (q.qe - q.qs < q.re - q.rs) l_M = q.qe - q.qs, l_D = (q.re - q.rs) - l_M;
					else l_M = q.re - q.rs, l_I = (q.qe - q.qs) - l_M;
					const int clip5 = q.rev ? qlen - q.qe : q.qs, clip3 = q.rev ? q.qs : qlen - q.qe;
					out += C.seq_names[q.rid], out += ',', put_int(out, q.rs + 1), out += ',', out += "+-"[q.rev], out += ',';
					if (clip5) put_int(out, clip5), out += 'S';
					if (l_M) put_int(out, l_M), out += 'M';
					if (l_I) put_int(out, l_I), out += 'I';
					if (l_D) put_int(out, l_D), out += 'D';
					if (clip3) put_int(out, clip3), out += 'S';
					out += ',', put_int(out, q.mapq), out += ',', put_int(out, q.blen - q.mlen + q.n_ambi), out += ';';
				}
			}
		}
		out.need(16);
		out += "\trl:i:0\n";
	}
}

} // namespace

static void put_int(std::string &s, long v) { s += std::to_string(v); }

extern "C" int gd_sam_header(int n_seq, const char *const *seq_names, const int32_t *ref_len, char **sam, size_t *sam_len)
{ // mm_write_sam_hdr, format.c:128-137 (the @PG line carries the command line and is the caller's)
	if (n_seq < 0 || !sam || !sam_len) return GD_ERR_ARG;
	std::string s;
	for (int i = 0; i < n_seq; ++i) s += "@SQ\tSN:", s += seq_names[i], s += "\tLN:", put_int(s, ref_len[i]), s += '\n';
	*sam = (char *)malloc(s.size() + 1);
	if (!*sam) return GD_ERR_ARG;
	memcpy(*sam, s.c_str(), s.size() + 1);
	*sam_len = s.size();
	return GD_OK;
}

extern "C" void gd_free(void *p) { free(p); }

static int sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq, const char *qual,
                     const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int n_seq, const char *const *seq_names,
                     const int64_t *ref_off, const int32_t *ref_len, const char *ref, const gd_sr_post_opt_t *opt, char **sam,
                     size_t *sam_len, bool long_read, int64_t *sam_off, uint8_t *needs_stitch, char ***parts_out = nullptr,
                     size_t **part_len = nullptr, int *n_parts = nullptr);

extern "C" int gd_sr_sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                               const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar,
                               int n_seq, const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len,
                               const char *ref, const gd_sr_post_opt_t *opt, char **sam, size_t *sam_len)
{
	if (n < 0 || !opt || !sam || !sam_len || (n > 0 && (!names || !off || !len || !seq || !cand_off || !seq_names || !ref_off || !ref)))
		return GD_ERR_ARG;
	return sam_batch(n, names, off, len, seq, qual, cand_off, cand, cigar, n_seq, seq_names, ref_off, ref_len, ref, opt, sam, sam_len, false,
	                 nullptr, nullptr);
}

extern "C" int gd_sr_sam_batch_parts(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                                     const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar,
                                     int n_seq, const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len,
                                     const char *ref, const gd_sr_post_opt_t *opt, char ***parts, size_t **part_len, int *n_parts)
{
	if (n < 0 || !opt || !parts || !part_len || !n_parts ||
	    (n > 0 && (!names || !off || !len || !seq || !cand_off || !seq_names || !ref_off || !ref)))
		return GD_ERR_ARG;
	return sam_batch(n, names, off, len, seq, qual, cand_off, cand, cigar, n_seq, seq_names, ref_off, ref_len, ref, opt, nullptr, nullptr,
	                 false, nullptr, nullptr, parts, part_len, n_parts);
}

extern "C" int gd_lr_sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                               const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar,
                               int n_seq, const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len,
                               const char *ref, const gd_sr_post_opt_t *opt, char **sam, size_t *sam_len, int64_t *sam_off,
                               uint8_t *needs_stitch)
{
	if (n < 0 || !opt || !sam || !sam_len || (n > 0 && (!names || !off || !len || !seq || !cand_off || !seq_names || !ref_off || !ref)))
		return GD_ERR_ARG;
	return sam_batch(n, names, off, len, seq, qual, cand_off, cand, cigar, n_seq, seq_names, ref_off, ref_len, ref, opt, sam, sam_len, true,
	                 sam_off, needs_stitch);
}

static int sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq, const char *qual,
                     const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int n_seq, const char *const *seq_names,
                     const int64_t *ref_off, const int32_t *ref_len, const char *ref, const gd_sr_post_opt_t *opt, char **sam,
                     size_t *sam_len, bool long_read, int64_t *sam_off, uint8_t *needs_stitch, char ***parts_out, size_t **part_len,
                     int *n_parts)
{
	Ctx C = {n, names, off, len, seq, qual, cand_off, cand, cigar, n_seq, seq_names, ref_off, ref_len, ref, opt, long_read};
	int nt = opt->n_threads > 0 ? opt->n_threads : (int)std::thread::hardware_concurrency();
	nt = std::max(1, std::min(nt, (n + 255) / 256));
	std::vector<Out> parts((size_t)nt);
	auto work = [&](int t) { // contiguous read ranges: the concatenation is in input order
		const int64_t b = (int64_t)n * t / nt, e = (int64_t)n * (t + 1) / nt;
		Scratch T;
		{ // one allocation per thread: an estimate of its text (the flag-4 line of an unmapped 50 kbp read alone is 100 kB), so that
		  // the buffer is not grown -- and copied -- again and again
			size_t est = 4096;
			for (int64_t i = b; i < e; ++i) est += 2 * (size_t)len[i] + 96 + (size_t)(cand_off[i + 1] - cand_off[i]) * ((size_t)len[i] / 8 + 384);
			parts[t].need(est);
		}
		for (int64_t i = b; i < e; ++i) {
			if (sam_off) sam_off[i] = (int64_t)parts[t].size(); // relative to the thread's part; rebased below
			one_read(C, (int)i, parts[t], T);
			if (needs_stitch) needs_stitch[i] = T.stitch ? 1 : 0;
		}
	};
	auto run = [&](auto &&f) {
		if (nt == 1) f(0);
		else {
			std::vector<std::thread> th;
			for (int t = 0; t < nt; ++t) th.emplace_back(f, t);
			for (auto &x : th) x.join();
		}
	};
	run(work);
	std::vector<size_t> at((size_t)nt + 1, 0);
	for (int t = 0; t < nt; ++t) at[t + 1] = at[t] + parts[t].size();
	const size_t total = at[nt];
	if (sam_off) {
		for (int t = 0; t < nt; ++t)
			for (int64_t i = (int64_t)n * t / nt, e = (int64_t)n * (t + 1) / nt; i < e; ++i) sam_off[i] += (int64_t)at[t];
		sam_off[n] = (int64_t)total;
	}
	if (parts_out) { // hand the per-thread pieces over as they are (input order): nothing is copied
		*parts_out = (char **)malloc(sizeof(char *) * (size_t)nt), *part_len = (size_t *)malloc(sizeof(size_t) * (size_t)nt);
		if (!*parts_out || !*part_len) return GD_ERR_ARG;
		for (int t = 0; t < nt; ++t) (*parts_out)[t] = parts[t].b, (*part_len)[t] = parts[t].n, parts[t].b = nullptr;
		*n_parts = nt;
		if (sam_len) *sam_len = total;
		return GD_OK;
	}
	char *buf = (char *)malloc(total + 1);
	if (!buf) return GD_ERR_ARG;
	run([&](int t) { memcpy(buf + at[t], parts[t].data(), parts[t].size()); }); // first touch + copy on every core
	buf[total] = 0;
	*sam = buf, *sam_len = total;
	return GD_OK;
}
