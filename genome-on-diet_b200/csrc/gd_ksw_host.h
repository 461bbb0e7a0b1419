// gd_ksw_host.h -- host-side preparation shared by the CUDA launcher (gd_ksw.cu) and the
// logic tests (tests/emu): scoring constants, ring sizing, arena strides.
#pragma once
#include "gd_ksw.cuh"

namespace gd {

static inline uint32_t h_pack16(int v) { return ((uint32_t)(v & 0xff) << 8) | ((uint32_t)(v & 0xff) << 24); }
static inline uint32_t h_rep4(int v) { return (uint32_t)(v & 0xff) * 0x01010101u; }

// Mirrors the prologue of ksw_extd2_sse (GDiet-ShortReads/ksw2_extd2_sse.c:68-105).
static inline KswConsts ksw_make_consts(int m, const int8_t *mat, int q_, int e_, int q2_, int e2_, int zdrop,
                                        int end_bonus, int flag)
{
	KswConsts C;
	int8_t q = (int8_t)q_, e = (int8_t)e_, q2 = (int8_t)q2_, e2 = (int8_t)e2_;
	C.qe_seed = (int)q + (int)e; // evaluated before the swap in the reference
	C.degenerate = 0;
	if (m <= 1) {
		C.degenerate = 1;
		m = 2;
	}
	if ((int)q2 + e2 < (int)q + e) {
		int8_t tq = q, te = e;
		q = q2, e = e2, q2 = tq, e2 = te;
	}
	int min_sc = mat ? mat[1] : 0;
	if (mat)
		for (int t = 1; t < m * m; ++t) min_sc = min_sc < mat[t] ? min_sc : mat[t];
	if (!mat || -min_sc > 2 * ((int)q + e)) C.degenerate = 1;
	const int mch = mat ? mat[0] : 0, mis = mat ? mat[1] : 0;
	const int scn = mat ? (mat[m * m - 1] == 0 ? (int8_t)(-e2) : mat[m * m - 1]) : 0;
	const bool right = (flag & KSW_F_RIGHT) != 0;
	// tags: larger tag wins a tie inside the packed max. Left-aligned gaps: first of (H,E,F,E~,F~)
	// wins -> tags 7,6,5,4,3 (xor 7 gives 0..4); right-aligned: last wins -> tags 0..4.
	const uint32_t ts = right ? 0 : 7, ta = right ? 1 : 6, tb = right ? 2 : 5, ta2 = right ? 3 : 4, tb2 = right ? 4 : 3;
	C.TS4 = ts * 0x01010101u;
	C.TA = ta * 0x00010001u, C.TB = tb * 0x00010001u, C.TA2 = ta2 * 0x00010001u, C.TB2 = tb2 * 0x00010001u;
	C.TAGX = right ? 0u : 0x07070707u;
	C.MCH16 = h_pack16(mch);
	C.Q1 = h_pack16(q) + 0x00010001u;
	C.Q21 = h_pack16(q2) + 0x00010001u;
	C.NEGQE = h_pack16(-(int)(int8_t)(q + e));
	C.NEGQE2 = h_pack16(-(int)(int8_t)(q2 + e2));
	C.MCH4 = h_rep4(mch), C.MIS4 = h_rep4(mis), C.SCN4 = h_rep4(scn);
	{ // ring slots: A = (x << 8 | v), B = (x2 << 8 | u), C = (y << 8 | y2); the reference memsets
	  // u,v,x,y to -q-e and x2,y2 to -q2-e2 (ksw2_extd2_sse.c:111-116)
		const uint32_t i1 = (uint32_t)((-q - e) & 0xff), i2 = (uint32_t)((-q2 - e2) & 0xff);
		C.INIT_A = (i1 << 8 | i1) * 0x10001u;
		C.INIT_B = (i2 << 8 | i1) * 0x10001u;
		C.INIT_C = (i1 << 8 | i2) * 0x10001u;
	}
	C.q = q, C.e = e, C.q2 = q2, C.e2 = e2;
	int long_thres = e != e2 ? (q2 - q) / (e - e2) - 1 : 0;
	if (q2 + e2 + long_thres * e2 > q + e + long_thres * e) ++long_thres;
	C.long_thres = long_thres;
	C.long_diff = long_thres * (e - e2) - (q2 - q) - e2;
	C.zdrop = zdrop, C.end_bonus = end_bonus, C.flag = flag;
	C.force_slow_max = 0;
	return C;
}

static inline int h_ncol16(int qlen, int tlen, int w)
{
	if (w < 0) w = tlen > qlen ? tlen : qlen;
	int n = qlen < tlen ? qlen : tlen;
	n = n < w + 1 ? n : w + 1;
	return ((n + 15) / 16 + 1) * 16;
}

// Geometry of one launch, derived from upper bounds on the chunk's pairs.
struct KswGeom {
	int ring;        // columns per ring (multiple of 8)
	int group_smem;  // bytes per group: ring + staged sequences
	int t_stride, q_stride;
	int64_t p_stride;
};
// Lanes per pair: one 8-column chunk per lane and step.  Short rows keep 8 pairs per warp; pairs whose
// sequences would not fit next to the ring (staged only for G <= 8) use wide groups.
static inline int ksw_pick_group(int max_qlen, int max_tlen, int max_w, bool exact = false, int npairs = 1 << 30, int sms = 148)
{
	const int nch = h_ncol16(max_qlen, max_tlen, max_w) / 8; // chunks in the widest row
	int G = nch <= 24 ? 4 : nch <= 64 ? 8 : nch <= 160 ? 16 : 32;
	if (G <= 8 && max_qlen + max_tlen > 2048) G = 16;
	// Long pairs: the backtrack arena limits how many pairs one launch can hold (31 MB per 15 kbp HiFi pair, 141 MB
	// per 50 kbp ONT pair), so narrow gangs leave the SMs short of warps -- widen the gang, up to a block of two
	// warps per pair, until the launch has ~24 warps per SM.  (Measured: HiFi 2048 pairs 497 -> 600 GCUPS, ONT 512
	// pairs 324 -> 434 GCUPS; four warps per pair and exact mode do not gain.)
	if (!exact) {
		while (G >= 16 && G < 64 && (int64_t)npairs * G / 32 < (int64_t)sms * 24) G <<= 1;
		// A handful of very long pairs (fewer blocks than twice the SMs): two warps per pair leave half of every SM's four
		// schedulers idle, so the gang grows to four warps.  (ONT, 72 pairs of 50 kbp: 94 -> 109 GCUPS; with hundreds of pairs four
		// warps per pair do not gain, and eight warps per pair -- one step per 1300-wide row instead of two -- gained only 2 %:
		// the per-row barriers and bookkeeping, not the number of steps, bound so few pairs.)
		if (G == 64 && (int64_t)npairs < (int64_t)sms * 2) G = 128;
	}
	return G;
}
static inline KswGeom ksw_geometry(int max_qlen, int max_tlen, int max_w, bool exact, bool with_p, int G)
{
	KswGeom g;
	if (max_qlen < 1) max_qlen = 1;
	if (max_tlen < 1) max_tlen = 1;
	const int T16 = (max_tlen + 15) / 16 * 16;
	const int ncol16 = h_ncol16(max_qlen, max_tlen, max_w);
	// live columns of a row: [st-1, en+16] (left boundary slot .. the block the score row may run into)
	g.ring = ncol16 + 24 < T16 + 8 ? ncol16 + 24 : T16 + 8;
	g.t_stride = T16;
	g.q_stride = (max_qlen + 15) / 16 * 16 + 64;
	g.group_smem = ksw_group_smem_bytes(g.ring, exact, G <= 8 ? g.q_stride : 0); // the target travels in the ring records
	int pitch = ncol16; // the kernel's row pitch (ksw_ncol16)
#if GD_KSW_P32
	pitch = (pitch + 31) & ~31;
#endif
	g.p_stride = with_p ? (int64_t)(max_qlen + max_tlen - 1) * pitch : 0;
	return g;
}

} // namespace gd
