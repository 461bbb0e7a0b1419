// gd_ksw.cuh -- batched ksw2 dual-affine banded extension DP for sm_100a.
//
// Replaces ksw_extd2_sse / ksw_extd2_avx512 (GDiet-ShortReads/ksw2_extd2_sse.c:27-401,
// GDiet-ShortReads/ksw2_extd2_avx.c:72-915) for whole batches of (query,target) pairs.
//
// Mapping.  A *group* of G lanes (G = 4..32, a power of two) owns one pair at a time and walks
// its anti-diagonals r = 0..qlen+tlen-2.  One *step* of a group updates G/4 consecutive
// 16-cell blocks of the row; each lane owns 4 consecutive target columns of its block.  The
// reference's persistent per-column int8 state (u,v,x,y,x2,y2) and its score row s live in
// shared memory in a ring of R columns; the 16-cell range rounding, the stale score cells and
// the boundary injections of the reference are reproduced exactly because
// cells outside the true band leak into later rows.
//
// Arithmetic.  int8 with wrap-around is carried in the HIGH byte of each half of a 32-bit
// register (two cells per register), so VIADD.16x2 / VIMNMX.S16x2 / VIMNMX3.S16x2 give exact
// int8 modular add and signed compare on two cells per instruction.  The LOW byte of each half
// carries a small tag (which of H,E,F,E~,F~ a value belongs to); the 5-way max therefore
// returns the arg-max in its low byte for free, with exactly the reference's tie order
// (strict '>' for left-aligned gaps, '>=' with KSW_EZ_RIGHT).
//
// Backtrack.  One byte per cell, same bit layout as ksw2.h:127-130, written straight to HBM as
// one 32-bit store per lane-step (a group step writes 4*G contiguous bytes); the CIGAR is
// produced by gd_ksw_traceback_kernel, one thread per pair.
#pragma once
#include "gd_common.cuh"

namespace gd {

enum {
	KSW_F_SCORE_ONLY = 0x01,
	KSW_F_RIGHT = 0x02,
	KSW_F_GENERIC_SC = 0x04,
	KSW_F_APPROX_MAX = 0x08,
	KSW_F_APPROX_DROP = 0x10,
	KSW_F_EXTZ_ONLY = 0x40,
	KSW_F_REV_CIGAR = 0x80
};
#define GD_KSW_NEG_INF (-0x40000000)

// Per-pair result record (device + host). First 11 ints mirror ksw_extz_t (ksw2.h:31-40).
struct KswResult {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
	int32_t tb_i, tb_j; // traceback start cell (target, query) or -1: filled by the DP kernel
	int32_t rows_done;  // anti-diagonals executed (for cell accounting)
	int32_t pad0, pad1;
};

// Batch-uniform constants, precomputed on the host (gd_ksw_make_consts).
struct KswConsts {
	uint32_t MCH16, Q1, Q21, NEGQE, NEGQE2;
	uint32_t TS4;                // tag of the score term, replicated in all 4 bytes
	uint32_t TA, TB, TA2, TB2;   // 16x2 tags (low byte of each half)
	uint32_t TAGX;               // xor applied to the extracted tags -> 0..4 as in ksw2.h
	uint32_t MCH4, MIS4, SCN4;   // byte-replicated match / mismatch / ambiguous scores
	uint32_t INIT_U, INIT_X, INIT_Y, INIT_X2, INIT_Y2; // 16x2 initial column state
	int32_t q, e, q2, e2;        // after ordering the two pieces (ksw2_extd2_sse.c:78)
	int32_t qe_seed;             // q+e BEFORE ordering: seeds H (ksw2_extd2_sse.c:68,358,382)
	int32_t long_thres, long_diff;
	int32_t zdrop, end_bonus, flag;
	int32_t degenerate;          // m<=1 or -min(mat) > 2(q+e): every call returns the reset ez
};

// One launch works on pairs [base, base+n) of the caller's batch ("chunk"); the packed sequence
// arenas and the backtrack arena are indexed by the chunk-local pair number with uniform strides.
struct KswBatch {
	int32_t n, base;
	const int32_t *qlen, *tlen, *w; // per pair (global index); w may be NULL -> w_all
	int32_t w_all;
	const uint8_t *tpk;  // target codes, zero padded; pair i at tpk + i*t_stride (16-byte aligned)
	const uint8_t *qpk;  // reversed query, N(4)->8; pair i: qpk + i*q_stride + 16 is element 0,
	                     // 16 zero bytes in front and >= 32 behind
	int32_t t_stride, q_stride;
	uint8_t *p;          // backtrack arena, pair i at p + i*p_stride
	int64_t p_stride;
	KswResult *res;      // global index
	int32_t *ticket;     // dynamic pair dispenser (chunk-local)
	int32_t ring;        // R: columns per ring (multiple of 16)
	int32_t group_smem;  // bytes of shared memory per group
};

GD_DEV uint32_t pack16(int v) { return ((uint32_t)(v & 0xff) << 8) | ((uint32_t)(v & 0xff) << 24); }
GD_DEV int hi8(uint32_t h16) { return (int)(int8_t)(h16 >> 8); } // value of a 16-bit cell image

GD_DEV int gap_delta(int r, const KswConsts &C)
{ // first-row / first-column difference (ksw2_extd2_sse.c:158,162); int8 wrap applied by the caller
	return r == 0 ? -C.q - C.e : r < C.long_thres ? -C.e : r == C.long_thres ? C.long_diff : -C.e2;
}

// Two cells of the recurrence (ksw2_extd2_sse.c:38-66,228-274). All operands are 16x2 images.
// In: s (tag TS), xt1/x2t1 (left neighbours, tags TA/TA2), vt1, ut (no tag), y/y2 (tags TB/TB2).
// Out: new u,v,x,y,x2,y2 and the word zt whose low bytes hold the arg-max tag, plus the four
// "positive" words whose bit 15/31 is the continuation flag.
template <bool RIGHT>
GD_DEV void cell2(const KswConsts &C, uint32_t s, uint32_t xt1, uint32_t vt1, uint32_t x2t1, uint32_t ut, uint32_t &y,
                  uint32_t &y2, uint32_t &u_new, uint32_t &v_new, uint32_t &x_new, uint32_t &x2_new, uint32_t &zt,
                  uint32_t &fa, uint32_t &fb, uint32_t &fa2, uint32_t &fb2)
{
	uint32_t a = vadd2(xt1, vt1), b = vadd2(y, ut), a2 = vadd2(x2t1, vt1), b2 = vadd2(y2, ut);
	zt = vmax3(vmax3(s, a, b), a2, b2);
	uint32_t z = vmin2(zt & 0xff00ff00u, C.MCH16);
	uint32_t z1 = vadd2(z, 0x00010001u);
	u_new = vadd2(z1, ~vt1); // z - v[t-1]
	v_new = vadd2(z1, ~ut);  // z - u[t]
	uint32_t nz = ~z;
	uint32_t nzq = vadd2(nz, C.Q1), nzq2 = vadd2(nz, C.Q21); // q - z, q2 - z
	a = vadd2(a, nzq), b = vadd2(b, nzq), a2 = vadd2(a2, nzq2), b2 = vadd2(b2, nzq2);
	uint32_t ma = vmax2(a, C.TA), mb = vmax2(b, C.TB), ma2 = vmax2(a2, C.TA2), mb2 = vmax2(b2, C.TB2);
	if (!RIGHT) { // continuation iff value > 0  <=> high byte of max(value,0) >= 1
		fa = ma + 0x7f007f00u, fb = mb + 0x7f007f00u, fa2 = ma2 + 0x7f007f00u, fb2 = mb2 + 0x7f007f00u;
	} else { // continuation iff value >= 0 <=> sign bit clear
		fa = ~a, fb = ~b, fa2 = ~a2, fb2 = ~b2;
	}
	x_new = vadd2(ma, C.NEGQE), y = vadd2(mb, C.NEGQE), x2_new = vadd2(ma2, C.NEGQE2), y2 = vadd2(mb2, C.NEGQE2);
}

// 4 ksw2 backtrack bytes from the per-register words of a lane (A = cells 0,1; B = cells 2,3).
GD_DEV uint32_t make_dir4(const KswConsts &C, uint32_t ztA, uint32_t ztB, uint32_t faA, uint32_t faB, uint32_t fbA,
                          uint32_t fbB, uint32_t fa2A, uint32_t fa2B, uint32_t fb2A, uint32_t fb2B)
{
	uint32_t FA = prmt(faA, faB, 0x7531), FB = prmt(fbA, fbB, 0x7531);
	uint32_t FA2 = prmt(fa2A, fa2B, 0x7531), FB2 = prmt(fb2A, fb2B, 0x7531);
	uint32_t TG = prmt(ztA, ztB, 0x6420);
	uint32_t d = (TG & 0x07070707u) ^ C.TAGX;
	d |= (FA >> 4) & 0x08080808u;
	d |= (FB >> 3) & 0x10101010u;
	d |= (FA2 >> 2) & 0x20202020u;
	d |= (FB2 >> 1) & 0x40404040u;
	return d;
}

// Score bytes for 4 cells with the AVX-512 xor-table rule (ksw2_extd2_avx.c:187-208,312-313):
// pmat[(t ^ q') & 15] with pmat = {mch, mis x3, scN x9, 0 x3}; a byte with bit 7 set gives 0.
GD_DEV uint32_t score4(const KswConsts &C, uint32_t tc, uint32_t qc)
{
	uint32_t x = tc ^ qc, idx = x & 0x0f0f0f0fu;
	uint32_t nz = (idx + 0x7f7f7f7fu) | idx;  // bit7 of each byte: idx != 0
	uint32_t ge4 = idx + 0x7c7c7c7cu;         // bit7: idx >= 4
	uint32_t ge13 = idx + 0x73737373u;        // bit7: idx >= 13
	uint32_t m_nz = prmt(nz, 0, 0xba98), m_ge4 = prmt(ge4, 0, 0xba98);
	uint32_t m_zero = prmt(ge13 | x, 0, 0xba98); // idx >= 13 or bit 7 of the raw xor
	uint32_t sc = (C.MIS4 & m_nz) | (C.MCH4 & ~m_nz);
	sc = (C.SCN4 & m_ge4) | (sc & ~m_ge4);
	return sc & ~m_zero;
}

struct Bounds {
	int st0, en0, st, en;
};
GD_DEV bool row_bounds(int r, int qlen, int tlen, int w, Bounds &b)
{ // ksw2_extd2_sse.c:133-147
	int st0 = imax(imax(0, r - qlen + 1), (r - w + 1) >> 1);
	int en0 = imin(imin(tlen - 1, r), (r + w) >> 1);
	b.st0 = st0, b.en0 = en0, b.st = st0 & ~15, b.en = en0 | 15;
	return st0 <= en0;
}

GD_DEV int ksw_ncol16(int qlen, int tlen, int w)
{ // row pitch of the backtrack matrix (ksw2_extd2_sse.c:92-94)
	int n = imin(imin(qlen, tlen), w + 1);
	return ((n + 15) / 16 + 1) * 16;
}

// Shared-memory view of one group's column ring.
struct Ring {
	uint16_t *u, *v, *x, *y, *x2, *y2; // 16-bit cell images, R entries each
	uint8_t *s;                        // score row bytes
	int32_t *H;                        // exact-max scores (EXACT mode only)
	int R;
};
GD_DEV Ring ring_view(uint8_t *base, int R, bool exact)
{
	Ring g;
	g.R = R;
	g.u = (uint16_t *)base, g.v = g.u + R, g.x = g.v + R, g.y = g.x + R, g.x2 = g.y + R, g.y2 = g.x2 + R;
	g.s = (uint8_t *)(g.y2 + R);
	g.H = exact ? (int32_t *)(g.s + R) : 0;
	return g;
}
static inline int ksw_group_smem_bytes(int R, bool exact)
{
	int b = R * 13 + (exact ? R * 4 : 0);
	b = (b + 127) / 128 * 128 + 64; // == 64 (mod 128): two groups of a half-warp hit disjoint banks
	return b;
}

GD_DEV int wrap(int slot, int R) { return slot >= R ? slot - R : slot; }

GD_DEV uint2 lds2(const uint16_t *p) { return *(const uint2 *)p; }
GD_DEV void sts2(uint16_t *p, uint32_t a, uint32_t b)
{
	uint2 v;
	v.x = a, v.y = b;
	*(uint2 *)p = v;
}

// 64-bit key used by the exact-max row scan: larger score wins; ties resolved exactly like the
// 4-lane SSE scan of ksw2_extd2_sse.c:327-357: en0 first, then SIMD lane (t-st0)%4, then t, then the scalar tail.
GD_DEV long long hkey(int H, uint32_t prio) { return (long long)(((unsigned long long)(uint32_t)H << 32) | prio); }

template <int G, bool RIGHT, bool EXACT, bool WITH_P>
GD_DEV void ksw_group_body(const KswConsts &C, const KswBatch &B, uint8_t *smem_group, int li, uint32_t gmask)
{
	const int BPS = G / 4; // 16-cell blocks per step
	const int sub = li >> 2, c4 = (li & 3) * 4;
	Ring g = ring_view(smem_group, B.ring, EXACT);
	const int R = g.R;

	for (;;) {
		int lp = 0; // chunk-local pair number
		if (li == 0) lp = atomic_add(B.ticket, 1);
		lp = (int)shfl_idx(gmask, (uint32_t)lp, 0, G);
		if (lp >= B.n) break;
		const int pair = B.base + lp;

		const int qlen = B.qlen[pair], tlen = B.tlen[pair];
		int w = B.w ? B.w[pair] : B.w_all;
		KswResult res;
		res.max = 0, res.zdropped = 0, res.max_q = res.max_t = res.mqe_t = res.mte_q = -1;
		res.score = res.mqe = res.mte = GD_KSW_NEG_INF, res.n_cigar = 0, res.reach_end = 0;
		res.tb_i = res.tb_j = -1, res.rows_done = 0, res.pad0 = res.pad1 = 0;
		if (C.degenerate || qlen <= 0 || tlen <= 0) {
			if (li == 0) B.res[pair] = res;
			continue;
		}
		if (w < 0) w = imax(tlen, qlen);
		const int T16 = (tlen + 15) & ~15, nblk_t = T16 >> 4;
		const int ncol16 = ksw_ncol16(qlen, tlen, w);
		const uint8_t *tpk = B.tpk + (size_t)lp * B.t_stride;
		const uint8_t *qpk = B.qpk + (size_t)lp * B.q_stride + 16; // qpk[j] = mapped query[qlen-1-j]
		uint8_t *prow = WITH_P ? B.p + (size_t)lp * B.p_stride : 0;
		const int nrows = qlen + tlen - 1;

		int last_st = -1, last_en = -1;
		int st_slot = 0;   // ring slot of column `st_cur`
		int st_cur = 0;
		int init_hi = 0;   // blocks [0, init_hi) have been initialised in the ring
		int H0 = 0, H0_t = 0;
		int r, rows_exec = 0;
		sync_warp(gmask); // previous pair's readers are done with the ring

		for (r = 0; r < nrows; ++r) {
			Bounds bd;
			if (!row_bounds(r, qlen, tlen, w, bd)) {
				res.zdropped = 1;
				break;
			}
			const int st = bd.st, en = bd.en, st0 = bd.st0, en0 = bd.en0;
			if (st != st_cur) { // st only ever advances, by exactly 16
				st_slot = wrap(st_slot + (st - st_cur), R);
				st_cur = st;
			}
			// ---- phase A: ring initialisation of blocks that enter, boundary stores, score row ----
			{
				int bneed = imin((en0 + 15) >> 4, nblk_t - 1);
				const bool grow = init_hi <= bneed;
				while (init_hi <= bneed) { // lanes 0..3 of the group write one 16-cell block
					if (li < 4) {
						int slot = wrap(st_slot + (init_hi * 16 - st) + li * 4, R);
						sts2(g.u + slot, C.INIT_U, C.INIT_U);
						sts2(g.v + slot, C.INIT_U, C.INIT_U);
						sts2(g.x + slot, C.INIT_X, C.INIT_X);
						sts2(g.y + slot, C.INIT_Y, C.INIT_Y);
						sts2(g.x2 + slot, C.INIT_X2, C.INIT_X2);
						sts2(g.y2 + slot, C.INIT_Y2, C.INIT_Y2);
						*(uint32_t *)(g.s + slot) = 0;
						if (EXACT) {
							g.H[slot] = GD_KSW_NEG_INF, g.H[slot + 1] = GD_KSW_NEG_INF;
							g.H[slot + 2] = GD_KSW_NEG_INF, g.H[slot + 3] = GD_KSW_NEG_INF;
						}
					}
					++init_hi;
				}
				if (grow) sync_warp(gmask);
				if (en >= r && li == G - 1) { // ksw2_extd2_sse.c:160-163
					int slot = wrap(st_slot + (r - st), R);
					g.y[slot] = (uint16_t)(C.INIT_Y & 0xffff);
					g.y2[slot] = (uint16_t)(C.INIT_Y2 & 0xffff);
					g.u[slot] = (uint16_t)((gap_delta(r, C) & 0xff) << 8);
				}
				// score row over [st0, fe): ksw2_extd2_sse.c:166-180 extent, AVX-512 scoring rule
				const int fe = imin(st0 + (((en0 - st0) >> 4) + 1) * 16, T16);
				const int qshift = qlen - 1 - r; // qpk index of column t is qshift + t
				for (int t4 = (st0 & ~3) + li * 4; t4 < fe; t4 += G * 4) {
					uint32_t tc = *(const uint32_t *)(tpk + t4);
					int qi = qshift + t4 + 16; // >= 13; the arena has 16 zero bytes in front of qpk[0]
					const uint32_t *qw = (const uint32_t *)(qpk - 16 + (qi & ~3));
					uint32_t qc = funnel_r(qw[0], qw[1], (uint32_t)(qi & 3) * 8);
					uint32_t sc = score4(C, tc, qc);
					int slot = wrap(st_slot + (t4 - st), R);
					uint32_t *sp = (uint32_t *)(g.s + slot);
					int lo = st0 - t4, hi = fe - t4; // valid bytes: lo <= b < hi
					if (lo > 0 || hi < 4) {
						uint32_t m = 0xffffffffu;
						if (lo > 0) m &= 0xffffffffu << (8 * lo);
						if (hi < 4) m &= 0xffffffffu >> (8 * (4 - hi));
						sc = (sc & m) | (*sp & ~m);
					}
					*sp = sc;
				}
			}
			sync_warp(gmask);
			// ---- phase B: core update over blocks st/16 .. en/16 ----
			{
				// left boundary (ksw2_extd2_sse.c:149-159): becomes "previous send" of lane G-1
				uint32_t prev_xv, prev_x2;
				if (st > 0) {
					if (st - 1 >= last_st && st - 1 <= last_en) {
						int slot = st_slot == 0 ? R - 1 : st_slot - 1;
						prev_xv = (uint32_t)g.x[slot] | ((uint32_t)g.v[slot] << 16);
						prev_x2 = (uint32_t)g.x2[slot] << 16;
					} else {
						prev_xv = (C.INIT_X & 0xffffu) | (C.INIT_U << 16);
						prev_x2 = C.INIT_X2 << 16;
					}
				} else {
					prev_xv = (C.INIT_X & 0xffffu) | ((uint32_t)(gap_delta(r, C) & 0xff) << 24);
					prev_x2 = C.INIT_X2 << 16;
				}
				const int nblk = ((en - st) >> 4) + 1;
				uint8_t *pr = WITH_P ? prow + (size_t)r * ncol16 : 0;
				for (int b0 = 0; b0 < nblk; b0 += BPS) {
					const int blk = b0 + sub;
					const bool active = blk < nblk;
					const int slot = wrap(st_slot + blk * 16 + c4, R);
					uint2 U, V, X, Y, X2, Y2;
					uint32_t sw = 0;
					if (active) {
						U = lds2(g.u + slot), V = lds2(g.v + slot), X = lds2(g.x + slot), Y = lds2(g.y + slot);
						X2 = lds2(g.x2 + slot), Y2 = lds2(g.y2 + slot);
						sw = *(const uint32_t *)(g.s + slot);
					} else {
						U.x = U.y = V.x = V.y = X.x = X.y = Y.x = Y.y = X2.x = X2.y = Y2.x = Y2.y = 0;
					}
					// neighbour exchange: lane li needs x,v,x2 of column t-1 (old values)
					uint32_t send_xv = prmt(X.y, V.y, 0x7632), send_x2 = X2.y;
					uint32_t dep_xv = (li == G - 1) ? prev_xv : send_xv, dep_x2 = (li == G - 1) ? prev_x2 : send_x2;
					uint32_t rxv = shfl_idx(gmask, dep_xv, (li + G - 1) & (G - 1), G);
					uint32_t rx2 = shfl_idx(gmask, dep_x2, (li + G - 1) & (G - 1), G);
					prev_xv = send_xv, prev_x2 = send_x2;
					uint32_t xt1A = prmt(rxv, X.x, 0x5410), xt1B = prmt(X.x, X.y, 0x5432);
					uint32_t vt1A = prmt(rxv, V.x, 0x5432), vt1B = prmt(V.x, V.y, 0x5432);
					uint32_t x2t1A = prmt(rx2, X2.x, 0x5432), x2t1B = prmt(X2.x, X2.y, 0x5432);
					uint32_t sA = prmt(sw, C.TS4, 0x1404), sB = prmt(sw, C.TS4, 0x3424);
					uint32_t uA, vA, xA, x2A, ztA, faA, fbA, fa2A, fb2A;
					uint32_t uB, vB, xB, x2B, ztB, faB, fbB, fa2B, fb2B;
					cell2<RIGHT>(C, sA, xt1A, vt1A, x2t1A, U.x, Y.x, Y2.x, uA, vA, xA, x2A, ztA, faA, fbA, fa2A, fb2A);
					cell2<RIGHT>(C, sB, xt1B, vt1B, x2t1B, U.y, Y.y, Y2.y, uB, vB, xB, x2B, ztB, faB, fbB, fa2B, fb2B);
					if (active) {
						sts2(g.u + slot, uA, uB), sts2(g.v + slot, vA, vB), sts2(g.x + slot, xA, xB);
						sts2(g.y + slot, Y.x, Y.y), sts2(g.x2 + slot, x2A, x2B), sts2(g.y2 + slot, Y2.x, Y2.y);
						if (WITH_P)
							*(uint32_t *)(pr + blk * 16 + c4) =
							    make_dir4(C, ztA, ztB, faA, faB, fbA, fbB, fa2A, fa2B, fb2A, fb2B);
					}
				}
			}
			sync_warp(gmask);
			++rows_exec;
			// ---- phase C: score tracking ----
			if (!EXACT) { // ksw2_extd2_sse.c:367-383; lane 0 only unless a Z-drop decision must be shared
				if (li == 0) {
					if (r > 0) {
						bool in0 = H0_t >= st0 && H0_t <= en0, in1 = H0_t + 1 >= st0 && H0_t + 1 <= en0;
						int s0 = wrap(st_slot + (H0_t - st), R); // only dereferenced when in range
						if (in0 && in1) {
							int d0 = hi8(g.v[s0]), d1 = hi8(g.u[wrap(s0 + 1, R)]);
							if (d0 > d1) H0 += d0;
							else H0 += d1, ++H0_t;
						} else if (in0) {
							H0 += hi8(g.v[s0]);
						} else {
							++H0_t;
							H0 += hi8(g.u[wrap(st_slot + (H0_t - st), R)]);
						}
					} else H0 = hi8(g.v[st_slot]) - C.qe_seed, H0_t = 0;
				}
				if (C.flag & KSW_F_APPROX_DROP) {
					int stop = 0;
					if (li == 0) { // ksw_apply_zdrop, ksw2.h:172-188
						if (H0 > res.max) res.max = H0, res.max_t = H0_t, res.max_q = r - H0_t;
						else if (H0_t >= res.max_t && r - H0_t >= res.max_q) {
							int tl = H0_t - res.max_t, ql = (r - H0_t) - res.max_q, l = tl > ql ? tl - ql : ql - tl;
							if (C.zdrop >= 0 && res.max - H0 > C.zdrop + l * C.e2) res.zdropped = 1, stop = 1;
						}
					}
					stop = (int)shfl_idx(gmask, (uint32_t)stop, 0, G);
					if (stop) break;
				}
				if (li == 0 && r == nrows - 1 && en0 == tlen - 1) res.score = H0;
			} else { // ksw2_extd2_sse.c:323-366
				int max_H, max_t;
				if (r > 0) {
					const int en1 = st0 + ((en0 - st0) & ~3);
					// H[en0] first, from the not yet updated H[en0-1]
					const int se = wrap(st_slot + (en0 - st), R);
					int Hen;
					if (en0 > 0) Hen = g.H[se == 0 ? R - 1 : se - 1] + hi8(g.u[se]);
					else Hen = g.H[se] + hi8(g.v[se]);
					sync_warp(gmask);
					long long best = hkey(Hen, 0xffffffffu);
					for (int t4 = (st0 & ~3) + li * 4; t4 < en0; t4 += G * 4) {
						int slot = wrap(st_slot + (t4 - st), R);
						for (int j = 0; j < 4; ++j) {
							int t = t4 + j;
							if (t < st0 || t >= en0) continue;
							int h = g.H[slot + j] + hi8(g.v[slot + j]);
							g.H[slot + j] = h;
							uint32_t prio = t < en1 ? (0x40000000u | (uint32_t)(3 - ((t - st0) & 3)) << 28 | (0x0fffffffu - (uint32_t)t))
							                        : (0x0fffffffu - (uint32_t)t);
							long long k = hkey(h, prio);
							if (k > best) best = k;
						}
					}
					if (li == 0) g.H[se] = Hen;
					for (int d = 1; d < G; d <<= 1) {
						uint32_t lo = shfl_xor(gmask, (uint32_t)best, d, G);
						uint32_t hi = shfl_xor(gmask, (uint32_t)((unsigned long long)best >> 32), d, G);
						long long o = (long long)(((unsigned long long)hi << 32) | lo);
						if (o > best) best = o;
					}
					max_H = (int)(best >> 32);
					uint32_t pr = (uint32_t)best;
					max_t = pr == 0xffffffffu ? en0 : (int)(0x0fffffffu - (pr & 0x0fffffffu));
					sync_warp(gmask);
				} else {
					if (li == 0) g.H[st_slot] = hi8(g.v[st_slot]) - C.qe_seed;
					sync_warp(gmask);
					max_H = g.H[st_slot], max_t = 0;
				}
				{ // end-of-target / end-of-query bests (all lanes read the same words)
					int He = g.H[wrap(st_slot + (en0 - st), R)], Hs = g.H[wrap(st_slot + (st0 - st), R)];
					if (en0 == tlen - 1 && He > res.mte) res.mte = He, res.mte_q = r - en;
					if (r - st0 == qlen - 1 && Hs > res.mqe) res.mqe = Hs, res.mqe_t = st0;
					int stop = 0;
					if (max_H > res.max) res.max = max_H, res.max_t = max_t, res.max_q = r - max_t;
					else if (max_t >= res.max_t && r - max_t >= res.max_q) {
						int tl = max_t - res.max_t, ql = (r - max_t) - res.max_q, l = tl > ql ? tl - ql : ql - tl;
						if (C.zdrop >= 0 && res.max - max_H > C.zdrop + l * C.e2) res.zdropped = 1, stop = 1;
					}
					if (stop) break;
					if (r == nrows - 1 && en0 == tlen - 1) res.score = g.H[wrap(st_slot + (tlen - 1 - st), R)];
				}
			}
			last_st = st, last_en = en;
		}
		if (li == 0) {
			res.rows_done = rows_exec;
			if (WITH_P) { // choice of the traceback start, ksw2_extd2_sse.c:389-400
				if (!res.zdropped && !(C.flag & KSW_F_EXTZ_ONLY)) res.tb_i = tlen - 1, res.tb_j = qlen - 1;
				else if (!res.zdropped && (C.flag & KSW_F_EXTZ_ONLY) && res.mqe + C.end_bonus > res.max)
					res.reach_end = 1, res.tb_i = res.mqe_t, res.tb_j = qlen - 1;
				else if (res.max_t >= 0 && res.max_q >= 0) res.tb_i = res.max_t, res.tb_j = res.max_q;
			}
			B.res[pair] = res;
		}
	}
}

// Stage one pair into the padded arenas the DP kernel reads (see KswBatch): target zero padded,
// query reversed (qr[] of ksw2_extd2_sse.c:128) with N(4) -> 8 (ksw2_extd2_avx.c:187-190).
GD_DEV void ksw_pack_pair(const uint8_t *GD_RESTRICT q, int qlen, const uint8_t *GD_RESTRICT t, int tlen,
                          uint8_t *GD_RESTRICT tpk, int t_stride, uint8_t *GD_RESTRICT qpk, int q_stride, int lane,
                          int nlanes)
{
	for (int i = lane * 4; i < t_stride; i += nlanes * 4) {
		uint32_t wv = 0;
		for (int k = 0; k < 4; ++k)
			if (i + k < tlen) wv |= (uint32_t)t[i + k] << (8 * k);
		*(uint32_t *)(tpk + i) = wv;
	}
	for (int i = lane * 4; i < q_stride; i += nlanes * 4) {
		uint32_t wv = 0;
		for (int k = 0; k < 4; ++k) {
			int j = i + k - 16;
			if (j >= 0 && j < qlen) {
				uint32_t c = q[qlen - 1 - j];
				wv |= (c == 4 ? 8u : c) << (8 * k);
			}
		}
		*(uint32_t *)(qpk + i) = wv;
	}
}

// One thread per pair: walk the backtrack matrix (ksw_backtrack, ksw2.h:115-163 with is_rot=1)
// and emit the run-length CIGAR in walk order (i.e. reversed) into cig_tmp[pair*stride ...].
GD_DEV void ksw_traceback_one(const KswBatch &B, int flag, int lp, uint32_t *cig_out, int stride)
{
	const int pair = B.base + lp;
	KswResult *res = &B.res[pair];
	int i = res->tb_i, j = res->tb_j, n = 0, state = 0, overflow = 0;
	if (i < 0 || j < 0) {
		res->n_cigar = 0;
		return;
	}
	const int qlen = B.qlen[pair], tlen = B.tlen[pair];
	int w = B.w ? B.w[pair] : B.w_all;
	if (w < 0) w = imax(tlen, qlen);
	const int ncol16 = ksw_ncol16(qlen, tlen, w);
	const uint8_t *p = B.p + (size_t)lp * B.p_stride;
	uint32_t *cig = cig_out + (size_t)pair * stride;
	uint32_t cur = 0; // pending op: len<<4|op, 0 = none
	while (i >= 0 && j >= 0) {
		int r = i + j, force = -1;
		Bounds bd;
		row_bounds(r, qlen, tlen, w, bd);
		if (i < bd.st) force = 2;
		if (i > bd.en) force = 1;
		uint32_t cell = force < 0 ? p[(size_t)r * ncol16 + (i - bd.st)] : 0;
		if (state == 0) state = cell & 7;
		else if (!((cell >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = cell & 7;
		if (force >= 0) state = force;
		uint32_t op;
		if (state == 0) op = 0, --i, --j;
		else if (state == 1 || state == 3) op = 2, --i;
		else op = 1, --j;
		if (cur && (cur & 0xf) == op) cur += 16;
		else {
			if (cur) {
				if (n < stride) cig[n] = cur;
				else overflow = 1;
				++n;
			}
			cur = 16 | op;
		}
	}
	if (i >= 0) { // leading deletion
		if (cur && (cur & 0xf) == 2) cur += (uint32_t)(i + 1) << 4;
		else {
			if (cur) {
				if (n < stride) cig[n] = cur;
				else overflow = 1;
				++n;
			}
			cur = (uint32_t)(i + 1) << 4 | 2;
		}
	}
	if (j >= 0) { // leading insertion
		if (cur && (cur & 0xf) == 1) cur += (uint32_t)(j + 1) << 4;
		else {
			if (cur) {
				if (n < stride) cig[n] = cur;
				else overflow = 1;
				++n;
			}
			cur = (uint32_t)(j + 1) << 4 | 1;
		}
	}
	if (cur) {
		if (n < stride) cig[n] = cur;
		else overflow = 1;
		++n;
	}
	if (!overflow && !(flag & KSW_F_REV_CIGAR)) // ksw2.h:158-160: forward order unless REV_CIGAR
		for (int k = 0; k < n >> 1; ++k) {
			uint32_t t = cig[k];
			cig[k] = cig[n - 1 - k], cig[n - 1 - k] = t;
		}
	res->n_cigar = overflow ? -n : n;
}

} // namespace gd
