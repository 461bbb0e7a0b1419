// gd_ksw.cuh -- batched ksw2 dual-affine banded extension DP for sm_100a (kernel design v2).
//
// Replaces ksw_extd2_sse / ksw_extd2_avx512 (GDiet-ShortReads/ksw2_extd2_sse.c:27-401,
// GDiet-ShortReads/ksw2_extd2_avx.c:72-915) for whole batches of (query,target) pairs.
//
// Mapping.  A *group* of G lanes (G = 4..32) owns one pair at a time and walks its anti-diagonals
// r = 0..qlen+tlen-2.  A row is cut into 8-column *chunks* (aligned to multiples of 8 target
// columns); one *step* of a group updates G consecutive chunks, one chunk per lane, and the steps
// of a row run from the highest chunk down to the lowest.  The reference's persistent per-column
// int8 state (u,v,x,y,x2,y2), its score row s and (exact mode) H live in shared memory in a ring
// of R columns.  Because chunks are visited in descending order, the previous-row values of the
// left neighbour column (x[t-1], v[t-1], x2[t-1]) are still in the ring when a lane needs them: no
// shuffles and no carries between steps; the boundary injection of the reference
// (ksw2_extd2_sse.c:149-159) is one store into the ring slot of column st-1.  The 16-cell range
// rounding, the stale score cells and the boundary injections are reproduced exactly (SURVEY.md
// 8 A2) because cells outside the true band leak into later rows.
//
// Arithmetic.  int8 with wrap-around is carried in the HIGH byte of each half of a 32-bit
// register (two cells per register), so VIADD.16x2 / VIMNMX.S16x2 / VIMNMX3.S16x2 / VIADDMNMX.S16x2
// give exact int8 modular add and signed compare on two cells per instruction.  The LOW byte of
// each half carries a tag (which of H,E,F,E~,F~ a value belongs to); the 5-way max therefore
// returns the arg-max in its low byte for free, with exactly the reference's tie order (strict
// '>' for left-aligned gaps, '>=' with KSW_EZ_RIGHT).  Inside a chunk the 8 columns c0..c7 sit in
// four registers as (c0,c4) (c1,c5) (c2,c6) (c3,c7), so "shift by one column" is ONE byte permute
// per state array; every per-chunk array in shared memory and the backtrack bytes use the same
// order (position of column j: ((j&3)<<1)|(j>>2)).
//
// Exact mode (no KSW_EZ_APPROX_MAX).  H[t] (int32) is updated for the whole chunk; the row maximum
// with the reference's 4-lane tie order is one 32-bit key per cell: (score relative to the
// previous row's maximum, clamped) << 16 | priority of the column, reduced with VIMNMX3.  The
// column en0 and the <= 3 columns of the scalar tail of the reference's scan get their update from
// four dedicated lanes after the chunk sweep; sentinels keep them (and the dead columns left and
// right of [st0,en0]) out of the bulk maximum.  Rows the clamped keys cannot represent fall back
// to a literal restatement of the scan.
//
// Backtrack.  One byte per cell, same bit layout as ksw2.h:127-130, written straight to HBM as
// one 8-byte store per lane-step; the CIGAR is produced by gd_ksw_traceback_kernel.
#pragma once
#include "gd_common.cuh"

#ifndef GD_KSW_PREFETCH
#define GD_KSW_PREFETCH 0 // 1: load the next step's chunk before computing the current one
#endif
// Three experiments that ncu's bank-conflict / partial-sector findings suggested (round 2), measured with tools/ksw_kbench.cu on a
// B200 (profiles/r2_kbench_variants.txt; 262,144 pairs of 150x200, live / exact mode, DP kernel alone, identical results) and
// left OFF because every one of them is slower: the kernel is bound by the integer-ALU pipe (67 % busy at 10 warps per SM), not
// by shared-memory wavefronts or HBM.
//   base 541.8 / 382.0 GCUPS;  ST16 539.4 / 378.2;  NBSHFL 476.9 / 338.7 (SHFL costs more than a 4-way conflicted 2-byte load);
//   P32 534.0 / 370.4 (HBM reads 11.9 -> 0.6 KB per pair, writes 38.6 -> 35.4 KB: the read-modify-write of partial sectors is gone,
//   but HBM was at 10 % of its bandwidth to begin with);  all three 467.3 / 333.7;  all + PREFETCH 433.9 / 307.7.
#ifndef GD_KSW_ST16
#define GD_KSW_ST16 0 // 1: score bytes + target codes of a record move as ONE 16-byte access (8-byte accesses of 4 groups collide on 2 banks)
#endif
#ifndef GD_KSW_NBSHFL
#define GD_KSW_NBSHFL 0 // 1: the left neighbour's (x,v) / (x2,u) slot comes from the neighbour lane's registers (SHFL) instead of a 2-byte shared load (4-way conflict)
#endif
#ifndef GD_KSW_P32
#define GD_KSW_P32 0 // 1: backtrack rows are pitched and padded to whole 32-byte sectors (no partial-sector writes -> no read-modify-write in HBM)
#endif
#ifndef GD_KSW_DP4A
#define GD_KSW_DP4A 1 // 1: exact mode adds v to the 32-bit H column with IDP.4A (one instruction, not on the integer pipe) instead of PRMT + IADD
#endif
#ifndef GD_KSW_NOTIMAD
#define GD_KSW_NOTIMAD 0 // 1: the three bitwise complements of a cell pair are IMADs (x * -1 - 1) instead of LOP3s
#endif
#ifndef GD_KSW_HOTMEM
#define GD_KSW_HOTMEM 1 // 1: sweep constants come from device memory (stay in registers) instead of the constant bank
#endif

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
#define GD_KSW_REL_FLOOR (-30000) // exact-max keys: scores are compared relative to the previous row's maximum
#define GD_KSW_POS_MAX 8190       // exact-max keys: 13-bit column position inside the live window
#define GD_KSW_LUT_BYTES 1280     // per-block lookup tables at the start of shared memory
#define GD_KSW_QFRONT 16          // zero bytes in front of the reversed query

// Per-pair result record (device + host). First 11 ints mirror ksw_extz_t (ksw2.h:31-40).
struct KswResult {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
	int32_t tb_i, tb_j; // traceback start cell (target, query) or -1: filled by the DP kernel
	int32_t rows_done;  // anti-diagonals executed (for cell accounting)
	int32_t lead64;     // 1: the walk entered the AVX-512 build's lead-in cells (see ksw_lead64_pair); CIGAR from that model
	int32_t pad1;
};

// Batch-uniform constants, precomputed on the host (gd_ksw_make_consts).
struct KswConsts {
	uint32_t MCH16, Q1, Q21, NEGQE, NEGQE2;
	uint32_t TS4;                // tag of the score term, replicated in all 4 bytes
	uint32_t TA, TB, TA2, TB2;   // 16x2 tags (low byte of each half)
	uint32_t TAGX;               // xor applied to the extracted tags -> 0..4 as in ksw2.h
	uint32_t MCH4, MIS4, SCN4;   // byte-replicated match / mismatch / ambiguous scores
	uint32_t INIT_A, INIT_B, INIT_C; // initial column state, two columns per word, packed as in the ring (see REC_*)
	int32_t q, e, q2, e2;        // after ordering the two pieces (ksw2_extd2_sse.c:78)
	int32_t qe_seed;             // q+e BEFORE ordering: seeds H (ksw2_extd2_sse.c:68,358,382)
	int32_t long_thres, long_diff;
	int32_t zdrop, end_bonus, flag;
	int32_t degenerate;          // m<=1 or -min(mat) > 2(q+e): every call returns the reset ez
	int32_t force_slow_max;      // test hook: always take the literal row-max scan (exact mode)
};

struct KswHot;

// One launch works on pairs [base, base+n) of the caller's batch ("chunk"); the packed sequence
// arenas and the backtrack arena are indexed by the chunk-local pair number with uniform strides.
struct KswBatch {
	int32_t n, base;
	const int32_t *qlen, *tlen, *w; // per pair (global index); w may be NULL -> w_all
	int32_t w_all;
	const uint8_t *tpk;  // target codes, zero padded, 8-column chunks in strided order; pair i at tpk + i*t_stride
	const uint8_t *qpk;  // reversed query, N(4)->8; pair i: qpk + i*q_stride + GD_KSW_QFRONT is element 0,
	                     // GD_KSW_QFRONT zero bytes in front and >= 48 behind
	int32_t t_stride, q_stride;
	uint8_t *p;          // backtrack arena, pair i at p + i*p_stride
	int64_t p_stride;
	KswResult *res;      // global index
	int32_t *ticket;     // dynamic pair dispenser (chunk-local)
	const KswHot *hot;   // sweep constants in device memory (see KswHot)
	int32_t ring;        // R: columns per ring (multiple of 8)
	int32_t group_smem;  // bytes of shared memory per group
	int32_t *lead64;     // [0]: number of, [1..]: chunk-local pairs whose walk entered the AVX-512 lead-in cells (may be NULL)
};

GD_DEV uint32_t pack16(int v) { return ((uint32_t)(v & 0xff) << 8) | ((uint32_t)(v & 0xff) << 24); }
GD_DEV int hi8(uint32_t h16) { return (int)(int8_t)(h16 >> 8); } // value of a 16-bit cell image
GD_DEV int chunk_pos(int j) { return ((j & 3) << 1) | ((j >> 2) & 1); } // place of column j inside its chunk

GD_DEV int gap_delta(int r, const KswConsts &C)
{ // first-row / first-column difference (ksw2_extd2_sse.c:158,162); int8 wrap applied by the caller
	return r == 0 ? -C.q - C.e : r < C.long_thres ? -C.e : r == C.long_thres ? C.long_diff : -C.e2;
}

// Two cells of the recurrence (ksw2_extd2_sse.c:38-66,228-274). All operands are 16x2 images.
// In: s (tag TS), xt1/x2t1 (left neighbours, tags TA/TA2), vt1, ut (no tag), y/y2 (tags TB/TB2).
// Out: new u,v,x,y,x2,y2, the word zt whose low bytes hold the arg-max tag, and four words whose
// bit 15/31 is the continuation flag of E,F,E~,F~.
template <bool RIGHT, class CT>
GD_DEV void cell2(const CT &C, uint32_t s, uint32_t xt1, uint32_t vt1, uint32_t x2t1, uint32_t ut, uint32_t &y,
                  uint32_t &y2, uint32_t &u_new, uint32_t &v_new, uint32_t &x_new, uint32_t &x2_new, uint32_t &zt,
                  uint32_t &fa, uint32_t &fb, uint32_t &fa2, uint32_t &fb2)
{
	uint32_t a = vadd2(xt1, vt1), b = vadd2(y, ut), a2 = vadd2(x2t1, vt1), b2 = vadd2(y2, ut);
	zt = vmax3(vmax3(s, a, b), a2, b2);
	const uint32_t z = vmin2(zt & 0xff00ff00u, C.MCH16);
	const uint32_t z1 = fma_add(z, C.ONE, 0x00010001u); // low bytes of z are 0: no carry between the halves
#if GD_KSW_NOTIMAD
	// ~x = x * (-1) - 1 (mod 2^32): an IMAD on the FMA pipe instead of a LOP3 on the integer pipe; C.ONE is opaque to the compiler
	const uint32_t m1 = 0u - C.ONE;
	u_new = vadd2(z1, fma_add(vt1, m1, 0xffffffffu)); // z - v[t-1]
	v_new = vadd2(z1, fma_add(ut, m1, 0xffffffffu));  // z - u[t]
	const uint32_t nz = fma_add(z, m1, 0xffffffffu);
#else
	u_new = vadd2(z1, ~vt1); // z - v[t-1]
	v_new = vadd2(z1, ~ut);  // z - u[t]
	const uint32_t nz = ~z;
#endif
	const uint32_t nzq = vadd2(nz, C.Q1), nzq2 = vadd2(nz, C.Q21); // q - z, q2 - z
	uint32_t ma, mb, ma2, mb2;
	if (!RIGHT) { // continuation iff value > 0  <=> high byte of max(value,0) >= 1
		ma = vaddmax2(a, nzq, C.TA), mb = vaddmax2(b, nzq, C.TB), ma2 = vaddmax2(a2, nzq2, C.TA2), mb2 = vaddmax2(b2, nzq2, C.TB2);
		fa = fma_add(ma, C.ONE, 0x7f007f00u), fb = fma_add(mb, C.ONE, 0x7f007f00u);
		fa2 = fma_add(ma2, C.ONE, 0x7f007f00u), fb2 = fma_add(mb2, C.ONE, 0x7f007f00u);
	} else { // continuation iff value >= 0 <=> sign bit clear
		a = vadd2(a, nzq), b = vadd2(b, nzq), a2 = vadd2(a2, nzq2), b2 = vadd2(b2, nzq2);
		ma = vmax2(a, C.TA), mb = vmax2(b, C.TB), ma2 = vmax2(a2, C.TA2), mb2 = vmax2(b2, C.TB2);
		fa = ~a, fb = ~b, fa2 = ~a2, fb2 = ~b2;
	}
	x_new = vadd2(ma, C.NEGQE), y = vadd2(mb, C.NEGQE), x2_new = vadd2(ma2, C.NEGQE2), y2 = vadd2(mb2, C.NEGQE2);
}

// 4 ksw2 backtrack bytes from two registers' worth of cells (A = (ca,ca+4), B = (cb,cb+4)); the
// byte order is (ca, ca+4, cb, cb+4), i.e. the strided chunk order.
template <class CT>
GD_DEV uint32_t make_dir4(const CT &C, uint32_t ztA, uint32_t ztB, uint32_t faA, uint32_t faB, uint32_t fbA,
                          uint32_t fbB, uint32_t fa2A, uint32_t fa2B, uint32_t fb2A, uint32_t fb2B)
{ // selector 0xfdb9: sign-replicated bytes 1,3 of A then 1,3 of B -> 0xff where the flag bit is set
	uint32_t d = (prmt(ztA, ztB, 0x6420) & 0x07070707u) ^ C.TAGX;
	d |= prmt(faA, faB, 0xfdb9) & 0x08080808u;
	d |= prmt(fbA, fbB, 0xfdb9) & 0x10101010u;
	d |= prmt(fa2A, fa2B, 0xfdb9) & 0x20202020u;
	d |= prmt(fb2A, fb2B, 0xfdb9) & 0x40404040u;
	return d;
}

// Score bytes for 4 cells with the AVX-512 xor-table rule (ksw2_extd2_avx.c:187-208,312-313):
// pmat[(t ^ q') & 15] with pmat = {mch, mis x3, scN x9, 0 x3}; a byte with bit 7 set gives 0.
template <class CT>
GD_DEV uint32_t score4(const CT &C, uint32_t tc, uint32_t qc)
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
{ // row pitch of the backtrack matrix (ksw2_extd2_sse.c:92-94), rounded up to whole 32-byte sectors
	int n = imin(imin(qlen, tlen), w + 1);
	n = ((n + 15) / 16 + 1) * 16;
#if GD_KSW_P32
	n = (n + 31) & ~31;
#endif
	return n;
}

// Shared memory of one group: a ring of NR = R/8 chunk records followed by the staged sequences.
// One record holds everything the sweep needs for 8 columns, so a lane-step addresses it with
// immediate offsets.  The six int8 state arrays are stored two to a 16-bit slot (8 slots per array
// pair, chunk-strided order): A = (x << 8 | v), B = (x2 << 8 | u), C = (y << 8 | y2); then the score
// bytes S and the target codes T of the 8 columns (8 B each, chunk-strided; T is filled from the packed
// target arena when the block enters the window) and, in exact mode, H (8 x int32, natural column order).  Record sizes
// 80 / 96 B keep the four lanes of a group and the two groups that share a 128-bit access phase
// (group pitch == 64 resp. 16 mod 128) on disjoint banks.
enum { REC_A = 0, REC_B = 16, REC_C = 32, REC_S = 48, REC_T = 56, REC_H = 64 };
template <bool EXACT> struct RecSize { enum { value = EXACT ? 96 : 80 }; };
static inline int ksw_group_smem_bytes(int R, bool exact, int seq_bytes /* 0 when the sequences stay in global memory */)
{
	int b = (R / 8) * (exact ? 96 : 80) + seq_bytes;
	const int phase = exact ? 16 : 64; // smallest size >= b that is == phase (mod 128)
	return (b - phase + 127) / 128 * 128 + phase;
}

// Per-block lookup tables (first GD_KSW_LUT_BYTES of shared memory):
//   [0,1152)    uint2 fresh[16][9]: byte masks (chunk-strided order) of the columns j of a chunk with lo <= j < hi
//   [1152,1216) uint32 pk[4][4]:    exact-max priority constants of register k for c0 = (-st0)&3
GD_DEV void ksw_build_lut(uint8_t *lut, int tid, int nthreads)
{
	for (int i = tid; i < 144; i += nthreads) {
		const int lo = i / 9, hi = i % 9;
		uint32_t w0 = 0, w1 = 0;
		for (int j = 0; j < 8; ++j)
			if (j >= lo && j < hi) {
				const int p = chunk_pos(j);
				if (p < 4) w0 |= 0xffu << (8 * p);
				else w1 |= 0xffu << (8 * (p - 4));
			}
		uint2 m;
		m.x = w0, m.y = w1;
		((uint2 *)lut)[i] = m;
	}
	for (int i = tid; i < 16; i += nthreads) {
		const int c0 = i >> 2, k = i & 3;
		const uint32_t a = (0x8000u | (uint32_t)(3 - ((k + c0) & 3)) << 13) + (uint32_t)(GD_KSW_POS_MAX - k);
		((uint32_t *)(lut + 1152))[i] = a | ((a - 4) << 16);
	}
}

GD_DEV int wrap(int k, int n) { return k >= n ? k - n : k; }
// The record of column t sits at ring position (t >> 3) mod NR; the kernel follows the few columns it
// patches or reads one at a time (r, st-1, H0_t, en0, st0) with byte offsets that step through the ring.
GD_DEV int ring_fwd(int off, int rec_bytes, int ring_bytes) { return off + rec_bytes >= ring_bytes ? 0 : off + rec_bytes; }
GD_DEV int ring_bwd(int off, int rec_bytes, int ring_bytes) { return off == 0 ? ring_bytes - rec_bytes : off - rec_bytes; }
GD_DEV int pos2(int j) { return ((j & 3) << 2) | ((j >> 1) & 2); } // 2 * chunk_pos(j): byte offset of column j's 16-bit slot

// Record of the column at distance d = t - st from the first column of the row's 16-aligned range
// (d may be -1: the left-boundary column).  st_rec is the ring position of the record that holds column st.
GD_DEV uint8_t *col_rec(uint8_t *ring, int rec_bytes, int NR, int st_rec, int d)
{
	int k = st_rec + (d >> 3);
	if (k >= NR) k -= NR;
	if (k < 0) k += NR;
	return ring + k * rec_bytes;
}
GD_DEV uint16_t *col_slot(uint8_t *ring, int rec_bytes, int NR, int st_rec, int d, int arr)
{ // 16-bit slot of one column in array pair `arr` (REC_A/REC_B/REC_C): high byte / low byte as listed at REC_*
	return (uint16_t *)(col_rec(ring, rec_bytes, NR, st_rec, d) + arr + 2 * chunk_pos(d & 7));
}
GD_DEV int col_lo8(uint8_t *ring, int rec_bytes, int NR, int st_rec, int d, int arr)
{ // v (REC_A) or u (REC_B) of one column as a signed value
	return (int)(int8_t)(*col_slot(ring, rec_bytes, NR, st_rec, d, arr) & 0xff);
}
GD_DEV int32_t *col_H(uint8_t *ring, int rec_bytes, int NR, int st_rec, int d)
{
	return (int32_t *)(col_rec(ring, rec_bytes, NR, st_rec, d) + REC_H + 4 * (d & 7));
}

// Literal restatement of the row-maximum scan (ksw2_extd2_sse.c:327-357) over the updated H row:
// only used when the relative 16-bit keys of the fast path cannot represent a row (never on real data).
GD_DEV void row_max_literal(uint8_t *ring, int rec_bytes, int NR, int st_rec, int st, int st0, int en0, int &max_H,
                            int &max_t)
{
	const int en1 = st0 + ((en0 - st0) & ~3);
	max_H = *col_H(ring, rec_bytes, NR, st_rec, en0 - st), max_t = en0;
	int HH[4], tt[4];
	for (int i = 0; i < 4; ++i) HH[i] = max_H, tt[i] = max_t;
	for (int t = st0; t < en1; t += 4)
		for (int i = 0; i < 4; ++i) {
			const int h = *col_H(ring, rec_bytes, NR, st_rec, t + i - st);
			if (h > HH[i]) HH[i] = h, tt[i] = t;
		}
	for (int i = 0; i < 4; ++i)
		if (max_H < HH[i]) max_H = HH[i], max_t = tt[i] + i;
	for (int t = en1; t < en0; ++t) {
		const int h = *col_H(ring, rec_bytes, NR, st_rec, t - st);
		if (h > max_H) max_H = h, max_t = t;
	}
}

// The constants the chunk sweep touches every step.  They are read from device memory (written by the
// pack kernel) rather than from the kernel parameters: values that come from the constant bank are
// re-loaded by ptxas inside the loop (10 LDC per step), values loaded from global memory stay in registers.
struct KswHot {
	uint32_t MCH16, Q1, Q21, NEGQE, NEGQE2, TS4, TA, TB, TA2, TB2, TAGX, MCH4, MIS4, SCN4;
	uint32_t ONE;  // the constant 1, opaque to the compiler: fma_add(a, ONE, c) = a + c issues on the FMA pipe (IMAD)
	uint32_t pad1; // instead of the integer-ALU pipe that bounds the kernel
};
GD_DEV KswHot ksw_hot_from_consts(const KswConsts &C)
{
	KswHot h;
	h.MCH16 = C.MCH16, h.Q1 = C.Q1, h.Q21 = C.Q21, h.NEGQE = C.NEGQE, h.NEGQE2 = C.NEGQE2, h.TS4 = C.TS4, h.TA = C.TA;
	h.TB = C.TB, h.TA2 = C.TA2, h.TB2 = C.TB2, h.TAGX = C.TAGX, h.MCH4 = C.MCH4, h.MIS4 = C.MIS4, h.SCN4 = C.SCN4;
	h.ONE = 1, h.pad1 = 0;
	return h;
}
GD_DEV KswHot ksw_hot_load(const KswHot *p)
{
	KswHot h;
	const uint32_t *w = (const uint32_t *)p;
	uint32_t *o = (uint32_t *)&h;
	for (int i = 0; i < 16; ++i) o[i] = ld_volatile(w + i);
	return h;
}

// Everything one lane reads from shared memory for one chunk (loaded one step ahead of its use).
struct StepIn {
	uint4 SA, SB, SC, ha, hb;
	uint2 old, tw;
	uint32_t am1, bm1, q0, q1, q2;
};

GD_DEV uint4 rep4(uint32_t v)
{
	uint4 q;
	q.x = q.y = q.z = q.w = v;
	return q;
}

// Collectives of one "gang": the lanes that execute the row loop together.  G <= 32: a warp that carries
// 32/G pairs (groups) at once; G > 32 (block-per-pair, for the widest bands): one pair per thread block
// of G threads, where every value the row loop votes on is already uniform.
template <int G> struct Gang {
	static GD_MEM void sync() { if (G <= 32) sync_warp(0xffffffffu); else sync_block(); }
	static GD_MEM bool any(bool p) { return G <= 32 ? ballot(0xffffffffu, p) != 0 : p; }
	static GD_MEM int max_steps(int v) { return G <= 32 ? reduce_max(0xffffffffu, v) : v; }
	// value held by lane 0 of every group -> all lanes of that group (scratch: >= 1 int of shared memory per block)
	static GD_MEM int from_leader(int v, int lane, int *scratch)
	{
		if (G <= 32) return (int)shfl_idx(0xffffffffu, (uint32_t)v, lane & ~(G - 1), 32);
		if (lane == 0) scratch[0] = v;
		sync_block();
		const int r = scratch[0];
		sync_block();
		return r;
	}
	// maximum over the lanes of a group (scratch: >= G/32 ints)
	static GD_MEM int group_max(int v, int lane, int *scratch)
	{
		for (int dd = 1; dd < (G <= 32 ? G : 32); dd <<= 1) v = imax(v, (int)shfl_xor(0xffffffffu, (uint32_t)v, dd, 32));
		if (G <= 32) return v;
		if ((lane & 31) == 0) scratch[1 + (lane >> 5)] = v;
		sync_block();
		for (int i = 0; i < G / 32; ++i) v = imax(v, scratch[1 + i]);
		sync_block();
		return v;
	}
};

// One warp, 32/G pairs at a time.  All control flow is warp-uniform (every loop runs for the
// maximum trip count over the groups of the warp, per-lane work is predicated), so every barrier
// and vote uses the full mask.  A group that finishes its pair fetches the next one at once.
// MODE: 0 = KSW_EZ_APPROX_MAX, 1 = APPROX_MAX | APPROX_DROP, 2 = exact maximum (+ Z-drop, mqe, mte).
template <int G, bool RIGHT, int MODE, bool WITH_P>
GD_DEV void ksw_warp_body(const KswConsts &C, const KswBatch &B, uint8_t *smem_warp, const uint8_t *lut, int lane)
{
	const bool EXACT = MODE == 2;
	const uint32_t FULL = 0xffffffffu;
	const int REC = RecSize<EXACT>::value;
	const int li = lane & (G - 1);
	int *const scratch = (int *)(lut + 1216); // spare words of the lookup-table block (block-per-pair collectives)
	uint8_t *const ring = smem_warp + (size_t)(lane / G) * B.group_smem;
	const int NR = B.ring >> 3;
	// The target codes travel in the ring records.  Short pairs (G <= 8) stage the reversed query next to
	// the ring; long ones read the packed query arena in global memory (coalesced: the lanes of a step
	// read consecutive chunks).
	const bool SEQ_SMEM = G <= 8;
	const uint8_t *qsm = SEQ_SMEM ? ring + NR * REC : B.qpk;
	const uint8_t *tpk_g = B.tpk; // packed target of the current pair (feeds the records of entering blocks)
	const uint2 *lut_fresh = (const uint2 *)lut;
	const uint32_t *lut_pk = (const uint32_t *)(lut + 1152);
#if GD_KSW_HOTMEM
	const KswHot K = ksw_hot_load(B.hot);
#else
	const KswHot K = ksw_hot_from_consts(C);
#endif

	// ---- per-group state (identical in all lanes of a group) ----
	bool have = false, done = false;
	int pair = 0, qlen = 1, tlen = 1, w = 0, T16 = 16, nblk_t = 1, ncol16 = 32, nrows = 1;
	uint8_t *prow = 0; // backtrack row pointer of the current row
	const int RB = NR * REC; // ring bytes
	int r = 0, rows_exec = 0, st_rec = 0, st_cur = 0, init_hi = 1;
	int r_off = 0;                          // ring offset of the record of column r
	int T_off = 0;                          // approx mode: ... of column H0_t
	int en0_cur = 0, en0_off = 0, st0_cur = 0, st0_off = 0; // exact mode: ... of columns en0 and st0
	int H0 = 0, H0_t = 0;                   // approx mode
	int Mprev = 0, Hleft = 0, st0_prev = 0, Hs_prev = 0; // exact mode
	int32_t *Hs_ptr = (int32_t *)(ring + REC_H);          // exact mode: H[st0] of the previous row
	KswResult res;
	res.max = 0, res.zdropped = 0, res.max_q = res.max_t = res.mqe_t = res.mte_q = -1;
	res.score = res.mqe = res.mte = GD_KSW_NEG_INF, res.n_cigar = 0, res.reach_end = 0;
	res.tb_i = res.tb_j = -1, res.rows_done = 0, res.lead64 = res.pad1 = 0;

	for (;;) {
		// ================= fetch: groups without a pair take the next ticket =================
		const bool need = !have && !done;
		if (Gang<G>::any(need)) {
			int lp = 0;
			if (need && li == 0) lp = atomic_add(B.ticket, 1);
			lp = Gang<G>::from_leader(lp, lane, scratch);
			bool fresh = false;
			const uint8_t *tpk = 0, *qpk = 0;
			if (need) {
				if (lp >= B.n) done = true;
				else {
					pair = B.base + lp;
					qlen = B.qlen[pair], tlen = B.tlen[pair];
					w = B.w ? B.w[pair] : B.w_all;
					res.max = 0, res.zdropped = 0, res.max_q = res.max_t = res.mqe_t = res.mte_q = -1;
					res.score = res.mqe = res.mte = GD_KSW_NEG_INF, res.n_cigar = 0, res.reach_end = 0;
					res.tb_i = res.tb_j = -1, res.rows_done = 0;
					if (C.degenerate || qlen <= 0 || tlen <= 0) {
						if (li == 0) B.res[pair] = res; // reset record, ksw2_extd2_sse.c:75-76,100
						qlen = tlen = 1;
					} else {
						if (w < 0) w = imax(tlen, qlen);
						T16 = (tlen + 15) & ~15, nblk_t = T16 >> 4;
						ncol16 = ksw_ncol16(qlen, tlen, w);
						nrows = qlen + tlen - 1;
						tpk = B.tpk + (size_t)lp * B.t_stride;
						qpk = B.qpk + (size_t)lp * B.q_stride;
						prow = WITH_P ? B.p + (size_t)lp * B.p_stride : 0;
						r = 0, rows_exec = 0, st_rec = 0, st_cur = 0, init_hi = 1;
						r_off = T_off = 0, en0_cur = en0_off = st0_cur = st0_off = 0;
						H0 = -C.qe_seed, H0_t = 0; // row 0 adds v[0]: H0 = v[0] - qe (ksw2_extd2_sse.c:382)
						Mprev = -C.qe_seed, Hleft = GD_KSW_NEG_INF, st0_prev = 0, Hs_prev = GD_KSW_NEG_INF;
						have = true, fresh = true;
					}
				}
			}
			// stage the padded sequences of fresh pairs (strides are launch-uniform) and give block 0 of the
			// ring the reference's initial values; the previous pair's readers are past their last row
			{
				if (SEQ_SMEM) {
					const int nq = B.q_stride >> 2;
					uint32_t *const qdst = (uint32_t *)(ring + NR * REC);
					for (int i = li; i < nq; i += G)
						if (fresh) qdst[i] = ((const uint32_t *)qpk)[i];
				} else if (fresh) qsm = qpk;
				if (fresh) tpk_g = tpk;
				if (fresh && li < 2) {
					uint8_t *rc = ring + li * REC;
					*(uint4 *)(rc + REC_A) = rep4(C.INIT_A), *(uint4 *)(rc + REC_B) = rep4(C.INIT_B);
					*(uint4 *)(rc + REC_C) = rep4(C.INIT_C);
					uint2 z2;
					z2.x = z2.y = 0;
					*(uint2 *)(rc + REC_S) = z2;
					*(uint2 *)(rc + REC_T) = *(const uint2 *)(tpk + li * 8);
					if (EXACT) {
						uint4 h0 = rep4((uint32_t)GD_KSW_NEG_INF);
						*(uint4 *)(rc + REC_H + 16) = h0;
						if (li == 0) h0.x = (uint32_t)(-C.qe_seed); // the bulk update of row 0 adds v[0]: H[0] = v[0] - qe
						*(uint4 *)(rc + REC_H) = h0;
					}
				}
			}
			Gang<G>::sync();
			if (!Gang<G>::any(have)) break; // nothing left anywhere in this warp / block
		}
		// ================= one anti-diagonal for every group that has a pair =================
		// (groups without a pair run the same instructions on a one-cell dummy row; their stores are off)
		int st0 = imax(imax(0, r - qlen + 1), (r - w + 1) >> 1); // ksw2_extd2_sse.c:133-147
		int en0 = imin(imin(tlen - 1, r), (r + w) >> 1);
		bool finish = false;
		bool active = have;
		if (have && st0 > en0) res.zdropped = 1, finish = true, active = false; // band closed, ksw2_extd2_sse.c:142-145
		if (!active) st0 = 0, en0 = 0;
		const int st = st0 & ~15, en = en0 | 15;
		// ---- phase A: ring bookkeeping, boundary injections ----
		const bool adv = active && st != st_cur; // st moves by exactly one 16-column block
		if (adv) st_rec = wrap(st_rec + 2, NR), st_cur = st;
		if (EXACT && active) { // en0 and st0 grow by at most one per row
			if (en0 != en0_cur) {
				en0_cur = en0;
				if ((en0 & 7) == 0) en0_off = ring_fwd(en0_off, REC, RB);
			}
			if (st0 != st0_cur) {
				st0_cur = st0;
				if ((st0 & 7) == 0) st0_off = ring_fwd(st0_off, REC, RB);
			}
		}
		const int fe = active ? imin(st0 + (((en0 - st0) >> 4) + 1) * 16, T16) : 16; // score row is rewritten on [st0, fe)
		const int en1 = st0 + ((en0 - st0) & ~3); // exact mode: end of the 4-lane part of the row scan
		{
			// The block that enters the window gets the reference's initial values (its memset / kcalloc).
			// en0 grows by at most one per row, so at most one block enters, and none of this row's
			// single-column patches below can fall into it (they touch columns <= en0|15; block 0 is
			// initialised when the pair is fetched).
			const bool need_init = active && init_hi <= imin((en0 + 15) >> 4, nblk_t - 1);
			if (Gang<G>::any(need_init)) {
				if (need_init && li < 2) {
					uint8_t *rc = col_rec(ring, REC, NR, st_rec, init_hi * 16 + li * 8 - st);
					*(uint4 *)(rc + REC_A) = rep4(C.INIT_A), *(uint4 *)(rc + REC_B) = rep4(C.INIT_B);
					*(uint4 *)(rc + REC_C) = rep4(C.INIT_C);
					uint2 z2;
					z2.x = z2.y = 0;
					*(uint2 *)(rc + REC_S) = z2;
					*(uint2 *)(rc + REC_T) = *(const uint2 *)(tpk_g + init_hi * 16 + li * 8);
					if (EXACT)
						*(uint4 *)(rc + REC_H) = rep4((uint32_t)GD_KSW_NEG_INF), *(uint4 *)(rc + REC_H + 16) = rep4((uint32_t)GD_KSW_NEG_INF);
				}
				if (need_init) ++init_hi;
			}
		}
		{
			// Two single-column patches, one lane each (the address arithmetic is shared):
			//  lane G-1: y[r], y2[r], u[r] when column r is in the row's range (ksw2_extd2_sse.c:160-163);
			//  lane 0:   the left boundary x1, x21, v1 (ksw2_extd2_sse.c:149-159); it lives in the ring slot of
			//            column st-1, which already holds the previous row's values when that column was in its range.
			const int gd = gap_delta(r, C);
			uint8_t *pc = ring + (li == G - 1 ? r_off + pos2(r & 7) : ring_bwd(st_rec * REC, REC, RB) + 14);
			if (active && li == G - 1 && en >= r) {
				*(uint16_t *)(pc + REC_C) = (uint16_t)(C.INIT_C & 0xffff); // y, y2
				*(pc + REC_B) = (uint8_t)gd;                               // u (low byte of B)
			}
			if (active && li == 0 && !adv) { // st-1 was not in the previous row's range (or st == 0)
				const uint32_t v1 = st > 0 ? (C.INIT_A & 0xffu) : (uint32_t)(gd & 0xff);
				*(uint16_t *)(pc + REC_A) = (uint16_t)((C.INIT_A & 0xff00u) | v1); // x, v
				*(pc + REC_B + 1) = (uint8_t)(C.INIT_B >> 8);                      // x2 (high byte of B)
			}
		}
#if GD_KSW_P32
		// rows end on a 32-byte sector: a row of an odd number of 16-cell blocks gets 16 zero bytes behind it, so HBM never
		// sees a partially written sector (which costs a read-modify-write)
		if (WITH_P && active && li == 1 && ((en - st + 1) & 16)) *(uint4 *)(prow + (en - st + 1)) = rep4(0u);
#endif
		// Exact mode: column en0 (H[en0-1] + u[en0]) and the <= 3 columns of the scalar tail of the reference's
		// row scan are kept out of the bulk update.  Lane li < 4 of the group owns column en0-3+li.
		bool sp = false, sp_en0 = false;
		int sp_h = 0, sp_cc = 0;
		int32_t *sp_hp = 0;
		uint8_t *sp_bp = 0;
		if (EXACT) {
			const int tcol = en0 - 3 + li;
			sp_en0 = li == 3 && en0 > 0;
			sp = active && li < 4 && (sp_en0 || (li < 3 && tcol >= en1));
			// what the lane reads: H[tcol] of the previous row; lane 3 reads H[en0-1] instead (column en0 starts
			// from it, or, on one-cell rows, from the last score of the column that left the band on the left
			// = the previous row's H[st0])
			const int lcol = li == 3 ? en0 - 1 : tcol;
			const bool lok = active && li < 4 && lcol >= st0;
			const int lj = lok ? lcol & 7 : en0 & 7;
			uint8_t *lrec = ring + (lok && (en0 & 7) < en0 - lcol ? ring_bwd(en0_off, REC, RB) : en0_off);
			sp_h = *((const int32_t *)(lrec + REC_H) + lj);
			if (active && r > 0 && st0 > st0_prev) Hleft = Hs_prev;
			if (li == 3 && !lok) sp_h = Hleft;
			// what the lane owns: column tcol
			uint8_t *rsp = li == 3 ? ring + en0_off : lrec;
			const int jsp = li == 3 ? en0 & 7 : lj;
			sp_cc = tcol - st;
			sp_hp = (int32_t *)(rsp + REC_H) + jsp;
			sp_bp = rsp + pos2(jsp);
			Gang<G>::sync(); // all of the loads above precede the sentinel stores below
			if (sp) *sp_hp = GD_KSW_NEG_INF; // keeps the bulk scan off these cells
			if (active && li == 0 && r > 0 && st0 > st0_prev) *Hs_ptr = GD_KSW_NEG_INF; // ... and off the column that left
		}
		Gang<G>::sync();
		// ---- phase B: score row + core update, chunk by chunk from the right end of the row ----
		int run = (int)0x80000000; // exact mode: best (relative score << 16 | priority) key of this lane
		{
			const int cbeg = st >> 3;
			const int ctop = imax(en, fe - 1) >> 3; // chunk of lane G-1 in the first step
			const int nsteps = Gang<G>::max_steps(active ? (ctop - cbeg + G) / G : 0);
			const int qshift = active ? qlen - 1 - r + GD_KSW_QFRONT : GD_KSW_QFRONT; // qsm offset of the query base under column 0
			const uint32_t qsh = (uint32_t)(qshift & 3) * 8; // chunks start at multiples of 8: same byte phase in every step
			const int nMprev = -Mprev;
			uint32_t pk0 = 0, pk1 = 0, pk2 = 0, pk3 = 0;
			if (EXACT) {
				const uint4 pk = *(const uint4 *)(lut_pk + (((0 - st0) & 3) << 2));
				pk0 = pk.x, pk1 = pk.y, pk2 = pk.z, pk3 = pk.w;
			}
			// Lanes without a chunk in a step run the same instructions on chunk cbeg of their own ring (all
			// addresses stay in range) and only their stores are predicated off: no divergence.
			int c0 = ctop - (G - 1 - li);      // this lane's chunk in the current step (may be below cbeg)
			bool valid = active && c0 >= cbeg;
			int c = valid ? c0 : cbeg;
			int k = st_rec + (c - cbeg);       // ring position of the chunk's record
			if (k >= NR) k -= NR;
			auto load_step = [&](int cc, int kk, StepIn &in) {
				uint8_t *const rc = ring + kk * REC;
				uint8_t *const rp = ring + (kk == 0 ? NR - 1 : kk - 1) * REC; // record of the left neighbour chunk
				const int tb = cc << 3;
#if GD_KSW_ST16
				{
					const uint4 st = *(const uint4 *)(rc + REC_S); // REC_S and REC_T are adjacent
					in.old.x = st.x, in.old.y = st.y, in.tw.x = st.z, in.tw.y = st.w;
				}
#else
				in.tw = *(const uint2 *)(rc + REC_T);
				in.old = *(const uint2 *)(rc + REC_S);
#endif
				const uint32_t *qw = (const uint32_t *)(qsm + ((qshift + tb) & ~3));
				in.q0 = qw[0], in.q1 = qw[1], in.q2 = qw[2];
				in.SA = *(const uint4 *)(rc + REC_A), in.SB = *(const uint4 *)(rc + REC_B), in.SC = *(const uint4 *)(rc + REC_C);
#if GD_KSW_NBSHFL
				{ // lane li-1 of the gang holds the chunk to the left (its last slot = column tb-1); the first lane of a warp's
				  // gang segment, and the lane on the row's first chunk (whose left neighbour is the boundary slot), read the ring
					const uint32_t na = shfl_up(0xffffffffu, in.SA.w, 1, G <= 32 ? G : 32), nb = shfl_up(0xffffffffu, in.SB.w, 1, G <= 32 ? G : 32);
					in.am1 = na >> 16, in.bm1 = nb >> 16;
					if ((lane & ((G <= 32 ? G : 32) - 1)) == 0 || cc == cbeg)
						in.am1 = *(const uint16_t *)(rp + REC_A + 14), in.bm1 = *(const uint16_t *)(rp + REC_B + 14);
				}
#else
				in.am1 = *(const uint16_t *)(rp + REC_A + 14), in.bm1 = *(const uint16_t *)(rp + REC_B + 14);
#endif
				if (EXACT) in.ha = *(const uint4 *)(rc + REC_H), in.hb = *(const uint4 *)(rc + REC_H + 16);
			};
#if GD_KSW_PREFETCH
			StepIn nx;
			load_step(c, k, nx);
#endif
			for (int step = 0; step < nsteps; ++step) {
#if GD_KSW_PREFETCH
				const StepIn in = nx;
#else
				StepIn in;
				load_step(c, k, in);
#endif
				const bool cvalid = valid;
				const int tb = c << 3, d = tb - st;
				const bool core = cvalid && tb <= en;
				uint8_t *const rc = ring + k * REC;
				// the next step's chunk: G records further down; its loads do not touch anything this step stores
				c0 -= G;
				valid = active && c0 >= cbeg;
				c = valid ? c0 : cbeg;
				k = valid ? k - G : st_rec;
				if (k < 0) k += NR;
#if GD_KSW_PREFETCH
				if (step + 1 < nsteps) load_step(c, k, nx);
#endif
				// score bytes of the 8 columns (xor-table rule), merged with the stale row outside [st0,fe)
				uint32_t sw0, sw1;
				{
					const uint32_t qa = funnel_r(in.q0, in.q1, qsh), qb = funnel_r(in.q1, in.q2, qsh);
					const uint32_t f0 = score4(K, in.tw.x, prmt(qa, qb, 0x5140)), f1 = score4(K, in.tw.y, prmt(qa, qb, 0x7362));
					const int lo = imax(st0 - tb, 0), hi = imax(imin(fe - tb, 8), 0);
					const uint2 m = lut_fresh[lo * 9 + hi];
					sw0 = (f0 & m.x) | (in.old.x & ~m.x), sw1 = (f1 & m.y) | (in.old.y & ~m.y);
				}
				// unpack the slots into tagged 16x2 images; x,v,x2 come from the column to the left
				const uint4 SA = in.SA, SB = in.SB, SC = in.SC;
				const uint32_t SAs = prmt(in.am1, SA.w, 0x5410), SBs = prmt(in.bm1, SB.w, 0x5410); // (c-1, c3)
				const uint32_t HI = 0xff00ff00u;
				uint32_t u0, v0, x0, x20, zt0, fa0, fb0, fa20, fb20, y0 = (SC.x & HI) | K.TB, y20 = prmt(SC.x, K.TB2, 0x2404);
				uint32_t u1, v1, x1, x21, zt1, fa1, fb1, fa21, fb21, y1 = (SC.y & HI) | K.TB, y21 = prmt(SC.y, K.TB2, 0x2404);
				uint32_t u2, v2, x2, x22, zt2, fa2, fb2, fa22, fb22, y2 = (SC.z & HI) | K.TB, y22 = prmt(SC.z, K.TB2, 0x2404);
				uint32_t u3, v3, x3, x23, zt3, fa3, fb3, fa23, fb23, y3 = (SC.w & HI) | K.TB, y23 = prmt(SC.w, K.TB2, 0x2404);
				cell2<RIGHT>(K, prmt(sw0, K.TS4, 0x1404), (SAs & HI) | K.TA, prmt(SAs, 0, 0x2404), (SBs & HI) | K.TA2,
				             prmt(SB.x, 0, 0x2404), y0, y20, u0, v0, x0, x20, zt0, fa0, fb0, fa20, fb20);
				cell2<RIGHT>(K, prmt(sw0, K.TS4, 0x3424), (SA.x & HI) | K.TA, prmt(SA.x, 0, 0x2404), (SB.x & HI) | K.TA2,
				             prmt(SB.y, 0, 0x2404), y1, y21, u1, v1, x1, x21, zt1, fa1, fb1, fa21, fb21);
				cell2<RIGHT>(K, prmt(sw1, K.TS4, 0x1404), (SA.y & HI) | K.TA, prmt(SA.y, 0, 0x2404), (SB.y & HI) | K.TA2,
				             prmt(SB.z, 0, 0x2404), y2, y22, u2, v2, x2, x22, zt2, fa2, fb2, fa22, fb22);
				cell2<RIGHT>(K, prmt(sw1, K.TS4, 0x3424), (SA.z & HI) | K.TA, prmt(SA.z, 0, 0x2404), (SB.z & HI) | K.TA2,
				             prmt(SB.w, 0, 0x2404), y3, y23, u3, v3, x3, x23, zt3, fa3, fb3, fa23, fb23);
				Gang<G>::sync(); // every lane has read its left neighbour's old column before anybody stores
				if (cvalid) {
#if GD_KSW_ST16
					uint4 sv; // the target codes go back unchanged: one conflict-free 16-byte store
					sv.x = sw0, sv.y = sw1, sv.z = in.tw.x, sv.w = in.tw.y;
					*(uint4 *)(rc + REC_S) = sv;
#else
					uint2 sv;
					sv.x = sw0, sv.y = sw1;
					*(uint2 *)(rc + REC_S) = sv;
#endif
				}
				if (core) {
					uint4 o; // repack: selector 0x3715 = (hi.b3, lo.b3, hi.b1, lo.b1)
					o.x = prmt(x0, v0, 0x3715), o.y = prmt(x1, v1, 0x3715), o.z = prmt(x2, v2, 0x3715), o.w = prmt(x3, v3, 0x3715);
					*(uint4 *)(rc + REC_A) = o;
					o.x = prmt(x20, u0, 0x3715), o.y = prmt(x21, u1, 0x3715), o.z = prmt(x22, u2, 0x3715), o.w = prmt(x23, u3, 0x3715);
					*(uint4 *)(rc + REC_B) = o;
					o.x = prmt(y0, y20, 0x3715), o.y = prmt(y1, y21, 0x3715), o.z = prmt(y2, y22, 0x3715), o.w = prmt(y3, y23, 0x3715);
					*(uint4 *)(rc + REC_C) = o;
					if (WITH_P) {
						uint2 dv;
						dv.x = make_dir4(K, zt0, zt1, fa0, fa1, fb0, fb1, fa20, fa21, fb20, fb21);
						dv.y = make_dir4(K, zt2, zt3, fa2, fa3, fb2, fb3, fa22, fa23, fb22, fb23);
						*(uint2 *)(prow + d) = dv;
					}
				}
				if (EXACT) { // H[t] += v[t] on the whole chunk; columns outside [st0,en1) hold sentinels
					uint4 ha = in.ha, hb = in.hb;
#if GD_KSW_DP4A
					ha.x = add_sbyte(ha.x, v0, 1), hb.x = add_sbyte(hb.x, v0, 3); // v of columns (c, c+4) sits in bytes 1 and 3
					ha.y = add_sbyte(ha.y, v1, 1), hb.y = add_sbyte(hb.y, v1, 3);
					ha.z = add_sbyte(ha.z, v2, 1), hb.z = add_sbyte(hb.z, v2, 3);
					ha.w = add_sbyte(ha.w, v3, 1), hb.w = add_sbyte(hb.w, v3, 3);
#else
					ha.x += prmt(v0, 0, 0x9991), hb.x += prmt(v0, 0, 0xbbb3);
					ha.y += prmt(v1, 0, 0x9991), hb.y += prmt(v1, 0, 0xbbb3);
					ha.z += prmt(v2, 0, 0x9991), hb.z += prmt(v2, 0, 0xbbb3);
					ha.w += prmt(v3, 0, 0x9991), hb.w += prmt(v3, 0, 0xbbb3);
#endif
					// keys: (score relative to the previous row's maximum) << 16 | priority of the column
					const uint32_t nb2 = (uint32_t)((0 - d) & 0xffff) * 0x10001u;
					const uint32_t p0 = vadd2(pk0, nb2), p1 = vadd2(pk1, nb2), p2 = vadd2(pk2, nb2), p3 = vadd2(pk3, nb2);
					const int k0 = (int)prmt(p0, (uint32_t)iaddmax((int)ha.x, nMprev, GD_KSW_REL_FLOOR), 0x5410);
					const int k4 = (int)prmt(p0, (uint32_t)iaddmax((int)hb.x, nMprev, GD_KSW_REL_FLOOR), 0x5432);
					const int k1 = (int)prmt(p1, (uint32_t)iaddmax((int)ha.y, nMprev, GD_KSW_REL_FLOOR), 0x5410);
					const int k5 = (int)prmt(p1, (uint32_t)iaddmax((int)hb.y, nMprev, GD_KSW_REL_FLOOR), 0x5432);
					const int k2 = (int)prmt(p2, (uint32_t)iaddmax((int)ha.z, nMprev, GD_KSW_REL_FLOOR), 0x5410);
					const int k6 = (int)prmt(p2, (uint32_t)iaddmax((int)hb.z, nMprev, GD_KSW_REL_FLOOR), 0x5432);
					const int k3 = (int)prmt(p3, (uint32_t)iaddmax((int)ha.w, nMprev, GD_KSW_REL_FLOOR), 0x5410);
					const int k7 = (int)prmt(p3, (uint32_t)iaddmax((int)hb.w, nMprev, GD_KSW_REL_FLOOR), 0x5432);
					if (core) {
						*(uint4 *)(rc + REC_H) = ha, *(uint4 *)(rc + REC_H + 16) = hb;
						run = imax3(run, k0, k4), run = imax3(run, k1, k5), run = imax3(run, k2, k6), run = imax3(run, k3, k7);
					}
				}
			}
		}
		Gang<G>::sync();
		// ---- phase C: score tracking ----
		int stop = 0;
		if (!EXACT) { // ksw2_extd2_sse.c:367-383 (every lane of the group tracks the same H0)
			// v[T] and u[T+1] of the updated row, T = last_H0_t (T >= st0-1 always holds; row 0 starts from H0 = -qe)
			const int T = H0_t;
			const int T1_off = ((T + 1) & 7) == 0 ? ring_fwd(T_off, REC, RB) : T_off;
			const int d0 = (int)*(const int8_t *)(ring + T_off + pos2(T & 7) + REC_A);        // v[T]
			const int d1 = (int)*(const int8_t *)(ring + T1_off + pos2((T + 1) & 7) + REC_B); // u[T+1]
			const bool in0 = T >= st0 && T <= en0, in1 = T + 1 >= st0 && T + 1 <= en0;
			const bool take_v = in0 && (!in1 || d0 > d1);
			if (active) {
				H0 += take_v ? d0 : d1;
				if (!take_v) ++H0_t, T_off = T1_off;
				if (MODE == 1) { // KSW_EZ_APPROX_DROP: ksw_apply_zdrop, ksw2.h:172-188
					if (H0 > res.max) res.max = H0, res.max_t = H0_t, res.max_q = r - H0_t;
					else if (H0_t >= res.max_t && r - H0_t >= res.max_q) {
						const int tl = H0_t - res.max_t, ql = (r - H0_t) - res.max_q, l = tl > ql ? tl - ql : ql - tl;
						if (C.zdrop >= 0 && res.max - H0 > C.zdrop + l * C.e2) res.zdropped = 1, stop = 1;
					}
				}
				if (!stop && r == nrows - 1 && en0 == tlen - 1) res.score = H0; // a Z-drop on the last row leaves before the score is set (ksw2_extd2_sse.c:380-382)
			}
		} else { // ksw2_extd2_sse.c:323-366
			// the cells kept out of the bulk update: scan tail (H += v) and column en0 (H[en0-1] + u[en0])
			int hn = 0;
			if (sp) {
				hn = sp_h + (int)(int8_t) * (sp_bp + (sp_en0 ? REC_B : REC_A));
				*sp_hp = hn;
				const int relc = imin(imax(hn - Mprev, GD_KSW_REL_FLOOR), 32767);
				const uint32_t pr = sp_en0 ? 0xffffu : (uint32_t)(3 - li) << 13 | (uint32_t)(GD_KSW_POS_MAX - sp_cc);
				run = imax(run, (int)((uint32_t)relc << 16 | pr));
			} else if (li == 3) hn = *sp_hp; // en0 == 0: column 0 was updated by the bulk pass
			const int He = hn; // lane 3 only: H[en0] of this row (mte and the final score are tracked by lane 3)
			run = Gang<G>::group_max(run, lane, scratch);
			Gang<G>::sync();
			int32_t *const hs_ptr = (int32_t *)(ring + st0_off + REC_H) + (st0 & 7);
			const int Hs = *hs_ptr; // H[st0] of this row
			if (active) {
				const int rel = run >> 16;
				const uint32_t pr = (uint32_t)run & 0xffffu;
				int max_H = Mprev + rel, max_t = pr == 0xffffu ? en0 : st + (GD_KSW_POS_MAX - (int)(pr & 0x1fffu));
				if (rel <= GD_KSW_REL_FLOOR || rel >= 32767 || C.force_slow_max)
					row_max_literal(ring, REC, NR, st_rec, st, st0, en0, max_H, max_t);
#ifdef GD_HOST_EMU
				{ // logic tests only: the key-based maximum must agree with the literal scan on every row
					int lit_H, lit_t;
					row_max_literal(ring, REC, NR, st_rec, st, st0, en0, lit_H, lit_t);
					if (lit_H != max_H || lit_t != max_t) ++emu::rowmax_mismatches();
				}
#endif
				Mprev = max_H, st0_prev = st0, Hs_prev = Hs, Hs_ptr = hs_ptr;
				if (en0 == tlen - 1 && He > res.mte) res.mte = He, res.mte_q = r - en;
				if (r - st0 == qlen - 1 && Hs > res.mqe) res.mqe = Hs, res.mqe_t = st0;
				if (max_H > res.max) res.max = max_H, res.max_t = max_t, res.max_q = r - max_t;
				else if (max_t >= res.max_t && r - max_t >= res.max_q) {
					const int tl = max_t - res.max_t, ql = (r - max_t) - res.max_q, l = tl > ql ? tl - ql : ql - tl;
					if (C.zdrop >= 0 && res.max - max_H > C.zdrop + l * C.e2) res.zdropped = 1, stop = 1;
				}
				if (!stop && r == nrows - 1 && en0 == tlen - 1) res.score = He;
			}
		}
		if (active) {
			++rows_exec;
			if (WITH_P) prow += ncol16;
			++r;
			if ((r & 7) == 0) r_off = ring_fwd(r_off, REC, RB);
			if (stop || r == nrows) finish = true;
		}
		// ================= pair finished: publish the record =================
		if (finish) {
			if (li == 3) { // every lane holds the same record; in exact mode lane 3 alone tracks mte and the final score
				res.rows_done = rows_exec;
				if (WITH_P) { // choice of the traceback start, ksw2_extd2_sse.c:389-400
					if (!res.zdropped && !(C.flag & KSW_F_EXTZ_ONLY)) res.tb_i = tlen - 1, res.tb_j = qlen - 1;
					else if (!res.zdropped && (C.flag & KSW_F_EXTZ_ONLY) && res.mqe + C.end_bonus > res.max)
						res.reach_end = 1, res.tb_i = res.mqe_t, res.tb_j = qlen - 1;
					else if (res.max_t >= 0 && res.max_q >= 0) res.tb_i = res.max_t, res.tb_j = res.max_q;
				}
				B.res[pair] = res;
			}
			have = false;
			qlen = tlen = 1, w = 0, r = 0; // idle geometry: one dummy cell per row until the next pair arrives
			if (!SEQ_SMEM) qsm = B.qpk; // any valid arena
			H0_t = 0, T_off = 0, r_off = 0;
		}
	}
}

// Stage one pair into the padded arenas the DP kernel reads (see KswBatch): target zero padded with
// every 8-column chunk in strided order, query reversed (qr[] of ksw2_extd2_sse.c:128) with
// N(4) -> 8 (ksw2_extd2_avx.c:187-190).
GD_DEV void ksw_pack_pair(const uint8_t *GD_RESTRICT q, int qlen, const uint8_t *GD_RESTRICT t, int tlen,
                          uint8_t *GD_RESTRICT tpk, int t_stride, uint8_t *GD_RESTRICT qpk, int q_stride, int lane,
                          int nlanes)
{
	for (int i = lane * 4; i < t_stride; i += nlanes * 4) {
		uint32_t wv = 0;
		for (int k = 0; k < 4; ++k) { // byte i+k of the arena holds column (chunk base) + j with chunk_pos(j) == (i+k)&7
			const int p = (i + k) & 7, j = (p >> 1) | ((p & 1) << 2), col = ((i + k) & ~7) + j;
			if (col < tlen) wv |= (uint32_t)t[col] << (8 * k);
		}
		*(uint32_t *)(tpk + i) = wv;
	}
	for (int i = lane * 4; i < q_stride; i += nlanes * 4) {
		uint32_t wv = 0;
		for (int k = 0; k < 4; ++k) {
			int j = i + k - GD_KSW_QFRONT;
			if (j >= 0 && j < qlen) {
				uint32_t c = q[qlen - 1 - j];
				wv |= (c == 4 ? 8u : c) << (8 * k);
			}
		}
		*(uint32_t *)(qpk + i) = wv;
	}
}

// ---------------------------------------------------------------------------------------------------
// The AVX-512 build's lead-in cells.  ksw_extd2_avx512 works on 64-cell vectors: the first vector of a row
// starts at stv = st0/64*64 and every lane of it is computed and stored (ksw2_extd2_avx.c:242,383,442-476),
// so the cells of [stv, st) -- st = st0/16*16, where the SSE build starts -- get state and backtrack bytes of
// their own, and off[r] = stv makes ksw_backtrack READ those bytes where the SSE build forces an insertion
// (ksw2.h:136).  No true cell ever depends on them (the boundary is blended into the lane of column st and
// last_st stays 16-aligned, :36,99,884), so scores, maxima and Z-drop are those of the 16-aligned sweep above,
// and a walk reaches them only by a deletion step off the band's left edge out of a cell that starts at a
// 16- but not 64-aligned column -- impossible with the presets (-q-2e < min score is needed), seen with random
// scoring and bands <= 3.  The walkers therefore only DETECT the case (one compare on the forced-insertion
// branch) and the few pairs that hit it are redone by ksw_lead64_pair: a literal cell-at-a-time model of the
// 64-lane rows (the lane-0 left neighbour of a first vector that starts left of st is byte 15 of its own
// 128-bit lane, i.e. column stv+15 of the previous row: index[0] = 15, :88-93), one thread block per pair,
// state in global scratch, followed by the walk over its own backtrack rows.
struct KswLead64 {
	const int32_t *list;  // [0] = count, [1..] = chunk-local pair numbers (KswBatch::lead64)
	const int64_t *qoff, *toff;
	const uint8_t *qbuf, *tbuf; // raw byte codes of the caller's batch
	uint8_t *scratch;     // per block: 10 * T64 state bytes, then the backtrack rows
	int64_t slot_bytes;
	int32_t T64, ncol64;  // launch-wide upper bounds the slots are sized for
	uint32_t *cigar;
	int32_t stride;
};

template <class BT> GD_DEV void ksw_lead64_report(const BT &B, KswResult *res, int lp)
{
	res->lead64 = 1;
	if (B.lead64) {
		const int k = atomic_add(B.lead64, 1);
		B.lead64[1 + k] = lp;
	}
}

struct CigOut { // run-length CIGAR writer of the walkers (ksw_push_cigar, ksw2.h:100-111), walk order
	uint32_t *cig;
	int stride, n, overflow;
	uint32_t cur;
	GD_MEM void push(uint32_t op, uint32_t len)
	{
		if (cur && (cur & 0xf) == op) cur += len << 4;
		else {
			flush();
			cur = len << 4 | op;
		}
	}
	GD_MEM void flush()
	{
		if (!cur) return;
		if (n < stride) cig[n] = cur;
		else overflow = 1;
		++n, cur = 0;
	}
};

// One block redoes one pair.  `tid`/`nt`: thread and block size; every thread owns the columns t with t % nt == tid.
GD_DEV void ksw_lead64_pair(const KswConsts &C, const KswBatch &B, const KswLead64 &L, int lp, uint8_t *scr, int tid, int nt)
{
	const int pair = B.base + lp;
	KswResult *res = &B.res[pair];
	const int qlen = B.qlen[pair], tlen = B.tlen[pair];
	int w = B.w ? B.w[pair] : B.w_all;
	if (w < 0) w = imax(tlen, qlen);
	const uint8_t *query = L.qbuf + L.qoff[pair], *target = L.tbuf + L.toff[pair];
	const int T64 = (tlen + 63) & ~63;
	const int ncol64 = ((imin(imin(qlen, tlen), w + 1) + 63) / 64 + 1) * 64; // ksw2_extd2_avx.c:145-146
	const bool right = (C.flag & KSW_F_RIGHT) != 0;
	// state: x, v, x2 of the previous and the current row (ping-pong: a cell reads its left neighbour's old values),
	// u, y, y2, s in place (same column, same thread)
	int8_t *xa = (int8_t *)scr, *va = xa + L.T64, *x2a = va + L.T64, *xb = x2a + L.T64, *vb = xb + L.T64, *x2b = vb + L.T64;
	int8_t *u = x2b + L.T64, *y = u + L.T64, *y2 = y + L.T64, *s = y2 + L.T64;
	uint8_t *p = (uint8_t *)(s + L.T64);
	const int8_t i1 = (int8_t)(-C.q - C.e), i2 = (int8_t)(-C.q2 - C.e2);
	const int8_t mch = (int8_t)hi8(C.MCH16), mis = (int8_t)(C.MIS4 & 0xff), scn = (int8_t)(C.SCN4 & 0xff);
	const int8_t q1 = (int8_t)C.q, q21 = (int8_t)C.q2, qe = (int8_t)(C.q + C.e), qe2 = (int8_t)(C.q2 + C.e2);
	for (int t = tid; t < T64; t += nt)
		xa[t] = va[t] = xb[t] = vb[t] = u[t] = y[t] = i1, x2a[t] = x2b[t] = y2[t] = i2, s[t] = 0;
	sync_block();
	const int nrows = res->rows_done; // rows the sweep executed (a Z-drop leaves the loop early)
	int last_st = -1, last_en = -1;
	for (int r = 0; r < nrows; ++r) {
		Bounds b;
		if (!row_bounds(r, qlen, tlen, w, b)) break;
		const int st = b.st, en = b.en, stv = b.st0 & ~63;
		const int gd = gap_delta(r, C);
		int8_t bx1 = i1, bx21 = i2, bv1 = i1; // ksw2_extd2_sse.c:149-159
		if (st > 0) {
			if (st - 1 >= last_st && st - 1 <= last_en) bx1 = xa[st - 1], bx21 = x2a[st - 1], bv1 = va[st - 1];
		} else bv1 = (int8_t)gd;
		const int ext_end = imin(b.st0 + (((b.en0 - b.st0) >> 4) + 1) * 16, T64); // the score row is rewritten on [st0, ext_end)
		for (int t = stv + ((tid - stv) % nt + nt) % nt; t <= imax(en, ext_end - 1); t += nt) {
			if (t >= b.st0 && t < ext_end) { // xor-table rule, ksw2_extd2_avx.c:187-208,312-313; zero padding outside the sequences
				const int qi = r - t;
				const uint32_t tc = t < tlen ? target[t] : 0, qc = (qi >= 0 && qi < qlen) ? query[qi] : 0;
				const uint32_t idx = tc ^ (qc == 4 ? 8u : qc);
				s[t] = (idx & 0x80) ? 0 : (idx & 15) == 0 ? mch : (idx & 15) < 4 ? mis : (idx & 15) < 13 ? scn : 0;
			}
			if (t > en) continue;
			int8_t x1, v1, x21;
			if (t == st) x1 = bx1, v1 = bv1, x21 = bx21;
			else if (t == stv) x1 = xa[stv + 15], v1 = va[stv + 15], x21 = x2a[stv + 15];
			else x1 = xa[t - 1], v1 = va[t - 1], x21 = x2a[t - 1];
			int8_t yt = y[t], y2t = y2[t], ut = u[t];
			if (t == r && en >= r) yt = i1, y2t = i2, ut = (int8_t)gd; // ksw2_extd2_sse.c:160-163
			int8_t z = s[t], a = (int8_t)(x1 + v1), bb = (int8_t)(yt + ut), a2 = (int8_t)(x21 + v1), b2 = (int8_t)(y2t + ut);
			uint32_t d = 0;
			if (!right) {
				if (a > z) d = 1, z = a;
				if (bb > z) d = 2, z = bb;
				if (a2 > z) d = 3, z = a2;
				if (b2 > z) d = 4, z = b2;
			} else {
				if (a >= z) d = 1, z = a;
				if (bb >= z) d = 2, z = bb;
				if (a2 >= z) d = 3, z = a2;
				if (b2 >= z) d = 4, z = b2;
			}
			if (z > mch) z = mch;
			u[t] = (int8_t)(z - v1), vb[t] = (int8_t)(z - ut);
			const int8_t tq = (int8_t)(z - q1), tq2 = (int8_t)(z - q21);
			a = (int8_t)(a - tq), bb = (int8_t)(bb - tq), a2 = (int8_t)(a2 - tq2), b2 = (int8_t)(b2 - tq2);
			if (!right) {
				if (a > 0) d |= 0x08; else a = 0;
				if (bb > 0) d |= 0x10; else bb = 0;
				if (a2 > 0) d |= 0x20; else a2 = 0;
				if (b2 > 0) d |= 0x40; else b2 = 0;
			} else {
				if (a >= 0) d |= 0x08; else a = 0;
				if (bb >= 0) d |= 0x10; else bb = 0;
				if (a2 >= 0) d |= 0x20; else a2 = 0;
				if (b2 >= 0) d |= 0x40; else b2 = 0;
			}
			xb[t] = (int8_t)(a - qe), y[t] = (int8_t)(bb - qe), x2b[t] = (int8_t)(a2 - qe2), y2[t] = (int8_t)(b2 - qe2);
			p[(size_t)r * ncol64 + (t - stv)] = (uint8_t)d;
		}
		sync_block();
		int8_t *sw;
		sw = xa, xa = xb, xb = sw, sw = va, va = vb, vb = sw, sw = x2a, x2a = x2b, x2b = sw;
		last_st = st, last_en = en;
	}
	if (tid != 0) return;
	// the walk (ksw_backtrack, ksw2.h:115-163) with off[r] = stv, off_end[r] = en (ksw2_extd2_avx.c:442,524)
	CigOut co;
	co.cig = L.cigar + (size_t)pair * L.stride, co.stride = L.stride, co.n = 0, co.overflow = 0, co.cur = 0;
	int i = res->tb_i, j = res->tb_j, state = 0;
	while (i >= 0 && j >= 0) {
		const int r = i + j;
		int force = -1;
		Bounds bd;
		row_bounds(r, qlen, tlen, w, bd);
		const int stv = bd.st0 & ~63;
		if (i < stv) force = 2;
		if (i > bd.en) force = 1;
		const uint32_t cell = force < 0 ? p[(size_t)r * ncol64 + (i - stv)] : 0;
		if (state == 0) state = cell & 7;
		else if (!((cell >> (state + 2)) & 1)) state = 0;
		if (state == 0) state = cell & 7;
		if (force >= 0) state = force;
		if (state == 0) co.push(0, 1), --i, --j;
		else if (state == 1 || state == 3) co.push(2, 1), --i;
		else co.push(1, 1), --j;
	}
	if (i >= 0) co.push(2, (uint32_t)(i + 1));
	if (j >= 0) co.push(1, (uint32_t)(j + 1));
	co.flush();
	if (!co.overflow && !(C.flag & KSW_F_REV_CIGAR))
		for (int k = 0; k < co.n >> 1; ++k) {
			const uint32_t t = co.cig[k];
			co.cig[k] = co.cig[co.n - 1 - k], co.cig[co.n - 1 - k] = t;
		}
	res->n_cigar = co.overflow ? -co.n : co.n;
}

// One thread per pair: walk the backtrack matrix (ksw_backtrack, ksw2.h:115-163 with is_rot=1)
// and emit the run-length CIGAR in walk order (i.e. reversed) into cig_tmp[pair*stride ...].
GD_DEV void ksw_traceback_one(const KswBatch &B, int flag, int lp, uint32_t *cig_out, int stride)
{
	const int pair = B.base + lp;
	KswResult *res = &B.res[pair];
	int i = res->tb_i, j = res->tb_j, n = 0, state = 0, overflow = 0, lead64 = 0;
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
		if (i < bd.st) force = 2, lead64 |= i >= (bd.st0 & ~63);
		if (i > bd.en) force = 1;
		const int cc = i - bd.st;
		uint32_t cell = force < 0 ? p[(size_t)r * ncol16 + (cc & ~7) + chunk_pos(cc & 7)] : 0;
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
	if (lead64) ksw_lead64_report(B, res, lp);
}

// Same walk for long pairs, one WARP per pair: the backtrack bytes the path can touch next (32 rows x 80
// columns ending at the current cell -- a step moves up one or two rows and at most one column to the
// left) are fetched into shared memory with one coalesced 8-byte load per lane and chunk, then every lane
// walks the tile redundantly (identical state, broadcast reads) and lane 0 writes the CIGAR.  This turns
// ~10^5 dependent global loads per ONT-sized pair into ~10^5 / 24 tile fills.
#define GD_KSW_TB_ROWS 32
#define GD_KSW_TB_CHUNKS 10
GD_DEV void ksw_traceback_warp(const KswBatch &B, int flag, int lp, uint32_t *cig_out, int stride, uint2 *tile, int lane)
{
	const int pair = B.base + lp;
	KswResult *res = &B.res[pair];
	int i = res->tb_i, j = res->tb_j, n = 0, state = 0, overflow = 0, lead64 = 0;
	if (i < 0 || j < 0) {
		if (lane == 0) res->n_cigar = 0;
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
		// ---- fill the tile: rows r_top .. r_top-31, columns [wlo, wlo+80) ----
		const int r_top = i + j, r_bot = imax(r_top - (GD_KSW_TB_ROWS - 1), 0);
		const int wlo = ((i >> 3) - (GD_KSW_TB_CHUNKS - 1)) * 8; // may be negative: those chunks are outside every row
		sync_warp(0xffffffffu);                                   // the previous tile's readers are done
		{
			const int rr = r_top - lane;
			Bounds rb;
			rb.st = 0, rb.en = -1;
			if (rr >= 0) row_bounds(rr, qlen, tlen, w, rb);
			for (int c = 0; c < GD_KSW_TB_CHUNKS; ++c) {
				const int cb = wlo + 8 * c;
				uint2 v;
				v.x = v.y = 0;
				if (rr >= 0 && cb >= rb.st && cb <= rb.en) v = *(const uint2 *)(p + (size_t)rr * ncol16 + (cb - rb.st));
				tile[lane * GD_KSW_TB_CHUNKS + c] = v;
			}
		}
		sync_warp(0xffffffffu);
		// ---- walk while the path stays inside the tile's rows ----
		while (i >= 0 && j >= 0 && i + j >= r_bot) {
			const int r = i + j;
			int force = -1;
			Bounds bd;
			row_bounds(r, qlen, tlen, w, bd);
			if (i < bd.st) force = 2, lead64 |= i >= (bd.st0 & ~63);
			if (i > bd.en) force = 1;
			const int cc = i - wlo; // 0..79
			const uint8_t *row = (const uint8_t *)(tile + (r_top - r) * GD_KSW_TB_CHUNKS);
			uint32_t cell = force < 0 ? row[(cc & ~7) + chunk_pos(cc & 7)] : 0;
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
					if (n < stride) {
						if (lane == 0) cig[n] = cur;
					} else overflow = 1;
					++n;
				}
				cur = 16 | op;
			}
		}
	}
	if (i >= 0) { // leading deletion
		if (cur && (cur & 0xf) == 2) cur += (uint32_t)(i + 1) << 4;
		else {
			if (cur) {
				if (n < stride) {
					if (lane == 0) cig[n] = cur;
				} else overflow = 1;
				++n;
			}
			cur = (uint32_t)(i + 1) << 4 | 2;
		}
	}
	if (j >= 0) { // leading insertion
		if (cur && (cur & 0xf) == 1) cur += (uint32_t)(j + 1) << 4;
		else {
			if (cur) {
				if (n < stride) {
					if (lane == 0) cig[n] = cur;
				} else overflow = 1;
				++n;
			}
			cur = (uint32_t)(j + 1) << 4 | 1;
		}
	}
	if (cur) {
		if (n < stride) {
			if (lane == 0) cig[n] = cur;
		} else overflow = 1;
		++n;
	}
	sync_warp(0xffffffffu);
	if (!overflow && !(flag & KSW_F_REV_CIGAR)) // ksw2.h:158-160: forward order unless REV_CIGAR (all lanes swap disjoint entries)
		for (int k = lane; k < n >> 1; k += 32) {
			uint32_t t = cig[k];
			cig[k] = cig[n - 1 - k], cig[n - 1 - k] = t;
		}
	if (lane == 0) res->n_cigar = overflow ? -n : n;
	if (lane == 0 && lead64) ksw_lead64_report(B, res, lp);
}

} // namespace gd
