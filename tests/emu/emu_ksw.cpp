// tests/emu/emu_ksw.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs the *device code* of genome-on-diet_b200/csrc/gd_ksw.cuh under the fiber SIMT emulator so
// its logic can be compared with the oracle on a machine without a GPU.
#define GD_HOST_EMU 1
#include "simt_emu.h"
#include "gd_ksw_host.h"
#include <algorithm>
#include <vector>

using namespace gd;

template <int G, bool RIGHT, int MODE, bool WITH_P>
static void run_dp(const KswConsts &C, const KswBatch &B, int threads)
{
	if (G > 32) threads = G; // block-per-pair: one block of G threads, two blocks so that the ticket is shared
	int groups = G > 32 ? 1 : threads / G;
	emu::launch(G > 32 ? 2 : 1, threads, GD_KSW_LUT_BYTES + (size_t)groups * B.group_smem, [&]() {
		int tid = emu::thread_idx();
		uint8_t *sm = (uint8_t *)emu::smem();
		ksw_build_lut(sm, tid, threads);
		emu::sync_block();
		if (G <= 32)
			ksw_warp_body<G, RIGHT, MODE, WITH_P>(C, B, sm + GD_KSW_LUT_BYTES + (size_t)(tid >> 5) * (32 / (G <= 32 ? G : 32)) * B.group_smem,
			                                      sm, tid & 31);
		else ksw_warp_body<G, RIGHT, MODE, WITH_P>(C, B, sm + GD_KSW_LUT_BYTES, sm, tid);
	});
}

template <int G, bool RIGHT>
static void dispatch2(const KswConsts &C, const KswBatch &B, int threads, int mode, bool with_p)
{
	if (mode == 0) with_p ? run_dp<G, RIGHT, 0, true>(C, B, threads) : run_dp<G, RIGHT, 0, false>(C, B, threads);
	else if (mode == 1) with_p ? run_dp<G, RIGHT, 1, true>(C, B, threads) : run_dp<G, RIGHT, 1, false>(C, B, threads);
	else with_p ? run_dp<G, RIGHT, 2, true>(C, B, threads) : run_dp<G, RIGHT, 2, false>(C, B, threads);
}
template <int G>
static void dispatch(const KswConsts &C, const KswBatch &B, int threads, bool right, bool exact, bool with_p)
{
	const int mode = exact ? 2 : (C.flag & KSW_F_APPROX_DROP) ? 1 : 0;
	if (right) dispatch2<G, true>(C, B, threads, mode, with_p);
	else dispatch2<G, false>(C, B, threads, mode, with_p);
}

extern "C" int emu_ksw_batch(int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf, const int32_t *tlen,
                             const int64_t *toff, const uint8_t *tbuf, const int32_t *w, int m, const int8_t *mat,
                             int q, int e, int q2, int e2, int zdrop, int end_bonus, int flag, int G, int threads,
                             KswResult *res, uint32_t *cigar, int cigar_stride, int *lead64_count)
{
	KswConsts C = ksw_make_consts(m, mat, q, e, q2, e2, zdrop, end_bonus, flag & 0xff);
	C.force_slow_max = (flag >> 8) & 1; // test hook: bit 8 of the emulator's flag argument
	flag &= 0xff;
	int max_q = 1, max_t = 1, max_w = 0;
	for (int i = 0; i < n; ++i) {
		int ww = w[i] < 0 ? (tlen[i] > qlen[i] ? tlen[i] : qlen[i]) : w[i];
		if (qlen[i] > max_q) max_q = qlen[i];
		if (tlen[i] > max_t) max_t = tlen[i];
		if (ww > max_w) max_w = ww;
	}
	const bool exact = !(flag & KSW_F_APPROX_MAX), with_p = !(flag & KSW_F_SCORE_ONLY), right = (flag & KSW_F_RIGHT) != 0;
	KswGeom g = ksw_geometry(max_q, max_t, max_w, exact, with_p, G);
	std::vector<uint8_t> tpk((size_t)n * g.t_stride), qpk((size_t)n * g.q_stride), p((size_t)n * g.p_stride + 16);
	memset(p.data(), 0xAA, p.size()); // poison: any read of an unwritten backtrack byte shows up
	for (int i = 0; i < n; ++i)
		ksw_pack_pair(qbuf + qoff[i], qlen[i], tbuf + toff[i], tlen[i], tpk.data() + (size_t)i * g.t_stride, g.t_stride,
		              qpk.data() + (size_t)i * g.q_stride, g.q_stride, 0, 1);
	int ticket = 0;
	KswHot hot = ksw_hot_from_consts(C);
	KswBatch B;
	B.hot = &hot;
	B.n = n, B.base = 0, B.qlen = qlen, B.tlen = tlen, B.w = w, B.w_all = 0;
	B.tpk = tpk.data(), B.qpk = qpk.data(), B.t_stride = g.t_stride, B.q_stride = g.q_stride;
	B.p = p.data(), B.p_stride = g.p_stride, B.res = res, B.ticket = &ticket, B.ring = g.ring, B.group_smem = g.group_smem;
	emu::rowmax_mismatches() = 0;
	switch (G) {
	case 4: dispatch<4>(C, B, threads, right, exact, with_p); break;
	case 8: dispatch<8>(C, B, threads, right, exact, with_p); break;
	case 16: dispatch<16>(C, B, threads, right, exact, with_p); break;
	case 32: dispatch<32>(C, B, threads, right, exact, with_p); break;
	case 64: dispatch<64>(C, B, threads, right, exact, with_p); break;
	case 128: dispatch<128>(C, B, threads, right, exact, with_p); break;
	default: return -1;
	}
	std::vector<int32_t> l64(n + 4, 0);
	B.lead64 = l64.data();
	if (with_p) {
		for (int i = 0; i < n; ++i) ksw_traceback_one(B, flag, i, cigar, cigar_stride);
		// the pairs whose walk entered the AVX-512 lead-in cells: the slow model, one 64-thread block per pair
		KswLead64 L;
		L.list = l64.data(), L.qoff = qoff, L.toff = toff, L.qbuf = qbuf, L.tbuf = tbuf;
		L.T64 = (max_t + 63) / 64 * 64;
		L.ncol64 = ((std::min(std::min(max_q, max_t), max_w + 1) + 63) / 64 + 1) * 64;
		L.slot_bytes = (int64_t)10 * L.T64 + (int64_t)(max_q + max_t - 1) * L.ncol64;
		L.cigar = cigar, L.stride = cigar_stride;
		std::vector<uint8_t> scr((size_t)L.slot_bytes, 0x55);
		L.scratch = scr.data();
		for (int k = 0; k < l64[0]; ++k)
			emu::launch(1, 64, 0, [&]() { ksw_lead64_pair(C, B, L, l64[1 + k], scr.data(), emu::thread_idx(), 64); });
	}
	if (lead64_count) *lead64_count = l64[0];
	if (emu::rowmax_mismatches()) return -2; // fast row maximum disagreed with the literal scan
	return 0;
}

// != 0: every later launch of this library keeps all blocks resident and runs blocks and threads in pseudo-random orders
extern "C" void emu_set_shuffle(unsigned seed) { emu::shuffle_seed() = seed; }
