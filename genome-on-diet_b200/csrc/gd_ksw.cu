// gd_ksw.cu -- kernels and launcher of the batched ksw_extd2 DP (see gd_ksw.cuh for the design).
#include "gd_ctx.h"
#include <mutex>
#include <map>
#include <tuple>
#include <utility>
#include "gd_ksw_host.h"
#include <algorithm>

using namespace gd;

static_assert(sizeof(KswResult) == sizeof(gd_extz_t), "result layouts must match");

// --------------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------------
// G <= 32: every warp of the block carries 32/G pairs; G = 64 / 128: block-per-pair (blockDim.x == G)
template <int G, bool RIGHT, int MODE, bool WITH_P>
__global__ void __launch_bounds__(128) gd_ksw_dp_kernel(const KswConsts C, const KswBatch B)
{
	extern __shared__ __align__(128) uint8_t gd_smem[];
	const int tid = threadIdx.x;
	ksw_build_lut(gd_smem, tid, blockDim.x);
	__syncthreads();
	if (G <= 32) {
		uint8_t *warp_smem = gd_smem + GD_KSW_LUT_BYTES + (size_t)(tid >> 5) * (32 / (G <= 32 ? G : 32)) * B.group_smem;
		ksw_warp_body<G, RIGHT, MODE, WITH_P>(C, B, warp_smem, gd_smem, tid & 31);
	} else ksw_warp_body<G, RIGHT, MODE, WITH_P>(C, B, gd_smem + GD_KSW_LUT_BYTES, gd_smem, tid);
}

// one warp per pair: raw byte codes -> padded arenas
__global__ void __launch_bounds__(128)
    gd_ksw_pack_kernel(int n, int base, const int32_t *__restrict__ qlen, const int64_t *__restrict__ qoff,
                       const uint8_t *__restrict__ qbuf, const int32_t *__restrict__ tlen,
                       const int64_t *__restrict__ toff, const uint8_t *__restrict__ tbuf, uint8_t *tpk, int t_stride,
                       uint8_t *qpk, int q_stride, const KswConsts C, KswHot *hot)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) *hot = ksw_hot_from_consts(C);
	const int warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
	for (int lp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; lp < n; lp += warps) {
		const int pair = base + lp;
		int ql = qlen[pair], tl = tlen[pair];
		if (ql < 0) ql = 0;
		if (tl < 0) tl = 0;
		ksw_pack_pair(qbuf + qoff[pair], ql, tbuf + toff[pair], tl, tpk + (size_t)lp * t_stride, t_stride,
		              qpk + (size_t)lp * q_stride, q_stride, lane, 32);
	}
}

__global__ void __launch_bounds__(128) gd_ksw_traceback_kernel(const KswBatch B, int flag, uint32_t *cigar, int stride)
{
	const int lp = blockIdx.x * blockDim.x + threadIdx.x;
	if (lp < B.n) ksw_traceback_one(B, flag, lp, cigar, stride);
}

// long pairs: one warp per pair, backtrack bytes staged through shared memory tiles
__global__ void __launch_bounds__(128) gd_ksw_traceback_warp_kernel(const KswBatch B, int flag, uint32_t *cigar, int stride)
{
	__shared__ uint2 tiles[4][GD_KSW_TB_ROWS * GD_KSW_TB_CHUNKS];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int lp = blockIdx.x * 4 + wid;
	if (lp < B.n) ksw_traceback_warp(B, flag, lp, cigar, stride, tiles[wid], lane);
}

// the pairs whose walk entered the AVX-512 build's lead-in cells (normally none): one block per pair, see ksw_lead64_pair
__global__ void __launch_bounds__(128) gd_ksw_lead64_kernel(const KswConsts C, const KswBatch B, const KswLead64 L)
{
	const int cnt = L.list[0];
	for (int k = blockIdx.x; k < cnt; k += gridDim.x) {
		ksw_lead64_pair(C, B, L, L.list[1 + k], L.scratch + (size_t)blockIdx.x * L.slot_bytes, threadIdx.x, blockDim.x);
		__syncthreads();
	}
}

__global__ void gd_ksw_nocigar_kernel(int n, KswResult *res)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) res[i].n_cigar = 0;
}

// exclusive scan of max(n_cigar,0) over n pairs (single block; n_cigar arrays are small)
// off[i] = *run_base + (sum of earlier counts); off[n] and *run_base receive the end, *h_end too (pinned host word)
__global__ void __launch_bounds__(1024)
    gd_ksw_cigar_scan_kernel(int n, const KswResult *res, int64_t *off, int64_t *run_base, int64_t *h_end)
{ // h_end[0] receives the end position; h_end[-1]... is not touched; a negative n_cigar (stride overflow) makes the end negative
	__shared__ int64_t warp_sum[32];
	__shared__ int64_t carry;
	const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
	int bad = 0;
	if (tid == 0) carry = run_base ? *run_base : 0;
	__syncthreads();
	for (int base = 0; base < n; base += 1024) {
		const int i = base + tid;
		int64_t v = 0;
		if (i < n) {
			int c = res[i].n_cigar;
			v = c > 0 ? c : 0;
			if (c < 0) bad = 1;
		}
		int64_t inc = v;
		for (int d = 1; d < 32; d <<= 1) {
			int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += o;
		}
		if (lane == 31) warp_sum[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			int64_t ws = warp_sum[lane], wi = ws;
			for (int d = 1; d < 32; d <<= 1) {
				int64_t o = __shfl_up_sync(0xffffffffu, wi, d);
				if (lane >= d) wi += o;
			}
			warp_sum[lane] = wi - ws; // exclusive
		}
		__syncthreads();
		const int64_t excl = carry + warp_sum[wid] + inc - v;
		if (i < n) off[i] = excl;
		__syncthreads();
		if (tid == 1023) carry = excl + v;
		__syncthreads();
	}
	bad = __syncthreads_or(bad);
	if (tid == 0) {
		off[n] = carry;
		if (run_base) *run_base = carry;
		if (h_end) *h_end = bad ? -1 - carry : carry; // negative: some pair overflowed its CIGAR stride
	}
}

__global__ void __launch_bounds__(256)
    gd_ksw_cigar_gather_kernel(int n, const KswResult *res, const uint32_t *cigar, int stride, const int64_t *off,
                               uint32_t *out, int64_t cap)
{
	const int warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const int c = res[i].n_cigar;
		const int64_t o = off[i];
		if (c <= 0 || o + c > cap) continue;
		for (int k = lane; k < c; k += 32) out[o + k] = cigar[(size_t)i * stride + k];
	}
}

__global__ void gd_exact_match_kernel(int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                      const int64_t *toff, const uint8_t *tbuf, uint8_t *equal)
{ // exact_match_sse.c:27-88 == memcmp over qlen bytes; one warp per pair
	const int warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const int ql = qlen[i];
		const uint8_t *q = qbuf + qoff[i], *t = tbuf + toff[i];
		int diff = 0;
		for (int k = lane; k < ql; k += 32) diff |= q[k] != t[k];
		diff = __any_sync(0xffffffffu, diff);
		if (lane == 0) equal[i] = (ql > 0 && !diff) ? 1 : 0;
	}
}

// --------------------------------------------------------------------------------------------
// launcher
// --------------------------------------------------------------------------------------------
typedef void (*dp_kernel_t)(const KswConsts, const KswBatch);

template <int G, bool RIGHT> static dp_kernel_t pick_mode2(int mode, bool with_p)
{
	switch (mode) {
	case 0: return with_p ? gd_ksw_dp_kernel<G, RIGHT, 0, true> : gd_ksw_dp_kernel<G, RIGHT, 0, false>;
	case 1: return with_p ? gd_ksw_dp_kernel<G, RIGHT, 1, true> : gd_ksw_dp_kernel<G, RIGHT, 1, false>;
	default: return with_p ? gd_ksw_dp_kernel<G, RIGHT, 2, true> : gd_ksw_dp_kernel<G, RIGHT, 2, false>;
	}
}
template <int G> static dp_kernel_t pick_mode(bool right, int mode, bool with_p)
{
	return right ? pick_mode2<G, true>(mode, with_p) : pick_mode2<G, false>(mode, with_p);
}
static dp_kernel_t pick_kernel(int G, bool right, int mode, bool with_p)
{
	switch (G) {
	case 4: return pick_mode<4>(right, mode, with_p);
	case 8: return pick_mode<8>(right, mode, with_p);
	case 16: return pick_mode<16>(right, mode, with_p);
	case 32: return pick_mode<32>(right, mode, with_p);
	case 64: return pick_mode<64>(right, mode, with_p);
	default: return pick_mode<128>(right, mode, with_p);
	}
}

int gd_ksw_run_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff, const uint8_t *d_qbuf,
                      const int32_t *d_tlen, const int64_t *d_toff, const uint8_t *d_tbuf, const int32_t *d_w,
                      int w_all, int max_qlen, int max_tlen, int max_w, const gd_ksw_params_t *prm, gd_extz_t *d_ez,
                      uint32_t *d_cigar, int cigar_stride)
{
	if (n <= 0) return GD_OK;
	if (!prm || !d_ez) {
		ctx->err = "gd_ksw: null params / result pointer";
		return GD_ERR_ARG;
	}
	if (prm->flag & KSW_F_GENERIC_SC) {
		ctx->err = "gd_ksw: KSW_EZ_GENERIC_SC is not supported (never used by the Genome-on-Diet call sites)";
		return GD_ERR_ARG;
	}
	const int flag = prm->flag;
	const bool exact = !(flag & KSW_F_APPROX_MAX), right = (flag & KSW_F_RIGHT) != 0;
	const bool with_p = !(flag & KSW_F_SCORE_ONLY) && d_cigar != nullptr && cigar_stride > 0;
	KswConsts C = ksw_make_consts(prm->m, prm->mat, prm->q, prm->e, prm->q2, prm->e2, prm->zdrop, prm->end_bonus, flag);
	if (max_w < 0) max_w = std::max(max_qlen, max_tlen);
	int G = (int)ctx->opt_ksw_group;
	if (G != 4 && G != 8 && G != 16 && G != 32 && G != 64 && G != 128) G = ksw_pick_group(max_qlen, max_tlen, max_w, exact, n, ctx->sms);
	KswGeom geo = ksw_geometry(max_qlen, max_tlen, max_w, exact, with_p, G);
	if (geo.ring > GD_KSW_POS_MAX - 32 && exact) {
		ctx->err = "gd_ksw: band too wide for the exact-max keys (ring > 8158 columns)";
		return GD_ERR_ARG;
	}
	// block shape: 1..4 warps per block, whichever packs the most resident warps into the SM's shared
	// memory (each block also pays for the lookup tables and the driver's 1 KB); when even one warp's
	// worth of groups does not fit, widen the group (fewer pairs per warp)
	int threads = 0;
	for (;;) {
		if (G > 32) { // block-per-pair
			if (GD_KSW_LUT_BYTES + (size_t)geo.group_smem <= ctx->smem_optin) threads = G;
		} else {
			const size_t per_warp = (size_t)(32 / G) * geo.group_smem;
			int best_warps = 0;
			for (int wpb = 1; wpb <= 4; ++wpb) {
				const size_t blk = GD_KSW_LUT_BYTES + wpb * per_warp;
				if (blk > ctx->smem_optin) break;
				const int resident = (int)std::min<size_t>(32, ctx->smem_per_sm / (blk + 1024)) * wpb;
				if (resident >= best_warps) best_warps = resident, threads = wpb * 32;
			}
		}
		if (threads) break;
		if (G < 128) {
			G <<= 1;
			geo = ksw_geometry(max_qlen, max_tlen, max_w, exact, with_p, G);
			continue;
		}
		ctx->err = "gd_ksw: band too wide for the shared-memory column ring (needs > 227 KB per pair)";
		return GD_ERR_ARG;
	}
	const int groups_per_block = G > 32 ? 1 : threads / G;
	const size_t smem = GD_KSW_LUT_BYTES + (size_t)groups_per_block * geo.group_smem;
	const int mode = exact ? 2 : (flag & KSW_F_APPROX_DROP) ? 1 : 0;
	dp_kernel_t kern = pick_kernel(G, right, mode, with_p);
	if (smem > ctx->smem_optin) {
		ctx->err = "ksw_extd2: the column ring of this shape does not fit the shared memory of one block";
		return GD_ERR_ARG;
	}
	{ // The kernel's dynamic shared-memory limit only ever grows (a high-water mark per device and kernel, under a lock):
	  // host threads with their own contexts launch the same kernel with different sizes (one call per candidate in
	  // the drop-in path), and a per-call setting by one thread could be lowered by another between its set and its launch.
		static std::mutex mu;
		static std::map<std::pair<int, const void *>, size_t> high;
		std::lock_guard<std::mutex> lk(mu);
		size_t &h = high[std::make_pair(ctx->device, (const void *)kern)];
		if (smem > h) {
			GD_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			h = smem;
		}
	}
	int occ = 0;
	{ // resident blocks per SM of this (kernel, block shape, shared memory): queried once, the launcher runs per batch
		static std::mutex mu;
		static std::map<std::tuple<int, const void *, int, size_t>, int> cache;
		std::lock_guard<std::mutex> lk(mu);
		const auto key = std::make_tuple(ctx->device, (const void *)kern, threads, smem);
		auto it = cache.find(key);
		if (it == cache.end()) {
			GD_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
			cache[key] = occ;
		} else occ = it->second;
	}
	if (occ < 1) occ = 1;
	if (ctx->opt_ksw_blocks_per_sm > 0) occ = std::min<long>(occ, ctx->opt_ksw_blocks_per_sm);

	// chunking: the backtrack arena is the only large scratch
	size_t budget;
	if (ctx->opt_p_budget_mb > 0) budget = (size_t)ctx->opt_p_budget_mb << 20;
	else if (!with_p || (size_t)n * (size_t)std::max<int64_t>(geo.p_stride, 1) + 64 <= ctx->parena.cap) budget = ctx->parena.cap; // fits as is
	else { // the arena has to grow: size it from the memory that is free now (cudaMemGetInfo is not free, so only here)
		size_t fr = 0, tot = 0;
		GD_CUDA_OK(ctx, cudaMemGetInfo(&fr, &tot));
		// (a call of the mapping stage runs its slices on two contexts: each may take at most 40 % of what is free, and never
		// more than 64 GB -- the first lane used to leave too little for the second: out of memory at 20,000 HiFi reads)
		budget = std::max<size_t>(ctx->parena.cap, (size_t)(fr * 0.4));
		budget = std::min<size_t>(budget, (size_t)64 << 30);
	}
	int chunk = n;
	if (with_p) {
		size_t fit = std::max<size_t>(1, budget / (size_t)std::max<int64_t>(geo.p_stride, 1));
		chunk = (int)std::min<size_t>(fit, (size_t)n);
	}
	chunk = std::min(chunk, 1 << 22);
	int rc;
	if ((rc = gd_reserve(ctx, ctx->tpk, (size_t)chunk * geo.t_stride + 64))) return rc;
	if ((rc = gd_reserve(ctx, ctx->qpk, (size_t)chunk * geo.q_stride + 64))) return rc;
	if (with_p && (rc = gd_reserve(ctx, ctx->parena, (size_t)chunk * geo.p_stride + 64))) return rc;
	if ((rc = gd_reserve(ctx, ctx->ticket, 256))) return rc;
	// scratch of the lead-in model (gd_ksw_lead64_kernel): a few block slots, each the state columns + the 64-lane backtrack rows of one pair
	KswLead64 L64;
	int l64_blocks = 0;
	if (with_p) {
		const int mq = std::max(max_qlen, 1), mt = std::max(max_tlen, 1);
		L64.T64 = (mt + 63) / 64 * 64;
		L64.ncol64 = ((std::min(std::min(mq, mt), max_w + 1) + 63) / 64 + 1) * 64;
		L64.slot_bytes = ((int64_t)10 * L64.T64 + (int64_t)(mq + mt - 1) * L64.ncol64 + 255) / 256 * 256;
		l64_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(std::min(chunk, ctx->sms), ((int64_t)128 << 20) / L64.slot_bytes));
		if ((rc = gd_reserve(ctx, ctx->lead64_scr, (size_t)l64_blocks * L64.slot_bytes))) return rc;
		if ((rc = gd_reserve(ctx, ctx->lead64_list, ((size_t)chunk + 4) * sizeof(int32_t)))) return rc;
		L64.list = (const int32_t *)ctx->lead64_list.p, L64.scratch = (uint8_t *)ctx->lead64_scr.p;
		L64.qoff = d_qoff, L64.toff = d_toff, L64.qbuf = d_qbuf, L64.tbuf = d_tbuf;
		L64.cigar = d_cigar, L64.stride = cigar_stride;
	}

	ctx->stat_ksw_ring = geo.ring, ctx->stat_ksw_group = G, ctx->stat_ksw_chunks = 0;
	cudaStream_t s = ctx->stream;
	for (int base = 0; base < n; base += chunk) {
		const int cn = std::min(chunk, n - base);
		KswBatch B;
		B.n = cn, B.base = base, B.qlen = d_qlen, B.tlen = d_tlen, B.w = d_w, B.w_all = w_all;
		B.tpk = (const uint8_t *)ctx->tpk.p, B.qpk = (const uint8_t *)ctx->qpk.p;
		B.t_stride = geo.t_stride, B.q_stride = geo.q_stride;
		B.p = (uint8_t *)ctx->parena.p, B.p_stride = geo.p_stride;
		B.res = (KswResult *)d_ez, B.ticket = (int32_t *)ctx->ticket.p;
		B.hot = (const KswHot *)((uint8_t *)ctx->ticket.p + 64);
		B.ring = geo.ring, B.group_smem = geo.group_smem;
		B.lead64 = with_p ? (int32_t *)ctx->lead64_list.p : nullptr;

		GD_CUDA_OK(ctx, cudaMemsetAsync(ctx->ticket.p, 0, 4, s));
		if (with_p) GD_CUDA_OK(ctx, cudaMemsetAsync(ctx->lead64_list.p, 0, 4, s));
		{
			int blocks = std::min((cn + 3) / 4, ctx->sms * 16);
			gd_ksw_pack_kernel<<<blocks, 128, 0, s>>>(cn, base, d_qlen, d_qoff, d_qbuf, d_tlen, d_toff, d_tbuf,
			                                          (uint8_t *)ctx->tpk.p, geo.t_stride, (uint8_t *)ctx->qpk.p,
			                                          geo.q_stride, C, (KswHot *)((uint8_t *)ctx->ticket.p + 64));
		}
		{
			int blocks = std::min((cn + groups_per_block - 1) / groups_per_block, ctx->sms * occ);
			GdKernelTimer tm(ctx, &ctx->tm_dp);
			kern<<<blocks, threads, smem, s>>>(C, B);
		}
		if (with_p) {
			if (max_qlen + max_tlen >= 2048) gd_ksw_traceback_warp_kernel<<<(cn + 3) / 4, 128, 0, s>>>(B, flag, d_cigar, cigar_stride);
			else gd_ksw_traceback_kernel<<<(cn + 127) / 128, 128, 0, s>>>(B, flag, d_cigar, cigar_stride);
			gd_ksw_lead64_kernel<<<l64_blocks, 128, 0, s>>>(C, B, L64); // blocks leave at once unless a walk reported a pair
		}
		ctx->stat_launches += with_p ? 4 : 2;
		ctx->stat_ksw_chunks++;
	}
	GD_CUDA_OK(ctx, cudaGetLastError());
	return GD_OK;
}

int gd_ksw_compact_cigars(gd_ctx *ctx, int n, const gd_extz_t *d_ez, const uint32_t *d_cigar, int cigar_stride,
                          int64_t *d_off, uint32_t *d_compact, int64_t compact_cap, int64_t *d_run_base, int64_t *h_end)
{
	cudaStream_t s = ctx->stream;
	gd_ksw_cigar_scan_kernel<<<1, 1024, 0, s>>>(n, (const KswResult *)d_ez, d_off, d_run_base, h_end);
	ctx->stat_launches++;
	if (d_compact && d_cigar) {
		int blocks = std::min((n + 7) / 8, ctx->sms * 8);
		gd_ksw_cigar_gather_kernel<<<blocks, 256, 0, s>>>(n, (const KswResult *)d_ez, d_cigar, cigar_stride, d_off,
		                                                  d_compact, compact_cap);
		ctx->stat_launches++;
	}
	GD_CUDA_OK(ctx, cudaGetLastError());
	return GD_OK;
}

extern "C" int gd_exact_match_batch_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff,
                                           const uint8_t *d_qbuf, const int64_t *d_toff, const uint8_t *d_tbuf,
                                           uint8_t *d_equal)
{
	if (!ctx) return GD_ERR_ARG;
	if (n <= 0) return GD_OK;
	int blocks = std::min((n + 3) / 4, ctx->sms * 16);
	gd_exact_match_kernel<<<blocks, 128, 0, ctx->stream>>>(n, d_qlen, d_qoff, d_qbuf, d_toff, d_tbuf, d_equal);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	return GD_OK;
}
