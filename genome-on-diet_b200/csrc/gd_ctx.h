// gd_ctx.h -- the context object behind the C ABI (include/gdiet_cuda.h): one CUDA stream plus
// grow-only device / pinned-host staging buffers.  One context per host thread.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>
#include "../../include/gdiet_cuda.h"

struct GdBuf { // grow-only device buffer
	void *p = nullptr;
	size_t cap = 0;
};
struct GdPinned { // grow-only pinned host buffer
	void *p = nullptr;
	size_t cap = 0;
};

struct gd_ctx {
	int device = 0;
	int sms = 0;
	size_t smem_optin = 0;  // max dynamic shared memory per block
	size_t smem_per_sm = 0; // shared memory per SM
	cudaStream_t stream = nullptr;
	cudaStream_t copy_stream = nullptr; // host -> device
	cudaStream_t d2h_stream = nullptr;  // device -> host (PCIe is full duplex)
	cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
	std::string err;
	// options
	long opt_ksw_group = 0;
	long opt_p_budget_mb = 0; // 0 = auto
	long opt_ksw_blocks_per_sm = 0;
	long opt_sketch_chunk = 0;
	long opt_ksw_slice = 0;    // pairs per pipeline slice of the host-buffer DP call (0 = auto)
	long opt_time_kernels = 0; // 1: bracket every DP / sketch kernel launch with CUDA events (bench.py roofline)
	// mapping stage: the slices of a call are dealt to this many lanes = contexts, each driven by its own host thread, so that the
	// transfers, 8-byte read-backs and small kernels of one slice overlap the DP kernel of the others (10 M reads -> SAM text on one
	// B200: 14.5 / 16.5 / 16.9 / 17.3 M reads/s with 2 / 3 / 4 / 6 lanes; long reads use two at most: every lane owns a backtrack arena)
	long opt_map_lanes = 4;
	gd_ctx *peer = nullptr;    // the next lane (created on first use; a chain: peer, peer->peer, ...)
	// stats
	long stat_launches = 0;
	long stat_ksw_ring = 0, stat_ksw_group = 0, stat_ksw_chunks = 0;
	// per-kernel device time (opt_time_kernels): event pairs recorded on `stream`, summed when a stat is read
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> tm_dp, tm_sketch;
	std::vector<cudaEvent_t> tm_pool;
	double tm_dp_us = 0, tm_sketch_us = 0;
	long tm_dp_n = 0, tm_sketch_n = 0;
	// DP scratch
	GdBuf tpk, qpk, parena, ticket, cig_tmp, cig_off, cig_compact, res, lead64_scr, lead64_list;
	GdBuf d_qlen, d_tlen, d_w, d_qoff, d_toff, d_qbuf, d_tbuf;
	GdPinned h_stage, h_res, h_cig, h_misc;
	// sketch scratch
	GdBuf sk_seq, sk_off, sk_len, sk_rid, sk_out, sk_out_off, sk_state, sk_misc, sk_jobs, sk_out2;
	GdPinned h_sk_stage, h_sk_out, h_sk_misc;
	// short-read mapping scratch (gd_map.cu)
	GdBuf mp_seq, mp_off, mp_len, mp_seed_n, mp_seed_first, mp_state, mp_hoff, mp_ht, mp_hq, mp_cand_tmp, mp_ncand, mp_coff, mp_cand,
	    mp_qbuf, mp_tbuf, mp_pair, mp_ez, mp_cig, mp_cnt, mp_cpool, mp_tmp;
	GdPinned h_mp;
	// SAM text produced on the device (gd_sr_map_sam_batch): per-slice scratch, and two pinned text buffers that alternate
	// between calls (the pieces of a call stay valid until the call after the next one)
	GdBuf mp_names, mp_qual, mp_rnames, mp_rcoff, mp_cpool2, mp_slen, mp_soff, mp_text;
	GdPinned h_names, h_sam[2];
	size_t h_sam_used[2] = {0, 0};
	long sam_calls = 0;
};

#define GD_CUDA_OK(ctx, call)                                                                          \
	do {                                                                                           \
		cudaError_t e__ = (call);                                                              \
		if (e__ != cudaSuccess) {                                                              \
			char b__[512];                                                                 \
			snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			(ctx)->err = b__;                                                              \
			return GD_ERR_CUDA;                                                            \
		}                                                                                      \
	} while (0)

// Device-time bracket around one kernel launch: a no-op unless the "time_kernels" option is set.
struct GdKernelTimer {
	gd_ctx *ctx;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *list;
	cudaEvent_t e0 = nullptr, e1 = nullptr;
	static cudaEvent_t get(gd_ctx *c)
	{
		cudaEvent_t e = nullptr;
		if (!c->tm_pool.empty()) e = c->tm_pool.back(), c->tm_pool.pop_back();
		else cudaEventCreate(&e);
		return e;
	}
	GdKernelTimer(gd_ctx *c, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> *l) : ctx(c), list(l)
	{
		if (!ctx->opt_time_kernels) return;
		e0 = get(ctx), e1 = get(ctx);
		cudaEventRecord(e0, ctx->stream);
	}
	~GdKernelTimer()
	{
		if (!e0) return;
		cudaEventRecord(e1, ctx->stream);
		list->push_back(std::make_pair(e0, e1));
	}
};
static inline void gd_timer_collect(gd_ctx *ctx, std::vector<std::pair<cudaEvent_t, cudaEvent_t>> &l, double &us, long &n)
{
	for (auto &pr : l) {
		float ms = 0;
		cudaEventSynchronize(pr.second);
		if (cudaEventElapsedTime(&ms, pr.first, pr.second) == cudaSuccess) us += 1e3 * ms, ++n;
		ctx->tm_pool.push_back(pr.first), ctx->tm_pool.push_back(pr.second);
	}
	l.clear();
}

static inline int gd_reserve(gd_ctx *ctx, GdBuf &b, size_t bytes)
{
	if (bytes <= b.cap) return GD_OK;
	if (b.p) GD_CUDA_OK(ctx, cudaFree(b.p));
	b.p = nullptr, b.cap = 0;
	size_t want = bytes + std::min<size_t>(bytes / 8, (size_t)256 << 20) + 256; // (slack: at most 256 MB)
	GD_CUDA_OK(ctx, cudaMalloc(&b.p, want));
	b.cap = want;
	return GD_OK;
}
static inline int gd_reserve_pinned(gd_ctx *ctx, GdPinned &b, size_t bytes)
{
	if (bytes <= b.cap) return GD_OK;
	if (b.p) GD_CUDA_OK(ctx, cudaFreeHost(b.p));
	b.p = nullptr, b.cap = 0;
	size_t want = bytes + bytes / 8 + 256;
	GD_CUDA_OK(ctx, cudaHostAlloc(&b.p, want, cudaHostAllocDefault));
	b.cap = want;
	return GD_OK;
}

// implemented in gd_sketch.cu: device-resident read sketching for the mapper
struct GdReadSketch {
	const int64_t *job_off; // [n*JW+1]
	const uint64_t *raw;    // x,y pairs of every job
	const int64_t *c3, *c2; // [n*W] capped list lengths of mm_sketch3 / mm_sketch2
	const uint32_t *ret3;   // [n*W] return value of mm_sketch3
	int JW, crop;
	int64_t raw_cap;
};
int gd_sketch_reads_device_raw(gd_ctx *ctx, int n, const int64_t *d_off, const int32_t *d_len, const char *d_buf,
                               int64_t max_len, int64_t sum_len, int w, int k, const char *Z, int W, float max_seeds,
                               uint32_t max_nb_seeds, GdReadSketch *out);

// implemented in gd_ksw.cu
int gd_ksw_run_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff, const uint8_t *d_qbuf,
                      const int32_t *d_tlen, const int64_t *d_toff, const uint8_t *d_tbuf, const int32_t *d_w,
                      int w_all, int max_qlen, int max_tlen, int max_w, const gd_ksw_params_t *prm, gd_extz_t *d_ez,
                      uint32_t *d_cigar, int cigar_stride);
// offsets (+ optional gather into a dense pool); d_run_base: device word with the pool position to start from
// (updated to the end), h_end: pinned host word that receives the end -- both may be NULL
int gd_ksw_compact_cigars(gd_ctx *ctx, int n, const gd_extz_t *d_ez, const uint32_t *d_cigar, int cigar_stride,
                          int64_t *d_off /*n+1*/, uint32_t *d_compact, int64_t compact_cap, int64_t *d_run_base = nullptr,
                          int64_t *h_end = nullptr);
