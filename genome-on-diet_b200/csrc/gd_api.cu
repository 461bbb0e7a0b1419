// gd_api.cu -- C ABI of libgdiet_cuda.so (include/gdiet_cuda.h): context management, the host-buffer
// batched DP entry point and the drop-in ksw_extd2_sse / ksw_extd2_avx512 symbols.
#include "gd_ctx.h"
#include <algorithm>
#include <mutex>
#include <stdlib.h>
#include <string.h>

static std::string g_init_err;

// kalloc of the host program (GDiet-ShortReads/kalloc.h:11-15) when we are linked into it.
extern "C" {
void *krealloc(void *km, void *ptr, size_t size) __attribute__((weak));
void *kmalloc(void *km, size_t size) __attribute__((weak));
}
void *gd_host_realloc(void *km, void *ptr, size_t size)
{
	if (krealloc) return krealloc(km, ptr, size);
	if (km) {
		fprintf(stderr, "[gdiet_cuda] a kalloc pool was passed but the host program exports no krealloc()\n");
		abort();
	}
	return realloc(ptr, size);
}
void *gd_host_malloc(void *km, size_t size)
{
	if (kmalloc) return kmalloc(km, size);
	if (km) {
		fprintf(stderr, "[gdiet_cuda] a kalloc pool was passed but the host program exports no kmalloc()\n");
		abort();
	}
	return malloc(size);
}

extern "C" int gd_init(int device, gd_ctx **out)
{
	if (!out) return GD_ERR_ARG;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev <= 0) {
		g_init_err = std::string("no CUDA device: ") + cudaGetErrorString(e);
		return GD_ERR_NO_DEVICE;
	}
	if (device < 0 || device >= ndev) {
		g_init_err = "device index out of range";
		return GD_ERR_NO_DEVICE;
	}
	cudaDeviceProp prop;
	if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
		g_init_err = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
		return GD_ERR_NO_DEVICE;
	}
	if (prop.major != 10) {
		g_init_err = "libgdiet_cuda.so is built for sm_100a only (found sm_" + std::to_string(prop.major) +
		             std::to_string(prop.minor) + ")";
		return GD_ERR_NO_DEVICE;
	}
	gd_ctx *ctx = new gd_ctx();
	ctx->device = device;
	ctx->sms = prop.multiProcessorCount;
	ctx->smem_optin = prop.sharedMemPerBlockOptin;
	ctx->smem_per_sm = prop.sharedMemPerMultiprocessor;
	if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
	    cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) {
		g_init_err = "cudaStreamCreate failed";
		delete ctx;
		return GD_ERR_CUDA;
	}
	for (int i = 0; i < 4; ++i) cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming);
	*out = ctx;
	return GD_OK;
}

extern "C" void gd_destroy(gd_ctx *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	GdBuf *bufs[] = {&ctx->tpk, &ctx->qpk, &ctx->parena, &ctx->ticket, &ctx->cig_tmp, &ctx->cig_off, &ctx->cig_compact,
	                 &ctx->res, &ctx->d_qlen, &ctx->d_tlen, &ctx->d_w, &ctx->d_qoff, &ctx->d_toff, &ctx->d_qbuf,
	                 &ctx->d_tbuf, &ctx->sk_seq, &ctx->sk_off, &ctx->sk_len, &ctx->sk_rid, &ctx->sk_out,
	                 &ctx->sk_out_off, &ctx->sk_state, &ctx->sk_misc, &ctx->sk_jobs, &ctx->sk_out2};
	for (GdBuf *b : bufs)
		if (b->p) cudaFree(b->p);
	GdPinned *pins[] = {&ctx->h_stage, &ctx->h_res, &ctx->h_cig, &ctx->h_misc, &ctx->h_sk_stage, &ctx->h_sk_out, &ctx->h_sk_misc};
	for (GdPinned *b : pins)
		if (b->p) cudaFreeHost(b->p);
	for (int i = 0; i < 4; ++i)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	gd_timer_collect(ctx, ctx->tm_dp, ctx->tm_dp_us, ctx->tm_dp_n);
	gd_timer_collect(ctx, ctx->tm_sketch, ctx->tm_sketch_us, ctx->tm_sketch_n);
	for (cudaEvent_t e : ctx->tm_pool) cudaEventDestroy(e);
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->copy_stream);
	delete ctx;
}

extern "C" const char *gd_strerror(const gd_ctx *ctx) { return ctx ? ctx->err.c_str() : g_init_err.c_str(); }

extern "C" int gd_set_option(gd_ctx *ctx, const char *key, long value)
{
	if (!ctx || !key) return GD_ERR_ARG;
	if (!strcmp(key, "ksw_group")) ctx->opt_ksw_group = value;
	else if (!strcmp(key, "p_budget_mb")) ctx->opt_p_budget_mb = value;
	else if (!strcmp(key, "ksw_blocks_per_sm")) ctx->opt_ksw_blocks_per_sm = value;
	else if (!strcmp(key, "sketch_chunk")) ctx->opt_sketch_chunk = value;
	else if (!strcmp(key, "time_kernels")) ctx->opt_time_kernels = value;
	else {
		ctx->err = std::string("unknown option ") + key;
		return GD_ERR_ARG;
	}
	return GD_OK;
}

extern "C" long gd_get_stat(const gd_ctx *cctx, const char *key)
{
	if (!cctx || !key) return -1;
	gd_ctx *ctx = const_cast<gd_ctx *>(cctx);
	if (!strncmp(key, "ksw_dp_", 7) || !strncmp(key, "sketch_", 7)) { // device time of the hot kernels (option "time_kernels")
		gd_timer_collect(ctx, ctx->tm_dp, ctx->tm_dp_us, ctx->tm_dp_n);
		gd_timer_collect(ctx, ctx->tm_sketch, ctx->tm_sketch_us, ctx->tm_sketch_n);
		if (!strcmp(key, "ksw_dp_us")) return (long)(ctx->tm_dp_us + 0.5);
		if (!strcmp(key, "ksw_dp_launches")) return ctx->tm_dp_n;
		if (!strcmp(key, "sketch_us")) return (long)(ctx->tm_sketch_us + 0.5);
		if (!strcmp(key, "sketch_launches")) return ctx->tm_sketch_n;
		if (!strcmp(key, "ksw_dp_reset") || !strcmp(key, "sketch_reset")) {
			ctx->tm_dp_us = ctx->tm_sketch_us = 0, ctx->tm_dp_n = ctx->tm_sketch_n = 0;
			return 0;
		}
	}
	if (!strcmp(key, "kernel_launches")) return ctx->stat_launches;
	if (!strcmp(key, "ksw_ring")) return ctx->stat_ksw_ring;
	if (!strcmp(key, "ksw_group")) return ctx->stat_ksw_group;
	if (!strcmp(key, "ksw_chunks")) return ctx->stat_ksw_chunks;
	if (!strcmp(key, "device_sms")) return ctx->sms;
	return -1;
}

extern "C" void *gd_stream(gd_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" int gd_ksw_extd2_batch_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff,
                                         const uint8_t *d_qbuf, const int32_t *d_tlen, const int64_t *d_toff,
                                         const uint8_t *d_tbuf, const int32_t *d_w, int w_all, int max_qlen,
                                         int max_tlen, int max_w, const gd_ksw_params_t *prm, gd_extz_t *d_ez,
                                         uint32_t *d_cigar, int cigar_stride)
{
	if (!ctx) return GD_ERR_ARG;
	cudaSetDevice(ctx->device);
	return gd_ksw_run_device(ctx, n, d_qlen, d_qoff, d_qbuf, d_tlen, d_toff, d_tbuf, d_w, w_all, max_qlen, max_tlen,
	                         max_w, prm, d_ez, d_cigar, cigar_stride);
}

extern "C" int gd_ksw_extd2_batch(gd_ctx *ctx, int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf,
                                  const int32_t *tlen, const int64_t *toff, const uint8_t *tbuf, const int32_t *w,
                                  int w_all, const gd_ksw_params_t *prm, gd_extz_t *ez, int64_t *cigar_off,
                                  uint32_t *cigar, int64_t cigar_cap)
{
	if (!ctx) return GD_ERR_ARG;
	if (n < 0 || !prm || !ez || (n > 0 && (!qlen || !qoff || !qbuf || !tlen || !toff || !tbuf))) {
		ctx->err = "gd_ksw_extd2_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (cigar_off) cigar_off[0] = 0;
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	// extents and bounds
	int64_t qbytes = 0, tbytes = 0;
	int max_q = 1, max_t = 1, max_w = 0;
	for (int i = 0; i < n; ++i) {
		const int ql = std::max(qlen[i], 0), tl = std::max(tlen[i], 0);
		qbytes = std::max<int64_t>(qbytes, qoff[i] + ql);
		tbytes = std::max<int64_t>(tbytes, toff[i] + tl);
		max_q = std::max(max_q, ql), max_t = std::max(max_t, tl);
		int ww = w ? w[i] : w_all;
		if (ww < 0) ww = std::max(ql, tl);
		max_w = std::max(max_w, ww);
	}
	const bool want_cigar = !(prm->flag & 0x01) && cigar_off != nullptr;
	const int stride = want_cigar ? max_q + max_t : 0;
	int rc;
	if ((rc = gd_reserve(ctx, ctx->d_qlen, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_tlen, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_qoff, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_toff, (size_t)n * 8))) return rc;
	if (w && (rc = gd_reserve(ctx, ctx->d_w, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_qbuf, (size_t)qbytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_tbuf, (size_t)tbytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->res, (size_t)n * sizeof(gd_extz_t)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->cig_off, (size_t)(n + 1) * 8))) return rc;
	if (want_cigar && (rc = gd_reserve(ctx, ctx->cig_tmp, (size_t)n * stride * 4))) return rc;
	cudaStream_t s = ctx->stream;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_qlen.p, qlen, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tlen.p, tlen, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_qoff.p, qoff, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_toff.p, toff, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	if (w) GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_w.p, w, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_qbuf.p, qbuf, (size_t)qbytes, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->d_tbuf.p, tbuf, (size_t)tbytes, cudaMemcpyHostToDevice, s));
	rc = gd_ksw_run_device(ctx, n, (const int32_t *)ctx->d_qlen.p, (const int64_t *)ctx->d_qoff.p,
	                       (const uint8_t *)ctx->d_qbuf.p, (const int32_t *)ctx->d_tlen.p, (const int64_t *)ctx->d_toff.p,
	                       (const uint8_t *)ctx->d_tbuf.p, w ? (const int32_t *)ctx->d_w.p : nullptr, w_all, max_q, max_t,
	                       max_w, prm, (gd_extz_t *)ctx->res.p, want_cigar ? (uint32_t *)ctx->cig_tmp.p : nullptr, stride);
	if (rc) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ez, ctx->res.p, (size_t)n * sizeof(gd_extz_t), cudaMemcpyDeviceToHost, s));
	if (cigar_off) {
		// offsets first (tiny), then a compact gather sized from the total
		rc = gd_ksw_compact_cigars(ctx, n, (const gd_extz_t *)ctx->res.p, nullptr, stride, (int64_t *)ctx->cig_off.p,
		                           nullptr, 0);
		if (rc) return rc;
		GD_CUDA_OK(ctx, cudaMemcpyAsync(cigar_off, ctx->cig_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
		GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
		const int64_t total = cigar_off[n];
		if (cigar && total > 0) {
			if (total > cigar_cap) {
				ctx->err = "gd_ksw_extd2_batch: cigar buffer too small";
				return GD_ERR_CAPACITY;
			}
			if ((rc = gd_reserve(ctx, ctx->cig_compact, (size_t)total * 4))) return rc;
			rc = gd_ksw_compact_cigars(ctx, n, (const gd_extz_t *)ctx->res.p, (const uint32_t *)ctx->cig_tmp.p, stride,
			                           (int64_t *)ctx->cig_off.p, (uint32_t *)ctx->cig_compact.p, total);
			if (rc) return rc;
			GD_CUDA_OK(ctx, cudaMemcpyAsync(cigar, ctx->cig_compact.p, (size_t)total * 4, cudaMemcpyDeviceToHost, s));
		}
	}
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	for (int i = 0; i < n; ++i)
		if (ez[i].n_cigar < 0) {
			ctx->err = "gd_ksw_extd2_batch: internal cigar stride overflow";
			return GD_ERR_CAPACITY;
		}
	return GD_OK;
}

// ---- drop-in single-call entry points (GDiet-ShortReads/ksw2.h:68-69, ksw2_extd2_avx.h:38) ----
gd_ctx *gd_thread_ctx()
{
	static thread_local gd_ctx *tctx = nullptr;
	if (!tctx) {
		int dev = 0;
		const char *env = getenv("GDIET_DEVICE");
		if (env) dev = atoi(env);
		if (gd_init(dev, &tctx) != GD_OK) {
			fprintf(stderr, "[gdiet_cuda] FATAL: %s (there is no CPU fallback)\n", gd_strerror(nullptr));
			abort();
		}
	}
	return tctx;
}

static void extd2_one(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                      const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus,
                      int flag, ksw_extz_t *ez)
{
	// ksw_reset_extz (ksw2.h:165-170): m_cigar / cigar are kept, everything else reset
	ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
	ez->max = 0, ez->score = ez->mqe = ez->mte = -0x40000000;
	ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
	if (m <= 1 || qlen <= 0 || tlen <= 0) return;
	gd_ctx *ctx = gd_thread_ctx();
	gd_ksw_params_t prm = {m, mat, q, e, q2, e2, zdrop, end_bonus, flag};
	const int64_t zero = 0;
	gd_extz_t r;
	int64_t coff[2] = {0, 0};
	const int cap = qlen + tlen + 2;
	uint32_t *tmp = (uint32_t *)malloc((size_t)cap * 4);
	int rc = gd_ksw_extd2_batch(ctx, 1, &qlen, &zero, query, &tlen, &zero, target, nullptr, w, &prm, &r, coff, tmp, cap);
	if (rc != GD_OK) {
		fprintf(stderr, "[gdiet_cuda] FATAL: ksw_extd2: %s\n", gd_strerror(ctx));
		abort();
	}
	ez->max = (uint32_t)r.max, ez->zdropped = (uint32_t)r.zdropped;
	ez->max_q = r.max_q, ez->max_t = r.max_t, ez->mqe = r.mqe, ez->mqe_t = r.mqe_t, ez->mte = r.mte, ez->mte_q = r.mte_q;
	ez->score = r.score, ez->reach_end = r.reach_end;
	if (r.n_cigar > 0) {
		if (r.n_cigar > ez->m_cigar) { // grow like ksw_push_cigar (ksw2.h:100-111): doubling from 4
			int mc = ez->m_cigar ? ez->m_cigar : 4;
			while (mc < r.n_cigar) mc <<= 1;
			ez->cigar = (uint32_t *)gd_host_realloc(km, ez->cigar, (size_t)mc << 2);
			ez->m_cigar = mc;
		}
		memcpy(ez->cigar, tmp, (size_t)r.n_cigar * 4);
	}
	ez->n_cigar = r.n_cigar;
	free(tmp);
}

extern "C" void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                              const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                              int end_bonus, int flag, ksw_extz_t *ez)
{
	extd2_one(km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, ez);
}

extern "C" void ksw_extd2_avx512(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                                 const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop,
                                 int end_bonus, int flag, ksw_extz_t *ez)
{
	extd2_one(km, qlen, query, tlen, target, m, mat, q, e, q2, e2, w, zdrop, end_bonus, flag, ez);
}
