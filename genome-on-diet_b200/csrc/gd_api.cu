// gd_api.cu -- C ABI of libgdiet_cuda.so (include/gdiet_cuda.h): context management, the host-buffer
// batched DP entry point and the drop-in ksw_extd2_sse / ksw_extd2_avx512 symbols.
#include "gd_ctx.h"
#include <algorithm>
#include <map>
#include <mutex>
#include <thread>
#include <vector>
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
	if (const char *e = getenv("GDIET_MAP_LANES")) // lanes of the mapping stage (contexts + host threads that drive alternate slices)
		if (atol(e) >= 1 && atol(e) <= 8) ctx->opt_map_lanes = atol(e);
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
	if (ctx->peer) gd_destroy(ctx->peer), ctx->peer = nullptr;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	GdBuf *bufs[] = {&ctx->tpk, &ctx->qpk, &ctx->parena, &ctx->ticket, &ctx->cig_tmp, &ctx->cig_off, &ctx->cig_compact,
	                 &ctx->res, &ctx->d_qlen, &ctx->d_tlen, &ctx->d_w, &ctx->d_qoff, &ctx->d_toff, &ctx->d_qbuf,
	                 &ctx->d_tbuf, &ctx->sk_seq, &ctx->sk_off, &ctx->sk_len, &ctx->sk_rid, &ctx->sk_out,
	                 &ctx->sk_out_off, &ctx->sk_state, &ctx->sk_misc, &ctx->sk_jobs, &ctx->sk_out2,
	                 &ctx->mp_seq, &ctx->mp_off, &ctx->mp_len, &ctx->mp_seed_n, &ctx->mp_seed_first, &ctx->mp_state, &ctx->mp_hoff,
	                 &ctx->mp_ht, &ctx->mp_hq, &ctx->mp_cand_tmp, &ctx->mp_ncand, &ctx->mp_coff, &ctx->mp_cand, &ctx->mp_qbuf,
	                 &ctx->mp_tbuf, &ctx->mp_pair, &ctx->mp_ez, &ctx->mp_cig, &ctx->mp_cnt, &ctx->mp_cpool, &ctx->mp_tmp,
	                 &ctx->lead64_scr, &ctx->lead64_list, &ctx->mp_names, &ctx->mp_qual, &ctx->mp_rnames, &ctx->mp_rcoff, &ctx->mp_cpool2,
	                 &ctx->mp_slen, &ctx->mp_soff, &ctx->mp_text};
	for (GdBuf *b : bufs)
		if (b->p) cudaFree(b->p);
	GdPinned *pins[] = {&ctx->h_stage, &ctx->h_res, &ctx->h_cig, &ctx->h_misc, &ctx->h_sk_stage, &ctx->h_sk_out, &ctx->h_sk_misc, &ctx->h_mp,
	                    &ctx->h_names, &ctx->h_sam[0], &ctx->h_sam[1]};
	for (GdPinned *b : pins)
		if (b->p) cudaFreeHost(b->p);
	for (int i = 0; i < 4; ++i)
		if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
	gd_timer_collect(ctx, ctx->tm_dp, ctx->tm_dp_us, ctx->tm_dp_n);
	gd_timer_collect(ctx, ctx->tm_sketch, ctx->tm_sketch_us, ctx->tm_sketch_n);
	for (cudaEvent_t e : ctx->tm_pool) cudaEventDestroy(e);
	cudaStreamDestroy(ctx->stream);
	cudaStreamDestroy(ctx->copy_stream);
	if (ctx->d2h_stream) cudaStreamDestroy(ctx->d2h_stream);
	delete ctx;
}

// page-locked host memory for a host's read / reference staging buffers (uploads from it run at PCIe speed and overlap)
extern "C" void *gd_pinned_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
		cudaGetLastError();
		return nullptr;
	}
	return p;
}
extern "C" void gd_pinned_free(void *p)
{
	if (p) cudaFreeHost(p);
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
	else if (!strcmp(key, "ksw_slice")) ctx->opt_ksw_slice = value;
	else if (!strcmp(key, "map_lanes")) ctx->opt_map_lanes = value;
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
	if (!strcmp(key, "kernel_launches")) {
		long t = 0;
		for (const gd_ctx *c = ctx; c; c = c->peer) t += c->stat_launches; // every lane of the mapping stage
		return t;
	}
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

// Host-buffer entry point.  The batch is cut into slices that flow through three streams: inputs go up on
// copy_stream, pack + DP + traceback + CIGAR compaction run on stream, results come down on d2h_stream, so
// the transfers of one slice overlap the kernels of its neighbours.  The dense CIGAR pool is filled in pair
// order by chaining the per-slice offset scans through a device-side running base.
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
	if (!ctx->d2h_stream) GD_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->d2h_stream, cudaStreamNonBlocking));
	// slices
	int slice = (int)ctx->opt_ksw_slice;
	if (slice <= 0) slice = n < 65536 ? n : std::max(32768, (n + 7) / 8);
	const int nsl = (n + slice - 1) / slice;
	// extents and bounds (whole batch and per slice)
	struct Sl {
		int b, n, max_q, max_t, max_w;
		int64_t q0, q1, t0, t1;
	};
	std::vector<Sl> sl(nsl);
	int64_t qbytes = 0, tbytes = 0;
	int max_q = 1, max_t = 1;
	auto scan_slice = [&](int k) {
		Sl &S = sl[k];
		S.b = k * slice, S.n = std::min(slice, n - S.b), S.max_q = S.max_t = 1, S.max_w = 0;
		S.q0 = S.t0 = INT64_MAX, S.q1 = S.t1 = 0;
		for (int i = S.b; i < S.b + S.n; ++i) {
			const int ql = std::max(qlen[i], 0), tl = std::max(tlen[i], 0);
			S.q0 = std::min<int64_t>(S.q0, qoff[i]), S.q1 = std::max<int64_t>(S.q1, qoff[i] + ql);
			S.t0 = std::min<int64_t>(S.t0, toff[i]), S.t1 = std::max<int64_t>(S.t1, toff[i] + tl);
			S.max_q = std::max(S.max_q, ql), S.max_t = std::max(S.max_t, tl);
			int ww = w ? w[i] : w_all;
			if (ww < 0) ww = std::max(ql, tl);
			S.max_w = std::max(S.max_w, ww);
		}
	};
	if (nsl > 1) { // one host thread per slice: the scan of a million descriptors is otherwise ~3 ms of the call
		std::vector<std::thread> th;
		for (int k = 0; k < nsl; ++k) th.emplace_back(scan_slice, k);
		for (auto &t : th) t.join();
	} else scan_slice(0);
	for (int k = 0; k < nsl; ++k) {
		const Sl &S = sl[k];
		if (S.q0 < 0 || S.t0 < 0) {
			ctx->err = "gd_ksw_extd2_batch: negative sequence offset";
			return GD_ERR_ARG;
		}
		qbytes = std::max(qbytes, S.q1), tbytes = std::max(tbytes, S.t1);
		max_q = std::max(max_q, S.max_q), max_t = std::max(max_t, S.max_t);
	}
	const bool want_cigar = !(prm->flag & 0x01) && cigar_off != nullptr;
	const int stride = want_cigar ? max_q + max_t : 0;
	const bool want_pool = want_cigar && cigar && cigar_cap > 0;
	int rc;
	if ((rc = gd_reserve(ctx, ctx->d_qlen, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_tlen, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_qoff, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_toff, (size_t)n * 8))) return rc;
	if (w && (rc = gd_reserve(ctx, ctx->d_w, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_qbuf, (size_t)qbytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->d_tbuf, (size_t)tbytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->res, (size_t)n * sizeof(gd_extz_t)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->cig_off, (size_t)(n + 2) * 8))) return rc;
	if (want_cigar && (rc = gd_reserve(ctx, ctx->cig_tmp, (size_t)std::min(slice, n) * stride * 4 * (nsl > 1 ? 2 : 1)))) return rc;
	if (want_pool && (rc = gd_reserve(ctx, ctx->cig_compact, (size_t)cigar_cap * 4))) return rc;
	if ((rc = gd_reserve_pinned(ctx, ctx->h_misc, (size_t)(nsl + 1) * 8))) return rc;
	int64_t *h_end = (int64_t *)ctx->h_misc.p;       // pool position after each slice (written by the device)
	int64_t *d_run = (int64_t *)ctx->cig_off.p + n + 1; // device-side running base
	cudaStream_t s = ctx->stream, up = ctx->copy_stream, down = ctx->d2h_stream;
	std::vector<cudaEvent_t> ev_up(nsl), ev_done(nsl);
	for (int k = 0; k < nsl; ++k) ev_up[k] = GdKernelTimer::get(ctx), ev_done[k] = GdKernelTimer::get(ctx);
	auto release = [&]() {
		for (int k = 0; k < nsl; ++k) ctx->tm_pool.push_back(ev_up[k]), ctx->tm_pool.push_back(ev_done[k]);
	};
	int status = GD_OK;
	// Every failure leaves through ONE exit that drains the three streams before the caller may free or reuse qbuf, ez,
	// cigar_off and cigar (copies that touch them can still be in flight) and gives the slice events back to the pool.
	auto work = [&]() -> int {
		GD_CUDA_OK(ctx, cudaMemsetAsync(d_run, 0, 8, s));
		// the previous call's kernels may still read the staging buffers: uploads start after them
		GD_CUDA_OK(ctx, cudaEventRecord(ev_done[0], s));
		GD_CUDA_OK(ctx, cudaStreamWaitEvent(up, ev_done[0], 0));
		for (int k = 0; k < nsl; ++k) { // ---- enqueue everything
			const Sl &S = sl[k];
			GD_CUDA_OK(ctx, cudaMemcpyAsync((int32_t *)ctx->d_qlen.p + S.b, qlen + S.b, (size_t)S.n * 4, cudaMemcpyHostToDevice, up));
			GD_CUDA_OK(ctx, cudaMemcpyAsync((int32_t *)ctx->d_tlen.p + S.b, tlen + S.b, (size_t)S.n * 4, cudaMemcpyHostToDevice, up));
			GD_CUDA_OK(ctx, cudaMemcpyAsync((int64_t *)ctx->d_qoff.p + S.b, qoff + S.b, (size_t)S.n * 8, cudaMemcpyHostToDevice, up));
			GD_CUDA_OK(ctx, cudaMemcpyAsync((int64_t *)ctx->d_toff.p + S.b, toff + S.b, (size_t)S.n * 8, cudaMemcpyHostToDevice, up));
			if (w) GD_CUDA_OK(ctx, cudaMemcpyAsync((int32_t *)ctx->d_w.p + S.b, w + S.b, (size_t)S.n * 4, cudaMemcpyHostToDevice, up));
			if (S.q1 > S.q0)
				GD_CUDA_OK(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_qbuf.p + S.q0, qbuf + S.q0, (size_t)(S.q1 - S.q0), cudaMemcpyHostToDevice, up));
			if (S.t1 > S.t0)
				GD_CUDA_OK(ctx, cudaMemcpyAsync((uint8_t *)ctx->d_tbuf.p + S.t0, tbuf + S.t0, (size_t)(S.t1 - S.t0), cudaMemcpyHostToDevice, up));
			GD_CUDA_OK(ctx, cudaEventRecord(ev_up[k], up));
			GD_CUDA_OK(ctx, cudaStreamWaitEvent(s, ev_up[k], 0));
			uint32_t *tmp = want_cigar ? (uint32_t *)ctx->cig_tmp.p + (size_t)(k & 1) * slice * stride : nullptr;
			gd_extz_t *d_res = (gd_extz_t *)ctx->res.p + S.b;
			rc = gd_ksw_run_device(ctx, S.n, (const int32_t *)ctx->d_qlen.p + S.b, (const int64_t *)ctx->d_qoff.p + S.b,
			                       (const uint8_t *)ctx->d_qbuf.p, (const int32_t *)ctx->d_tlen.p + S.b,
			                       (const int64_t *)ctx->d_toff.p + S.b, (const uint8_t *)ctx->d_tbuf.p,
			                       w ? (const int32_t *)ctx->d_w.p + S.b : nullptr, w_all, S.max_q, S.max_t, S.max_w, prm, d_res, tmp,
			                       stride);
			if (rc) return rc;
			if (cigar_off) {
				rc = gd_ksw_compact_cigars(ctx, S.n, d_res, want_pool ? tmp : nullptr, stride, (int64_t *)ctx->cig_off.p + S.b,
				                           want_pool ? (uint32_t *)ctx->cig_compact.p : nullptr, cigar_cap, d_run, h_end + k + 1);
				if (rc) return rc;
			}
			GD_CUDA_OK(ctx, cudaEventRecord(ev_done[k], s));
		}
		h_end[0] = 0;
		for (int k = 0; k < nsl; ++k) { // ---- results come down slice by slice while later slices compute
			const Sl &S = sl[k];
			GD_CUDA_OK(ctx, cudaStreamWaitEvent(down, ev_done[k], 0));
			GD_CUDA_OK(ctx, cudaMemcpyAsync(ez + S.b, (gd_extz_t *)ctx->res.p + S.b, (size_t)S.n * sizeof(gd_extz_t), cudaMemcpyDeviceToHost, down));
			if (cigar_off) {
				GD_CUDA_OK(ctx, cudaMemcpyAsync(cigar_off + S.b, (int64_t *)ctx->cig_off.p + S.b, (size_t)(S.n + 1) * 8, cudaMemcpyDeviceToHost, down));
				GD_CUDA_OK(ctx, cudaEventSynchronize(ev_done[k])); // h_end[k+1] is valid now
				if (h_end[k + 1] < 0) status = GD_ERR_ARG, h_end[k + 1] = -1 - h_end[k + 1]; // stride overflow inside the slice
				const int64_t lo = h_end[k], hi = h_end[k + 1];
				if (want_pool && hi > lo) {
					if (hi > cigar_cap) status = status == GD_OK ? GD_ERR_CAPACITY : status;
					else GD_CUDA_OK(ctx, cudaMemcpyAsync(cigar + lo, (uint32_t *)ctx->cig_compact.p + lo, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, down));
				}
			}
		}
		return GD_OK;
	};
	rc = work();
	if (rc) {
		cudaStreamSynchronize(up), cudaStreamSynchronize(s), cudaStreamSynchronize(down);
		release();
		return rc;
	}
	GD_CUDA_OK(ctx, cudaStreamSynchronize(down));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	release();
	if (status == GD_ERR_CAPACITY) {
		ctx->err = "gd_ksw_extd2_batch: cigar buffer too small";
		return status;
	}
	if (cigar && cigar_off && !want_pool && cigar_off[n] > 0) {
		ctx->err = "gd_ksw_extd2_batch: cigar buffer too small";
		return GD_ERR_CAPACITY;
	}
	if (status == GD_ERR_ARG || !cigar_off) { // (without CIGAR offsets nobody scanned n_cigar on the device)
		for (int i = 0; i < n; ++i)
			if (ez[i].n_cigar < 0) {
				ctx->err = "gd_ksw_extd2_batch: internal cigar stride overflow";
				return GD_ERR_CAPACITY;
			}
	}
	return GD_OK;
}

// ---- drop-in single-call entry points (GDiet-ShortReads/ksw2.h:68-69, ksw2_extd2_avx.h:38) ----
// The reference's kt_for / kt_pipeline start fresh pthreads for every mini-batch and join them (kthread.c:54-69,
// 139-160), so a context bound to the thread for good would be leaked once per worker and batch -- with its
// streams, events and grow-only device / pinned buffers.  Contexts therefore live in a per-device pool: a thread
// checks one out on its first drop-in call and its thread_local holder hands it back (buffers and all) when the
// thread exits; the next batch's workers reuse them.
namespace {
std::mutex g_pool_mu;
std::map<int, std::vector<gd_ctx *>> g_pool;
struct ThreadCtx {
	gd_ctx *ctx = nullptr;
	~ThreadCtx()
	{
		if (!ctx) return;
		cudaSetDevice(ctx->device);
		cudaStreamSynchronize(ctx->stream);
		std::lock_guard<std::mutex> lk(g_pool_mu);
		g_pool[ctx->device].push_back(ctx);
	}
};
} // namespace

gd_ctx *gd_thread_ctx()
{
	static thread_local ThreadCtx t;
	if (!t.ctx) {
		int dev = 0;
		const char *env = getenv("GDIET_DEVICE");
		if (env) dev = atoi(env);
		{
			std::lock_guard<std::mutex> lk(g_pool_mu);
			std::vector<gd_ctx *> &v = g_pool[dev];
			if (!v.empty()) t.ctx = v.back(), v.pop_back();
		}
		if (t.ctx) cudaSetDevice(dev);
		else if (gd_init(dev, &t.ctx) != GD_OK) {
			fprintf(stderr, "[gdiet_cuda] FATAL: %s (there is no CPU fallback)\n", gd_strerror(nullptr));
			abort();
		}
	}
	return t.ctx;
}

extern "C" long gd_thread_ctx_pool_size(int device)
{ // contexts parked in the pool of `device` (tests: the pool, not the number of threads ever started, bounds the footprint)
	std::lock_guard<std::mutex> lk(g_pool_mu);
	return (long)g_pool[device].size();
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
