// gd_sketch.cu -- kernels, launchers and C ABI of sparsified sketching (see gd_sketch.cuh).
#include "gd_ctx.h"
#include "gd_sketch.cuh"
#include <algorithm>
#include <string.h>
#include <vector>

using namespace gd;

void *gd_host_realloc(void *km, void *ptr, size_t size);
void *gd_host_malloc(void *km, size_t size);
gd_ctx *gd_thread_ctx();

// --------------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------------
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) gd_sketch_tile_kernel(const SketchParams S, SketchBatch B)
{
	extern __shared__ __align__(16) uint8_t gd_sk_smem[];
	if (B.tile_base) B.ntiles = B.tile_base[B.njobs];
	sketch_tile_body<THREADS>(S, B, (SketchSmem<THREADS> *)gd_sk_smem);
}
// design v3 of the tile body (gd_sketch.cuh); the default.  GDIET_SK_V=2 selects the kernel above (A/B measurements).
template <int THREADS, bool PACKED>
__global__ void __launch_bounds__(THREADS, 1024 / THREADS) gd_sketch_tile3_kernel(const SketchParams S, SketchBatch B)
{
	extern __shared__ __align__(16) uint8_t gd_sk_smem[];
	if (B.tile_base) B.ntiles = B.tile_base[B.njobs];
	sketch_tile_body3<THREADS, PACKED>(S, B, (SketchSmem3<THREADS> *)gd_sk_smem);
}

// jobs for index-build sketching: one per sequence, shift 0
__global__ void gd_sketch_ref_jobs_kernel(const SketchParams S, int n, const int64_t *off, const int32_t *len, const uint32_t *rid, SketchJob *jobs)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	SketchJob j;
	j.seq_off = off[i], j.len = len[i], j.shift = 0, j.rid = rid ? rid[i] : (uint32_t)i, j.dl = (int32_t)sk_diet_len((uint32_t)len[i], 0, S);
	jobs[i] = j;
}

// jobs for read sketching: per read W full-length jobs (shift 0..W-1) and, if crop, one extra job
// (shift 0 on the prefix (unsigned)(max_seeds*len), GDiet-ShortReads/sketch.c:2177-2183)
__global__ void gd_sketch_read_jobs_kernel(const SketchParams S, int n, const int64_t *off, const int32_t *len, int W, int crop, float max_seeds,
                                           SketchJob *jobs)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int JW = W + (crop ? 1 : 0);
	SketchJob j;
	j.seq_off = off[i], j.rid = 0;
	for (int s = 0; s < W; ++s) {
		j.len = len[i], j.shift = s, j.dl = (int32_t)sk_diet_len((uint32_t)len[i], (uint32_t)s, S);
		jobs[(size_t)i * JW + s] = j;
	}
	if (crop) {
		j.len = (int32_t)(unsigned)((float)max_seeds * len[i]), j.shift = 0, j.dl = (int32_t)sk_diet_len((uint32_t)j.len, 0, S);
		jobs[(size_t)i * JW + W] = j;
	}
}

// tiles per job -> written to tb[i]; scanned in place afterwards
__global__ void gd_sketch_tiles_per_job_kernel(const SketchParams S, int n, const SketchJob *jobs, int64_t *tb)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t dl = sk_diet_len((uint32_t)jobs[i].len, (uint32_t)jobs[i].shift, S);
	const int64_t t = (dl + S.TP - 1) / S.TP;
	tb[i] = t < 1 ? 1 : t;
}

// ---- generic exclusive scan of int64 (three phases, 4096 elements per block) ----
#define GD_SCAN_TILE 4096
__global__ void __launch_bounds__(1024) gd_scan_partial_kernel(int64_t n, const int64_t *in, int64_t *block_sum)
{
	__shared__ int64_t ws[32];
	const int64_t base = (int64_t)blockIdx.x * GD_SCAN_TILE;
	int64_t v = 0;
	for (int k = 0; k < 4; ++k) {
		int64_t i = base + threadIdx.x * 4 + k;
		if (i < n) v += in[i];
	}
	for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
	if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = v;
	__syncthreads();
	if (threadIdx.x < 32) {
		int64_t s = ws[threadIdx.x];
		for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
		if (threadIdx.x == 0) block_sum[blockIdx.x] = s;
	}
}
__global__ void __launch_bounds__(1024) gd_scan_blocksums_kernel(int64_t nb, int64_t *block_sum, int64_t *total)
{ // single block, in-place exclusive scan of block sums
	__shared__ int64_t ws[32];
	__shared__ int64_t carry;
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	if (threadIdx.x == 0) carry = 0;
	__syncthreads();
	for (int64_t base = 0; base < nb; base += 1024) {
		const int64_t i = base + threadIdx.x;
		const int64_t v = i < nb ? block_sum[i] : 0;
		int64_t inc = v;
		for (int d = 1; d < 32; d <<= 1) {
			int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
			if (lane >= d) inc += o;
		}
		if (lane == 31) ws[wid] = inc;
		__syncthreads();
		if (wid == 0) {
			int64_t s = ws[lane], si = s;
			for (int d = 1; d < 32; d <<= 1) {
				int64_t o = __shfl_up_sync(0xffffffffu, si, d);
				if (lane >= d) si += o;
			}
			ws[lane] = si - s;
		}
		__syncthreads();
		const int64_t excl = carry + ws[wid] + inc - v;
		if (i < nb) block_sum[i] = excl;
		__syncthreads();
		if (threadIdx.x == 1023) carry = excl + v;
		__syncthreads();
	}
	if (threadIdx.x == 0 && total) *total = carry;
}
__global__ void __launch_bounds__(1024) gd_scan_final_kernel(int64_t n, const int64_t *in, const int64_t *block_excl, int64_t *out)
{ // out may alias in
	__shared__ int64_t ws[32];
	const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
	const int64_t base = (int64_t)blockIdx.x * GD_SCAN_TILE + threadIdx.x * 4;
	int64_t v[4], s = 0;
	for (int k = 0; k < 4; ++k) {
		v[k] = base + k < n ? in[base + k] : 0;
		s += v[k];
	}
	int64_t inc = s;
	for (int d = 1; d < 32; d <<= 1) {
		int64_t o = __shfl_up_sync(0xffffffffu, inc, d);
		if (lane >= d) inc += o;
	}
	if (lane == 31) ws[wid] = inc;
	__syncthreads();
	if (wid == 0) {
		int64_t t = ws[lane], ti = t;
		for (int d = 1; d < 32; d <<= 1) {
			int64_t o = __shfl_up_sync(0xffffffffu, ti, d);
			if (lane >= d) ti += o;
		}
		ws[lane] = ti - t;
	}
	__syncthreads();
	int64_t run = block_excl[blockIdx.x] + ws[wid] + inc - s;
	for (int k = 0; k < 4; ++k) {
		if (base + k < n) out[base + k] = run;
		run += v[k];
	}
}

// exclusive scan: out[0..n) = prefix of in, out[n] = total (out has n+1 entries, may alias in)
static int gd_exclusive_scan(gd_ctx *ctx, int64_t n, const int64_t *in, int64_t *out, GdBuf &scratch)
{
	const int64_t nb = (n + GD_SCAN_TILE - 1) / GD_SCAN_TILE;
	int rc = gd_reserve(ctx, scratch, (size_t)(nb + 1) * 8);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	int64_t *bs = (int64_t *)scratch.p;
	if (nb > 0) gd_scan_partial_kernel<<<(unsigned)nb, 1024, 0, s>>>(n, in, bs);
	gd_scan_blocksums_kernel<<<1, 1024, 0, s>>>(nb, bs, out + n);
	if (nb > 0) gd_scan_final_kernel<<<(unsigned)nb, 1024, 0, s>>>(n, in, bs, out);
	ctx->stat_launches += 3;
	GD_CUDA_OK(ctx, cudaGetLastError());
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// launcher: run a job list through the tile kernel
// --------------------------------------------------------------------------------------------
static int make_params(gd_ctx *ctx, int w, int k, const char *Z, int W, SketchParams &S)
{
	if (w <= 0 || w >= 256 || k <= 0 || k > 28 || !Z || W <= 0 || W > 64) { // asserts of sketch.c:160
		ctx->err = "gd_sketch: need 0<w<256, 0<k<=28, 0<W<=64";
		return GD_ERR_ARG;
	}
	memset(&S, 0, sizeof(S));
	S.w = w, S.k = k, S.W = W, S.mask = (1ull << 2 * k) - 1;
	for (int g = 0; g < W; ++g)
		if (Z[g] == '1') S.ones_loc[S.ones++] = (uint8_t)g;
	if (S.ones == 0 || S.ones > 40) { // ones_loc[40] in the reference (sketch.c:170)
		ctx->err = "gd_sketch: pattern needs 1..40 ones";
		return GD_ERR_ARG;
	}
	return GD_OK;
}

// max_dl: upper bound of the sparsified length of any job; pos_total: upper bound of the sum.
// fixed_stride > 0 (short jobs only, one tile per job): job j's records go to d_out[j * fixed_stride ..], its count to
// d_out_cnt[j]; the caller sized d_out for njobs * fixed_stride records.
static int gd_sketch_run_jobs(gd_ctx *ctx, SketchParams &S, int njobs, const SketchJob *d_jobs, int64_t max_dl,
                              int64_t pos_total, const char *d_buf, int64_t *d_out_off, uint64_t *d_out, int64_t out_cap,
                              int64_t fixed_stride = 0, int32_t *d_out_cnt = nullptr, int pack_group = 0, int64_t pack_slots = 0)
{ // pack_group / pack_slots (fixed-stride mode): the jobs come in groups of pack_group (a read: one job per shift and the cropped
  // job) that need at most pack_slots tile slots (sparsified positions of the jobs that can emit, one N behind each)
	cudaStream_t s = ctx->stream;
	int rc;
	SketchBatch B;
	memset(&B, 0, sizeof(B));
	B.njobs = njobs, B.jobs = d_jobs, B.buf = d_buf, B.out_off = d_out_off, B.out = d_out, B.out_cap = out_cap;
	// short reads: one warp-sized tile (256 positions) per job; otherwise 2048-position tiles
	const int tp_small = sk_tile_emit(256, S.w, S.k);
	const bool small = tp_small > 0 && max_dl <= tp_small;
	// threads per block of the large tiles (8 positions per thread): smaller blocks wait less at the five block-wide barriers
	// of a tile, larger ones spend less of a tile on its halo of 3w+k-4 positions (developer switch GDIET_SK_THREADS = 64 | 128 | 256)
	static const int big_env = getenv("GDIET_SK_THREADS") ? atoi(getenv("GDIET_SK_THREADS")) : 0;
	int big_threads = (big_env == 64 || big_env == 128 || big_env == 256) ? big_env : 256;
	while (big_threads < 256 && sk_tile_emit(big_threads * 8, S.w, S.k) < big_threads * 4) big_threads *= 2; // wide windows need wide tiles
	int64_t ntiles_bound;
	if (fixed_stride > 0 && !small) {
		ctx->err = "sketch: fixed-stride output needs one tile per job";
		return GD_ERR_ARG;
	}
	if (small) {
		S.TP = tp_small, S.one_tile_per_job = 1;
		ntiles_bound = njobs;
		B.ntiles = njobs, B.tile_base = nullptr;
		B.fixed_stride = fixed_stride, B.out_cnt = d_out_cnt;
	} else {
		S.TP = sk_tile_emit(big_threads * 8, S.w, S.k), S.one_tile_per_job = 0;
		ntiles_bound = pos_total / S.TP + 2 * (int64_t)njobs + 2;
		if ((rc = gd_reserve(ctx, ctx->sk_misc, (size_t)(njobs + 1) * 8))) return rc;
		int64_t *tb = (int64_t *)ctx->sk_misc.p;
		gd_sketch_tiles_per_job_kernel<<<(njobs + 255) / 256, 256, 0, s>>>(S, njobs, d_jobs, tb);
		ctx->stat_launches++;
		if ((rc = gd_exclusive_scan(ctx, njobs, tb, tb, ctx->sk_out2))) return rc;
		B.tile_base = tb, B.ntiles = 0;
	}
	static const int sk_ver = getenv("GDIET_SK_V") ? atoi(getenv("GDIET_SK_V")) : 3;
	static const int sk_pack = getenv("GDIET_SK_PACK") ? atoi(getenv("GDIET_SK_PACK")) : 1;   // several reads per tile (A/B switch)
	int pack_threads = 0;
	if (small && fixed_stride > 0 && sk_ver != 2 && sk_pack && pack_group > 0 && pack_group <= 32 && pack_slots > 0) {
		// whole groups per tile: the tile shape (one, two or four warps) that leaves the fewest slots empty
		double best = 0.30; // one job per one-warp tile fills about this much with 150 bp reads
		for (int T = 32; T <= 128; T *= 2) {
			int64_t R = ((int64_t)T * 8 - (S.w - 1)) / pack_slots;
			if (R * pack_group > 32) R = 32 / pack_group;
			const double util = (double)(R * pack_slots) / (T * 8);
			if (R >= 1 && util > best + 0.05) best = util, pack_threads = T, B.pack_jobs = (int32_t)(R * pack_group);
		}
		if (pack_threads) B.ntiles = (njobs + B.pack_jobs - 1) / B.pack_jobs;
	}
	if ((rc = gd_reserve(ctx, ctx->sk_state, (size_t)ntiles_bound * 8 + 64))) return rc;
	GD_CUDA_OK(ctx, cudaMemsetAsync(ctx->sk_state.p, 0, (size_t)ntiles_bound * 8 + 64, s));
	B.status = (unsigned long long *)ctx->sk_state.p;
	B.ticket = (int32_t *)((char *)ctx->sk_state.p + (size_t)ntiles_bound * 8);
	auto launch = [&](auto kern, int threads, size_t smem) -> int {
		int occ = 0;
		if (smem > 48 * 1024) GD_CUDA_OK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		GD_CUDA_OK(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
		int blocks = (int)std::min<int64_t>(B.pack_jobs > 0 ? B.ntiles : ntiles_bound, (int64_t)ctx->sms * std::max(occ, 1));
		GdKernelTimer tm(ctx, &ctx->tm_sketch);
		kern<<<std::max(blocks, 1), threads, smem, s>>>(S, B);
		return GD_OK;
	};
	if (pack_threads == 32) rc = launch(gd_sketch_tile3_kernel<32, true>, 32, sizeof(SketchSmem3<32>));
	else if (pack_threads == 64) rc = launch(gd_sketch_tile3_kernel<64, true>, 64, sizeof(SketchSmem3<64>));
	else if (pack_threads == 128) rc = launch(gd_sketch_tile3_kernel<128, true>, 128, sizeof(SketchSmem3<128>));
	else if (sk_ver != 2) {
		if (small) rc = launch(gd_sketch_tile3_kernel<32, false>, 32, sizeof(SketchSmem3<32>));
		else if (big_threads == 64) rc = launch(gd_sketch_tile3_kernel<64, false>, 64, sizeof(SketchSmem3<64>));
		else if (big_threads == 128) rc = launch(gd_sketch_tile3_kernel<128, false>, 128, sizeof(SketchSmem3<128>));
		else rc = launch(gd_sketch_tile3_kernel<256, false>, 256, sizeof(SketchSmem3<256>));
	} else if (small) rc = launch(gd_sketch_tile_kernel<32>, 32, sizeof(SketchSmem<32>));
	else if (big_threads == 64) rc = launch(gd_sketch_tile_kernel<64>, 64, sizeof(SketchSmem<64>));
	else if (big_threads == 128) rc = launch(gd_sketch_tile_kernel<128>, 128, sizeof(SketchSmem<128>));
	else rc = launch(gd_sketch_tile_kernel<256>, 256, sizeof(SketchSmem<256>));
	if (rc) return rc;
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// C ABI: index-build sketching
// --------------------------------------------------------------------------------------------
extern "C" int gd_sketch_ref_batch_device(gd_ctx *ctx, int n, const int64_t *d_off, const int32_t *d_len,
                                          const uint32_t *d_rid, const char *d_buf, int64_t total_len, int w, int k,
                                          const char *Z, int W, int64_t *d_out_off, mm128_t *d_out, int64_t out_cap)
{
	if (!ctx) return GD_ERR_ARG;
	if (n <= 0) return GD_OK;
	cudaSetDevice(ctx->device);
	SketchParams S;
	int rc = make_params(ctx, w, k, Z, W, S);
	if (rc) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_jobs, (size_t)n * sizeof(SketchJob)))) return rc;
	gd_sketch_ref_jobs_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(S, n, d_off, d_len, d_rid, (SketchJob *)ctx->sk_jobs.p);
	ctx->stat_launches++;
	// without per-sequence lengths on the host, assume the longest job may need the multi-tile path
	return gd_sketch_run_jobs(ctx, S, n, (const SketchJob *)ctx->sk_jobs.p, total_len, total_len, d_buf, d_out_off,
	                          (uint64_t *)d_out, out_cap);
}

extern "C" int gd_sketch_ref_batch(gd_ctx *ctx, int n, const int64_t *off, const int32_t *len, const uint32_t *rid,
                                   const char *buf, int w, int k, const char *Z, int W, int64_t *out_off, mm128_t *out,
                                   int64_t out_cap)
{
	if (!ctx) return GD_ERR_ARG;
	if (n < 0 || (n > 0 && (!off || !len || !buf || !out_off))) {
		ctx->err = "gd_sketch_ref_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (out_off) out_off[0] = 0;
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	SketchParams S;
	int rc = make_params(ctx, w, k, Z, W, S);
	if (rc) return rc;
	int64_t bytes = 0, max_len = 0, sum_len = 0;
	for (int i = 0; i < n; ++i) {
		if (len[i] <= 0) { // assert(len > 0) in the reference
			ctx->err = "gd_sketch_ref_batch: sequence length must be > 0";
			return GD_ERR_ARG;
		}
		bytes = std::max<int64_t>(bytes, off[i] + len[i]);
		max_len = std::max<int64_t>(max_len, len[i]), sum_len += len[i];
	}
	cudaStream_t s = ctx->stream;
	if ((rc = gd_reserve(ctx, ctx->sk_seq, (size_t)bytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_off, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_len, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_rid, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_out_off, (size_t)(n + 1) * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_jobs, (size_t)n * sizeof(SketchJob)))) return rc;
	// every sparsified base can emit at most one record
	const int64_t worst = sum_len / S.W * S.ones + (int64_t)n * S.ones + 16;
	const int64_t dcap = out ? std::min<int64_t>(worst, std::max<int64_t>(out_cap, 0)) : 0;
	if ((rc = gd_reserve(ctx, ctx->sk_out, (size_t)dcap * 16 + 16))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_seq.p, buf, (size_t)bytes, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_off.p, off, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_len.p, len, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	if (rid) GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_rid.p, rid, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	gd_sketch_ref_jobs_kernel<<<(n + 255) / 256, 256, 0, s>>>(S, n, (const int64_t *)ctx->sk_off.p, (const int32_t *)ctx->sk_len.p,
	                                                      rid ? (const uint32_t *)ctx->sk_rid.p : nullptr,
	                                                      (SketchJob *)ctx->sk_jobs.p);
	ctx->stat_launches++;
	const int64_t max_dl = max_len / S.W * S.ones + S.ones;
	rc = gd_sketch_run_jobs(ctx, S, n, (const SketchJob *)ctx->sk_jobs.p, max_dl, worst, (const char *)ctx->sk_seq.p,
	                        (int64_t *)ctx->sk_out_off.p, (uint64_t *)ctx->sk_out.p, dcap);
	if (rc) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(out_off, ctx->sk_out_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const int64_t total = out_off[n];
	if (out) {
		if (total > out_cap) {
			ctx->err = "gd_sketch_ref_batch: output buffer too small";
			return GD_ERR_CAPACITY;
		}
		if (total > 0) {
			GD_CUDA_OK(ctx, cudaMemcpyAsync(out, ctx->sk_out.p, (size_t)total * 16, cudaMemcpyDeviceToHost, s));
			GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
		}
	}
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// C ABI: read sketching (mm_sketch2 + mm_sketch3 for every shift)
// --------------------------------------------------------------------------------------------
// final lengths of every list after the reference's caps (sketch.c:2010,2219-2222)
__global__ void gd_sketch_read_counts_kernel(int n, int W, int crop, uint32_t cap2_const, uint32_t max_nb, const int32_t *len,
                                             const int64_t *job_off, const uint64_t *raw, int64_t *c3, int64_t *c2,
                                             uint32_t *s3_ret, uint32_t *s2_counts, const int32_t *job_cnt = nullptr)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int JW = W + (crop ? 1 : 0);
	const int64_t *jo = job_off + (size_t)i * JW;
	const int32_t *jc = job_cnt ? job_cnt + (size_t)i * JW : nullptr; // fixed-stride lists carry their own counts
	uint32_t cap2 = cap2_const; // 0 = uncapped
	for (int s = 0; s < W; ++s) {
		const int64_t cnt = jc ? (int64_t)jc[s] : jo[s + 1] - jo[s];
		// mm_sketch3
		int64_t n3 = (max_nb != 0 && cnt > (int64_t)max_nb) ? (int64_t)max_nb : cnt;
		c3[(size_t)i * W + s] = n3;
		s3_ret[(size_t)i * W + s] = (n3 > 0 && (uint64_t)n3 == (uint64_t)max_nb) ? (uint32_t)(raw[2 * (jo[s] + n3 - 1) + 1] >> 1)
		                                                                        : (uint32_t)len[i];
		// mm_sketch2
		int64_t n2;
		if (crop && s == 0) {
			n2 = jc ? (int64_t)jc[W] : jo[W + 1] - jo[W]; // the cropped shift-0 job, uncapped
			cap2 = (uint32_t)n2;    // becomes the cap of every later shift
		} else n2 = (cap2 != 0 && cnt > (int64_t)cap2) ? (int64_t)cap2 : cnt;
		c2[(size_t)i * W + s] = n2;
		s2_counts[(size_t)i * W + s] = (uint32_t)n2;
	}
}

__global__ void gd_sketch_read_gather_kernel(int n, int W, int crop, const int64_t *job_off, const uint64_t *raw,
                                             const int64_t *o3, const int64_t *o2, uint64_t *s3, int64_t s3_cap, uint64_t *s2,
                                             int64_t s2_cap)
{ // one warp per read
	const int warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
	const int JW = W + (crop ? 1 : 0);
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const int64_t *jo = job_off + (size_t)i * JW;
		for (int s = 0; s < W; ++s) {
			const size_t q = (size_t)i * W + s;
			const int64_t n3 = o3[q + 1] - o3[q], n2 = o2[q + 1] - o2[q];
			const int64_t src3 = jo[s], src2 = (crop && s == 0) ? jo[W] : jo[s];
			if (s3 && o3[q] + n3 <= s3_cap)
				for (int64_t e = lane; e < 2 * n3; e += 32) s3[2 * o3[q] + e] = raw[2 * src3 + e];
			if (s2 && o2[q] + n2 <= s2_cap)
				for (int64_t e = lane; e < 2 * n2; e += 32) s2[2 * o2[q] + e] = raw[2 * src2 + e];
		}
	}
}

// s2 offsets are per (read,shift) internally; the ABI exposes per-read offsets
__global__ void gd_pick_stride_kernel(int n, int W, const int64_t *in, int64_t *out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i <= n) out[i] = in[(size_t)i * W];
}

extern "C" int gd_sketch_reads_batch(gd_ctx *ctx, int n, const int64_t *off, const int32_t *len, const char *buf, int w,
                                     int k, const char *Z, int W, float max_seeds, uint32_t max_nb_seeds,
                                     uint32_t *s2_counts, int64_t *s2_off, mm128_t *s2, int64_t s2_cap, int64_t *s3_off,
                                     uint32_t *s3_ret, mm128_t *s3, int64_t s3_cap)
{
	if (!ctx) return GD_ERR_ARG;
	if (n < 0 || (n > 0 && (!off || !len || !buf))) {
		ctx->err = "gd_sketch_reads_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (s2_off) s2_off[0] = 0;
	if (s3_off) s3_off[0] = 0;
	if (W == 1) { // mm_sketch2 with a pattern of length 1 walks one shift more than the pattern has (sketch.c:2143-2225): not pinned -> refused
		ctx->err = "gd_sketch_reads_batch: pattern length W = 1 is not supported (the reference's mm_sketch2 is not pinned for it)";
		return GD_ERR_ARG;
	}
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	SketchParams S;
	int rc = make_params(ctx, w, k, Z, W, S);
	if (rc) return rc;
	int64_t bytes = 0, max_len = 0, sum_len = 0;
	for (int i = 0; i < n; ++i) {
		if (len[i] < W) {
			ctx->err = "gd_sketch_reads_batch: read shorter than the pattern";
			return GD_ERR_ARG;
		}
		bytes = std::max<int64_t>(bytes, off[i] + len[i]);
		max_len = std::max<int64_t>(max_len, len[i]), sum_len += len[i];
	}
	const int crop = max_seeds < 1.0f ? 1 : 0;
	cudaStream_t s = ctx->stream;
	if ((rc = gd_reserve(ctx, ctx->sk_seq, (size_t)bytes + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_off, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_len, (size_t)n * 4))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_seq.p, buf, (size_t)bytes, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_off.p, off, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->sk_len.p, len, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	// every (read, shift) job + cap rules on the device (shared with the mapping stage)
	GdReadSketch K;
	if ((rc = gd_sketch_reads_device_raw(ctx, n, (const int64_t *)ctx->sk_off.p, (const int32_t *)ctx->sk_len.p, (const char *)ctx->sk_seq.p,
	                                     max_len, sum_len, w, k, Z, W, max_seeds, max_nb_seeds, &K)))
		return rc;
	// offsets, gather
	const size_t nq = (size_t)n * W;
	int64_t *c3 = const_cast<int64_t *>(K.c3), *c2 = const_cast<int64_t *>(K.c2);
	uint32_t *d_ret = const_cast<uint32_t *>(K.ret3), *d_cnt2 = d_ret + nq;
	int64_t *d_off2 = (int64_t *)(((uintptr_t)(d_cnt2 + nq) + 15) & ~(uintptr_t)15);
	if ((rc = gd_exclusive_scan(ctx, (int64_t)nq, c3, c3, ctx->sk_out2))) return rc;
	if ((rc = gd_exclusive_scan(ctx, (int64_t)nq, c2, c2, ctx->sk_out2))) return rc;
	gd_pick_stride_kernel<<<(n + 1 + 255) / 256, 256, 0, s>>>(n, W, c2, d_off2);
	ctx->stat_launches++;
	if (s3_off) GD_CUDA_OK(ctx, cudaMemcpyAsync(s3_off, c3, (nq + 1) * 8, cudaMemcpyDeviceToHost, s));
	if (s2_off) GD_CUDA_OK(ctx, cudaMemcpyAsync(s2_off, d_off2, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
	if (s3_ret) GD_CUDA_OK(ctx, cudaMemcpyAsync(s3_ret, d_ret, nq * 4, cudaMemcpyDeviceToHost, s));
	if (s2_counts) GD_CUDA_OK(ctx, cudaMemcpyAsync(s2_counts, d_cnt2, nq * 4, cudaMemcpyDeviceToHost, s));
	int64_t tot[2] = {0, 0};
	GD_CUDA_OK(ctx, cudaMemcpyAsync(&tot[0], c3 + nq, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(&tot[1], c2 + nq, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	if ((s3 && tot[0] > s3_cap) || (s2 && tot[1] > s2_cap)) {
		ctx->err = "gd_sketch_reads_batch: output buffer too small";
		return GD_ERR_CAPACITY;
	}
	if (s3 || s2) {
		// gather into device staging (reuse sk_out2 beyond the scan scratch), then copy out
		const size_t g3 = s3 ? (size_t)tot[0] * 16 : 0, g2 = s2 ? (size_t)tot[1] * 16 : 0;
		if ((rc = gd_reserve(ctx, ctx->sk_state, g3 + g2 + 64))) return rc;
		uint64_t *d3 = (uint64_t *)ctx->sk_state.p, *d2 = (uint64_t *)((char *)ctx->sk_state.p + g3);
		int blocks = std::min((n + 7) / 8, ctx->sms * 8);
		gd_sketch_read_gather_kernel<<<std::max(blocks, 1), 256, 0, s>>>(n, W, crop, K.job_off, K.raw, c3, c2, s3 ? d3 : nullptr,
		                                                            tot[0], s2 ? d2 : nullptr, tot[1]);
		ctx->stat_launches++;
		GD_CUDA_OK(ctx, cudaGetLastError());
		if (g3) GD_CUDA_OK(ctx, cudaMemcpyAsync(s3, d3, g3, cudaMemcpyDeviceToHost, s));
		if (g2) GD_CUDA_OK(ctx, cudaMemcpyAsync(s2, d2, g2, cudaMemcpyDeviceToHost, s));
		GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	}
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// device-resident read sketching for the mapper (gd_map.cu): every (read, shift) job plus the cropped job of
// mm_sketch2, cap rules applied, nothing copied or synchronised; the lists stay in the context's sketch scratch
// (list of (read i, shift s): raw[job_off[i*JW+s] ..], first c3 / c2 entries; mm_sketch2's shift-0 list is job W).
// --------------------------------------------------------------------------------------------
int gd_sketch_reads_device_raw(gd_ctx *ctx, int n, const int64_t *d_off, const int32_t *d_len, const char *d_buf,
                               int64_t max_len, int64_t sum_len, int w, int k, const char *Z, int W, float max_seeds,
                               uint32_t max_nb_seeds, GdReadSketch *out)
{
	SketchParams S;
	int rc = make_params(ctx, w, k, Z, W, S);
	if (rc) return rc;
	const int crop = max_seeds < 1.0f ? 1 : 0;
	const int JW = W + crop;
	const uint32_t cap2_const = crop ? 0u : (uint32_t)max_seeds;
	const int64_t njobs64 = (int64_t)n * JW;
	if (njobs64 > 0x7fffffff) {
		ctx->err = "read sketching: too many (read,shift) jobs in one call";
		return GD_ERR_ARG;
	}
	const int njobs = (int)njobs64;
	cudaStream_t s = ctx->stream;
	const int64_t per_job = max_len / S.W * S.ones + S.ones;
	int64_t worst = (sum_len / S.W * S.ones + (int64_t)n * S.ones) * JW + 16;
	// Short reads (one warp-sized tile per job): every job gets its own slot of per_job records -- a position emits at most
	// once -- so the tiles need neither the ordering ticket nor the look-back; the lists are addressed by job_off / counts.
	const int tp_small = sk_tile_emit(256, S.w, S.k);
	const bool fixed = tp_small > 0 && per_job <= tp_small && (int64_t)njobs * per_job * 16 <= (24ll << 30);
	int32_t *d_cnt = nullptr;
	if (fixed) {
		worst = (int64_t)njobs * per_job;
		if ((rc = gd_reserve(ctx, ctx->sk_misc, (size_t)njobs * 4 + 16))) return rc;
		d_cnt = (int32_t *)ctx->sk_misc.p;
	}
	if ((rc = gd_reserve(ctx, ctx->sk_jobs, (size_t)njobs * sizeof(SketchJob)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_out_off, (size_t)(njobs + 1) * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_out, (size_t)worst * 16 + 64))) return rc;
	gd_sketch_read_jobs_kernel<<<(n + 255) / 256, 256, 0, s>>>(S, n, d_off, d_len, W, crop, max_seeds, (SketchJob *)ctx->sk_jobs.p);
	ctx->stat_launches++;
	// slots a read needs in a packed tile: its W full-length jobs, and the cropped job if that one is long enough to emit
	const int64_t crop_dl = crop ? (int64_t)((double)max_seeds * max_len) / S.W * S.ones + S.ones : 0;
	const int64_t read_slots = (int64_t)W * (per_job + 1) + (crop && crop_dl >= S.w + S.k - 1 ? crop_dl + 1 : 0);
	rc = gd_sketch_run_jobs(ctx, S, njobs, (const SketchJob *)ctx->sk_jobs.p, per_job, worst, d_buf, (int64_t *)ctx->sk_out_off.p,
	                        (uint64_t *)ctx->sk_out.p, worst, fixed ? per_job : 0, d_cnt, JW, read_slots);
	if (rc) return rc;
	const size_t nq = (size_t)n * W;
	const size_t need = (nq + 1) * 8 * 2 + nq * 4 * 2 + (size_t)(n + 1) * 8 + 128; // c3 | c2 | ret | cnt2 | per-read offsets (host API)
	if ((rc = gd_reserve(ctx, ctx->sk_rid, need))) return rc;
	int64_t *c3 = (int64_t *)ctx->sk_rid.p, *c2 = c3 + nq + 1;
	uint32_t *d_ret = (uint32_t *)(c2 + nq + 1), *d_cnt2 = d_ret + nq;
	gd_sketch_read_counts_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, W, crop, cap2_const, max_nb_seeds, d_len, (const int64_t *)ctx->sk_out_off.p,
	                                                         (const uint64_t *)ctx->sk_out.p, c3, c2, d_ret, d_cnt2, d_cnt);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	out->job_off = (const int64_t *)ctx->sk_out_off.p, out->raw = (const uint64_t *)ctx->sk_out.p;
	out->c3 = c3, out->c2 = c2, out->ret3 = d_ret, out->JW = JW, out->crop = crop, out->raw_cap = worst;
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// drop-in single-sequence entry points (GDiet-ShortReads/mmpriv.h:63-68)
// --------------------------------------------------------------------------------------------
static void fatal(gd_ctx *ctx, const char *what)
{
	fprintf(stderr, "[gdiet_cuda] FATAL: %s: %s\n", what, gd_strerror(ctx));
	abort();
}

static void append_v(void *km, mm128_v *p, const mm128_t *src, size_t n)
{ // same growth as kv_push (kvec.h:79-85): capacity doubles from 2
	if (p->n + n > p->m) {
		size_t m = p->m ? p->m : 2;
		while (m < p->n + n) m <<= 1;
		p->a = (mm128_t *)gd_host_realloc(km, p->a, sizeof(mm128_t) * m);
		p->m = m;
	}
	if (n) memcpy(p->a + p->n, src, n * sizeof(mm128_t));
	p->n += n;
}

extern "C" void mm_sketch(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p,
                          const char *Z, int W)
{
	(void)is_hpc; // ignored by the reference as well (sketch.c:1647-1658)
	gd_ctx *ctx = gd_thread_ctx();
	const int64_t off = 0;
	int64_t oo[2] = {0, 0};
	int64_t cap = len + 16;
	std::vector<mm128_t> tmp((size_t)cap);
	if (gd_sketch_ref_batch(ctx, 1, &off, &len, &rid, str, w, k, Z, W, oo, tmp.data(), cap) != GD_OK) fatal(ctx, "mm_sketch");
	append_v(km, p, tmp.data(), (size_t)oo[1]);
}

extern "C" unsigned mm_sketch3(void *km, const char *str, const unsigned len, int w, int k, uint32_t rid, int is_hpc,
                               mm128_v *p, const char *Z, int W, int shift2, uint32_t MAX_NB_SEEDS)
{
	(void)is_hpc;
	gd_ctx *ctx = gd_thread_ctx();
	const int64_t off = 0;
	const int32_t l = (int32_t)len;
	const int shift = shift2 < 0 ? 0 : shift2;
	std::vector<int64_t> o3((size_t)W + 1);
	std::vector<uint32_t> ret((size_t)W);
	const int64_t cap = ((int64_t)len + 16) * W;
	std::vector<mm128_t> tmp((size_t)cap);
	// the cap of the reference counts entries already in *p (sketch.c:2010); callers pass p->n == 0
	uint32_t eff = MAX_NB_SEEDS;
	if (p->n && MAX_NB_SEEDS > p->n) eff = MAX_NB_SEEDS - (uint32_t)p->n;
	if (gd_sketch_reads_batch(ctx, 1, &off, &l, str, w, k, Z, W, 1.0f, eff, nullptr, nullptr, nullptr, 0, o3.data(),
	                          ret.data(), tmp.data(), cap) != GD_OK)
		fatal(ctx, "mm_sketch3");
	if (shift >= W) return len;
	const size_t n0 = (size_t)(o3[shift + 1] - o3[shift]);
	size_t base = p->n;
	append_v(km, p, tmp.data() + o3[shift], n0);
	if (rid) // read sketching always passes rid 0 (map.c:81); honour other values anyway
		for (size_t i = base; i < p->n; ++i) p->a[i].y |= (uint64_t)rid << 32;
	return ret[shift];
}

extern "C" mm_pattern_t mm_sketch2(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p,
                                   const char *Z, int W, const float max_seeds)
{
	(void)is_hpc;
	gd_ctx *ctx = gd_thread_ctx();
	mm_pattern_t pat;
	pat.n = (uint32_t)W;
	pat.shift_seeds_number = (uint32_t *)gd_host_malloc(km, (size_t)W * sizeof(uint32_t));
	const int64_t off = 0;
	int64_t o2[2] = {0, 0};
	const int64_t cap = ((int64_t)len + 16) * W;
	std::vector<mm128_t> tmp((size_t)cap);
	if (gd_sketch_reads_batch(ctx, 1, &off, &len, str, w, k, Z, W, max_seeds, 0, pat.shift_seeds_number, o2, tmp.data(), cap,
	                          nullptr, nullptr, nullptr, 0) != GD_OK)
		fatal(ctx, "mm_sketch2");
	size_t base = p->n;
	append_v(km, p, tmp.data(), (size_t)o2[1]);
	if (rid)
		for (size_t i = base; i < p->n; ++i) p->a[i].y |= (uint64_t)rid << 32;
	return pat;
}
