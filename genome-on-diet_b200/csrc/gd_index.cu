// gd_index.cu -- device-side index build and lookup: replaces the sketch -> bucket -> sort ->
// hash pipeline of mm_idx_gen (GDiet-ShortReads/index.c:216-304,389-420) and mm_idx_get (index.c:84-100).
//
// Build, all on the device: contigs are sketched by the sketch kernel (output already ascending in
// y = rid<<32|pos<<1|strand), ONE stable radix sort by minimizer value groups equal minimizers and keeps their
// positions ascending (what radix_sort_128x + radix_sort_64 give per bucket, index.c:225,255), a run-length pass
// yields the distinct minimizers, and one kernel inserts {minimizer, first, count} into the open-addressing
// table.  The sort / run-length / scan primitives are CUB (library code, like cuBLAS for a plain GEMM); the
// sketching, the table and every lookup are ours.
#include "gd_ctx.h"
#include "gd_index.cuh"
#include "gd_sketch.cuh"
#include <cub/cub.cuh>
#include <algorithm>
#include <string.h>
#include <string>
#include <vector>

using namespace gd;

#define IDX_CUDA(call)                                                                            \
	do {                                                                                          \
		cudaError_t e__ = (call);                                                                 \
		if (e__ != cudaSuccess) {                                                                 \
			char b__[512];                                                                        \
			snprintf(b__, sizeof b__, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
			ctx->err = b__;                                                                       \
			rc = GD_ERR_CUDA;                                                                     \
			goto fail;                                                                            \
		}                                                                                         \
	} while (0)

// --------------------------------------------------------------------------------------------
// kernels
// --------------------------------------------------------------------------------------------
__global__ void gd_idx_split_kernel(int64_t n, const uint64_t *xy, uint64_t *key, uint64_t *val)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const ulonglong2 r = *(const ulonglong2 *)(xy + 2 * i);
	key[i] = r.x >> 8; // the span byte is k for every record (sketch.c:62-67)
	val[i] = r.y;
}

__global__ void gd_idx_clear_kernel(int64_t slots, IdxSlot *tab)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < slots) tab[i].key = GD_IDX_EMPTY, tab[i].first = 0, tab[i].count = 0;
}

__global__ void gd_idx_insert_kernel(int64_t n_keys, const uint64_t *keys, const uint32_t *counts, const uint64_t *first,
                                     IdxSlot *tab, uint64_t mask)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_keys) return;
	const unsigned long long key = keys[i];
	uint64_t s = idx_mix(key) & mask;
	for (;;) {
		const unsigned long long old = atomicCAS(&tab[s].key, GD_IDX_EMPTY, key);
		if (old == GD_IDX_EMPTY) {
			tab[s].first = (uint32_t)first[i], tab[s].count = counts[i];
			return;
		}
		s = (s + 1) & mask;
	}
}

// mi->S: nt4 code of every base, 4 bits each, contigs back to back (index.c:351-356)
__global__ void gd_idx_pack_kernel(int64_t total, int n_seq, const uint64_t *seq_off, const int64_t *src_off, const char *buf,
                                   uint32_t *S)
{
	const int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	const uint64_t g0 = (uint64_t)wi * 8;
	if ((int64_t)g0 >= total) return;
	int lo = 0, hi = n_seq - 1; // last contig with seq_off <= g0
	while (lo < hi) {
		const int mid = (lo + hi + 1) >> 1;
		if (seq_off[mid] <= g0) lo = mid;
		else hi = mid - 1;
	}
	int c = lo;
	uint32_t v = 0;
	for (int j = 0; j < 8; ++j) {
		const uint64_t g = g0 + j;
		if ((int64_t)g >= total) break;
		while (c + 1 < n_seq && seq_off[c + 1] <= g) ++c;
		v |= (uint32_t)sk_nt4((unsigned char)buf[src_off[c] + (int64_t)(g - seq_off[c])]) << (4 * j);
	}
	S[wi] = v;
}

__global__ void gd_idx_get_kernel(IndexDev I, int64_t n, const uint64_t *minier, uint32_t *count, int64_t *first)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t f;
	const uint32_t c = idx_get(I, minier[i], f);
	count[i] = c, first[i] = c ? (int64_t)f : -1;
}

// --------------------------------------------------------------------------------------------
// build
// --------------------------------------------------------------------------------------------
extern "C" void gd_index_destroy(gd_index *idx)
{
	if (!idx) return;
	cudaSetDevice(idx->device);
	void *bufs[] = {idx->d_tab, idx->d_pos, idx->d_S, idx->d_seq_off, idx->d_seq_len, idx->d_keys, idx->d_counts};
	for (void *p : bufs)
		if (p) cudaFree(p);
	free(idx->h_seq_len), free(idx->h_seq_off);
	delete idx;
}

// off / len are host arrays; the ASCII sequence is in host memory (buf) or already in HBM (d_seq, not modified)
static int index_build_core(gd_ctx *ctx, int n_seq, const int64_t *off, const int32_t *len, const char *buf, const char *d_seq,
                            int w, int k, const char *Z, int W, gd_index **out)
{
	if (!ctx) return GD_ERR_ARG;
	if (!out || n_seq <= 0 || !off || !len || (!buf && !d_seq) || !Z) {
		ctx->err = "gd_index_build: bad argument";
		return GD_ERR_ARG;
	}
	*out = nullptr;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	int rc = GD_OK;
	int64_t total = 0, bytes = 0;
	std::vector<uint64_t> seq_off((size_t)n_seq + 1);
	for (int i = 0; i < n_seq; ++i) {
		if (len[i] <= 0 || off[i] < 0) {
			ctx->err = "gd_index_build: contig length must be > 0";
			return GD_ERR_ARG;
		}
		seq_off[i] = (uint64_t)total, total += len[i];
		bytes = std::max<int64_t>(bytes, off[i] + len[i]);
	}
	seq_off[n_seq] = (uint64_t)total;
	int ones = 0;
	for (int g = 0; g < W; ++g) ones += Z[g] == '1';
	if (ones == 0 || W <= 0) {
		ctx->err = "gd_index_build: bad pattern";
		return GD_ERR_ARG;
	}
	gd_index *idx = new gd_index();
	idx->device = ctx->device, idx->n_seq = n_seq, idx->total_len = total;
	idx->h_seq_len = (uint32_t *)malloc((size_t)n_seq * 4), idx->h_seq_off = (uint64_t *)malloc((size_t)(n_seq + 1) * 8);
	for (int i = 0; i < n_seq; ++i) idx->h_seq_len[i] = (uint32_t)len[i];
	memcpy(idx->h_seq_off, seq_off.data(), (size_t)(n_seq + 1) * 8);
	// temporaries of the build (freed before returning; the context's grow-only buffers are not used for them)
	char *d_buf = nullptr;
	int64_t *d_off = nullptr, *d_out_off = nullptr;
	int32_t *d_len = nullptr;
	uint64_t *d_xy = nullptr, *d_key = nullptr, *d_val = nullptr, *d_key2 = nullptr, *d_first = nullptr;
	void *d_tmp = nullptr, *d_nruns = nullptr;
	size_t tmp_bytes = 0;
	int64_t h_total = 0, cap = 0;
	const int64_t worst = total / W * ones + (int64_t)n_seq * ones + 16;
	if (!d_seq) IDX_CUDA(cudaMalloc(&d_buf, (size_t)bytes + 16));
	IDX_CUDA(cudaMalloc(&d_off, (size_t)n_seq * 8));
	IDX_CUDA(cudaMalloc(&d_len, (size_t)n_seq * 4));
	IDX_CUDA(cudaMalloc(&d_out_off, (size_t)(n_seq + 1) * 8));
	IDX_CUDA(cudaMalloc(&idx->d_seq_off, (size_t)(n_seq + 1) * 8));
	IDX_CUDA(cudaMalloc(&idx->d_seq_len, (size_t)n_seq * 4));
	if (!d_seq) IDX_CUDA(cudaMemcpyAsync(d_buf, buf, (size_t)bytes, cudaMemcpyHostToDevice, s));
	else d_buf = const_cast<char *>(d_seq);
	IDX_CUDA(cudaMemcpyAsync(d_off, off, (size_t)n_seq * 8, cudaMemcpyHostToDevice, s));
	IDX_CUDA(cudaMemcpyAsync(d_len, len, (size_t)n_seq * 4, cudaMemcpyHostToDevice, s));
	IDX_CUDA(cudaMemcpyAsync(idx->d_seq_off, seq_off.data(), (size_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, s));
	IDX_CUDA(cudaMemcpyAsync(idx->d_seq_len, idx->h_seq_len, (size_t)n_seq * 4, cudaMemcpyHostToDevice, s));
	// ---- 4-bit reference
	idx->s_words = (total + 7) / 8;
	IDX_CUDA(cudaMalloc(&idx->d_S, (size_t)idx->s_words * 4 + 16));
	gd_idx_pack_kernel<<<(unsigned)((idx->s_words + 255) / 256), 256, 0, s>>>(total, n_seq, (const uint64_t *)idx->d_seq_off, d_off,
	                                                                        d_buf, (uint32_t *)idx->d_S);
	ctx->stat_launches++;
	// ---- sketch every contig (expected density 2/(w+1) per sparsified base; retry with the hard bound if exceeded)
	cap = std::min<int64_t>(worst, (int64_t)((double)worst * 2.0 / (w + 1) * 1.25) + 4096 + (int64_t)n_seq * 8);
	for (int attempt = 0; attempt < 2; ++attempt) {
		IDX_CUDA(cudaMalloc(&d_xy, (size_t)cap * 16 + 16));
		rc = gd_sketch_ref_batch_device(ctx, n_seq, d_off, d_len, nullptr, d_buf, total, w, k, Z, W, d_out_off, (mm128_t *)d_xy, cap);
		if (rc) goto fail;
		IDX_CUDA(cudaMemcpyAsync(&h_total, d_out_off + n_seq, 8, cudaMemcpyDeviceToHost, s));
		IDX_CUDA(cudaStreamSynchronize(s));
		if (h_total <= cap) break;
		cudaFree(d_xy), d_xy = nullptr, cap = worst;
	}
	if (!d_seq) cudaFree(d_buf);
	d_buf = nullptr;
	idx->n_min = h_total;
	if (h_total > 0xffffffffll) {
		ctx->err = "gd_index_build: more than 2^32 minimizers (split the reference like -I does)";
		rc = GD_ERR_ARG;
		goto fail;
	}
	// ---- group by minimizer (stable: positions of one minimizer stay ascending)
	{
		const int64_t n = h_total;
		IDX_CUDA(cudaMalloc(&d_key, (size_t)(n + 1) * 8));
		IDX_CUDA(cudaMalloc(&d_val, (size_t)(n + 1) * 8));
		IDX_CUDA(cudaMalloc(&d_key2, (size_t)(n + 1) * 8));
		IDX_CUDA(cudaMalloc(&idx->d_pos, (size_t)(n + 1) * 8));
		if (n > 0) gd_idx_split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(n, d_xy, d_key, d_val);
		ctx->stat_launches++;
		cudaStreamSynchronize(s);
		cudaFree(d_xy), d_xy = nullptr;
		IDX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_key, d_key2, d_val, (uint64_t *)idx->d_pos, n, 0, 2 * k, s));
		IDX_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
		IDX_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_key, d_key2, d_val, (uint64_t *)idx->d_pos, n, 0, 2 * k, s));
		cudaStreamSynchronize(s);
		cudaFree(d_tmp), d_tmp = nullptr;
		cudaFree(d_val), d_val = nullptr;
		// distinct minimizers + counts -> d_key (reused) / d_counts
		IDX_CUDA(cudaMalloc(&idx->d_counts, (size_t)(n + 1) * 4));
		IDX_CUDA(cudaMalloc(&d_nruns, 8));
		IDX_CUDA(cudaMemsetAsync(d_nruns, 0, 8, s));
		IDX_CUDA(cub::DeviceRunLengthEncode::Encode(nullptr, tmp_bytes, d_key2, d_key, (uint32_t *)idx->d_counts, (int64_t *)d_nruns, n, s));
		IDX_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
		IDX_CUDA(cub::DeviceRunLengthEncode::Encode(d_tmp, tmp_bytes, d_key2, d_key, (uint32_t *)idx->d_counts, (int64_t *)d_nruns, n, s));
		IDX_CUDA(cudaMemcpyAsync(&idx->n_keys, d_nruns, 8, cudaMemcpyDeviceToHost, s));
		IDX_CUDA(cudaStreamSynchronize(s));
		cudaFree(d_tmp), d_tmp = nullptr;
		cudaFree(d_key2), d_key2 = nullptr;
		const int64_t nk = idx->n_keys;
		// first position of every minimizer
		IDX_CUDA(cudaMalloc(&d_first, (size_t)(nk + 1) * 8));
		IDX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (const uint32_t *)idx->d_counts, d_first, nk, s));
		IDX_CUDA(cudaMalloc(&d_tmp, tmp_bytes + 16));
		IDX_CUDA(cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, (const uint32_t *)idx->d_counts, d_first, nk, s));
		// table
		int64_t slots = 1024;
		while (slots < 2 * nk) slots <<= 1;
		idx->tab_slots = slots;
		IDX_CUDA(cudaMalloc(&idx->d_tab, (size_t)slots * sizeof(IdxSlot)));
		gd_idx_clear_kernel<<<(unsigned)((slots + 255) / 256), 256, 0, s>>>(slots, (IdxSlot *)idx->d_tab);
		if (nk > 0)
			gd_idx_insert_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, s>>>(nk, d_key, (const uint32_t *)idx->d_counts, d_first,
			                                                               (IdxSlot *)idx->d_tab, (uint64_t)slots - 1);
		ctx->stat_launches += 2;
		IDX_CUDA(cudaGetLastError());
		IDX_CUDA(cudaStreamSynchronize(s));
		cudaFree(d_tmp), d_tmp = nullptr;
		cudaFree(d_first), d_first = nullptr;
		idx->d_keys = d_key, d_key = nullptr; // keep the sorted distinct minimizers (export, max_occ)
		idx->device_bytes = (size_t)slots * sizeof(IdxSlot) + (size_t)(n + 1) * 8 + (size_t)idx->s_words * 4 + (size_t)(n + 1) * 12;
	}
	cudaFree(d_off), cudaFree(d_len), cudaFree(d_out_off), cudaFree(d_nruns);
	idx->d.tab = (const IdxSlot *)idx->d_tab, idx->d.tab_mask = (uint64_t)idx->tab_slots - 1;
	idx->d.pos = (const uint64_t *)idx->d_pos, idx->d.S = (const uint32_t *)idx->d_S;
	idx->d.seq_off = (const uint64_t *)idx->d_seq_off, idx->d.seq_len = (const uint32_t *)idx->d_seq_len;
	idx->d.n_seq = n_seq, idx->d.w = w, idx->d.k = k;
	*out = idx;
	return GD_OK;
fail:
	cudaStreamSynchronize(s);
	{
		void *tmps[] = {d_seq ? nullptr : (void *)d_buf, d_off, d_len, d_out_off, d_xy, d_key, d_val, d_key2, d_first, d_tmp, d_nruns};
		for (void *p : tmps)
			if (p) cudaFree(p);
	}
	gd_index_destroy(idx);
	return rc;
}

extern "C" int gd_index_build(gd_ctx *ctx, int n_seq, const int64_t *off, const int32_t *len, const char *buf, int w, int k,
                              const char *Z, int W, gd_index **out)
{
	return index_build_core(ctx, n_seq, off, len, buf, nullptr, w, k, Z, W, out);
}

extern "C" int gd_index_build_device(gd_ctx *ctx, int n_seq, const int64_t *off, const int32_t *len, const char *d_buf, int w,
                                     int k, const char *Z, int W, gd_index **out)
{
	return index_build_core(ctx, n_seq, off, len, nullptr, d_buf, w, k, Z, W, out);
}

extern "C" int64_t gd_index_stat(const gd_index *idx, const char *key)
{
	if (!idx || !key) return -1;
	if (!strcmp(key, "n_seq")) return idx->n_seq;
	if (!strcmp(key, "total_len")) return idx->total_len;
	if (!strcmp(key, "n_minimizers")) return idx->n_min;
	if (!strcmp(key, "n_keys")) return idx->n_keys;
	if (!strcmp(key, "table_slots")) return idx->tab_slots;
	if (!strcmp(key, "device_bytes")) return (int64_t)idx->device_bytes;
	if (!strcmp(key, "s_words")) return idx->s_words;
	return -1;
}

extern "C" int gd_index_get_batch(gd_ctx *ctx, const gd_index *idx, int64_t n, const uint64_t *minier, uint32_t *count,
                                  int64_t *first)
{
	if (!ctx) return GD_ERR_ARG;
	if (!idx || n < 0 || (n > 0 && (!minier || !count || !first))) {
		ctx->err = "gd_index_get_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	int rc;
	if ((rc = gd_reserve(ctx, ctx->sk_misc, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_out2, (size_t)n * 12))) return rc;
	uint64_t *d_m = (uint64_t *)ctx->sk_misc.p;
	int64_t *d_first = (int64_t *)ctx->sk_out2.p;
	uint32_t *d_cnt = (uint32_t *)(d_first + n);
	GD_CUDA_OK(ctx, cudaMemcpyAsync(d_m, minier, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	gd_idx_get_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(idx->d, n, d_m, d_cnt, d_first);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	GD_CUDA_OK(ctx, cudaMemcpyAsync(count, d_cnt, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(first, d_first, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	return GD_OK;
}

extern "C" int gd_index_export(gd_ctx *ctx, const gd_index *idx, uint64_t *keys, uint32_t *counts, uint64_t *positions,
                               uint32_t *S)
{
	if (!ctx || !idx) return GD_ERR_ARG;
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	if (keys && idx->n_keys) GD_CUDA_OK(ctx, cudaMemcpyAsync(keys, idx->d_keys, (size_t)idx->n_keys * 8, cudaMemcpyDeviceToHost, s));
	if (counts && idx->n_keys) GD_CUDA_OK(ctx, cudaMemcpyAsync(counts, idx->d_counts, (size_t)idx->n_keys * 4, cudaMemcpyDeviceToHost, s));
	if (positions && idx->n_min) GD_CUDA_OK(ctx, cudaMemcpyAsync(positions, idx->d_pos, (size_t)idx->n_min * 8, cudaMemcpyDeviceToHost, s));
	if (S && idx->s_words) GD_CUDA_OK(ctx, cudaMemcpyAsync(S, idx->d_S, (size_t)idx->s_words * 4, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	return GD_OK;
}

// mm_idx_cal_max_occ (index.c:182-201): thres = (k-th smallest occurrence count, k = (uint32)((1-f) n)) + 1
extern "C" int gd_index_cal_max_occ(gd_ctx *ctx, const gd_index *idx, float frac, int32_t *max_occ)
{
	if (!ctx || !idx || !max_occ) return GD_ERR_ARG;
	if (frac <= 0.f) {
		*max_occ = INT32_MAX;
		return GD_OK;
	}
	const int64_t n = idx->n_keys;
	if (n == 0) {
		*max_occ = 1;
		return GD_OK;
	}
	cudaSetDevice(ctx->device);
	cudaStream_t s = ctx->stream;
	int rc;
	size_t tmp_bytes = 0;
	GD_CUDA_OK(ctx, cub::DeviceRadixSort::SortKeys(nullptr, tmp_bytes, (const uint32_t *)idx->d_counts, (uint32_t *)nullptr, n, 0, 32, s));
	if ((rc = gd_reserve(ctx, ctx->sk_out2, (size_t)n * 4 + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->sk_misc, tmp_bytes + 16))) return rc;
	GD_CUDA_OK(ctx, cub::DeviceRadixSort::SortKeys(ctx->sk_misc.p, tmp_bytes, (const uint32_t *)idx->d_counts, (uint32_t *)ctx->sk_out2.p, n,
	                                               0, 32, s));
	const uint32_t kth = (uint32_t)((1. - frac) * n);
	uint32_t v = 0;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(&v, (uint32_t *)ctx->sk_out2.p + std::min<int64_t>(kth, n - 1), 4, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	*max_occ = (int32_t)(v + 1);
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// replication onto the other GPUs of the node: the index is built once, its device buffers are
// broadcast by the caller's communicator (ncclBroadcast over NVLink), every GPU then maps its own read shard
// --------------------------------------------------------------------------------------------
extern "C" int gd_index_meta(const gd_index *idx, gd_index_meta_t *m)
{
	if (!idx || !m) return GD_ERR_ARG;
	m->n_seq = idx->n_seq, m->total_len = idx->total_len, m->n_minimizers = idx->n_min, m->n_keys = idx->n_keys;
	m->table_slots = idx->tab_slots, m->s_words = idx->s_words, m->w = idx->d.w, m->k = idx->d.k;
	return GD_OK;
}

static void index_buffer_list(gd_index *idx, void **ptrs, size_t *bytes)
{
	ptrs[0] = idx->d_tab, bytes[0] = (size_t)idx->tab_slots * sizeof(IdxSlot);
	ptrs[1] = idx->d_pos, bytes[1] = (size_t)idx->n_min * 8;
	ptrs[2] = idx->d_S, bytes[2] = (size_t)idx->s_words * 4;
	ptrs[3] = idx->d_seq_off, bytes[3] = (size_t)(idx->n_seq + 1) * 8;
	ptrs[4] = idx->d_seq_len, bytes[4] = (size_t)idx->n_seq * 4;
	ptrs[5] = idx->d_keys, bytes[5] = (size_t)idx->n_keys * 8;
	ptrs[6] = idx->d_counts, bytes[6] = (size_t)idx->n_keys * 4;
}

extern "C" int gd_index_buffers(gd_index *idx, void **ptrs, size_t *bytes)
{
	if (!idx || !ptrs || !bytes) return GD_ERR_ARG;
	index_buffer_list(idx, ptrs, bytes);
	return GD_OK;
}

extern "C" int gd_index_alloc(gd_ctx *ctx, const gd_index_meta_t *m, gd_index **out)
{
	if (!ctx || !m || !out || m->n_seq <= 0 || m->table_slots <= 0 || (m->table_slots & (m->table_slots - 1))) {
		if (ctx) ctx->err = "gd_index_alloc: bad argument";
		return GD_ERR_ARG;
	}
	*out = nullptr;
	cudaSetDevice(ctx->device);
	gd_index *idx = new gd_index();
	idx->device = ctx->device, idx->n_seq = m->n_seq, idx->total_len = m->total_len, idx->n_min = m->n_minimizers;
	idx->n_keys = m->n_keys, idx->tab_slots = m->table_slots, idx->s_words = m->s_words;
	idx->d.w = m->w, idx->d.k = m->k, idx->d.n_seq = (int32_t)m->n_seq;
	void *ptrs[GD_INDEX_NBUF];
	size_t bytes[GD_INDEX_NBUF];
	index_buffer_list(idx, ptrs, bytes);
	void **slots[GD_INDEX_NBUF] = {&idx->d_tab, &idx->d_pos, &idx->d_S, &idx->d_seq_off, &idx->d_seq_len, &idx->d_keys, &idx->d_counts};
	for (int i = 0; i < GD_INDEX_NBUF; ++i) {
		if (cudaMalloc(slots[i], bytes[i] + 16) != cudaSuccess) {
			ctx->err = "gd_index_alloc: out of device memory";
			gd_index_destroy(idx);
			return GD_ERR_CUDA;
		}
		idx->device_bytes += bytes[i];
	}
	*out = idx;
	return GD_OK;
}

extern "C" int gd_index_commit(gd_ctx *ctx, gd_index *idx)
{ // after the buffers of an allocated index have been filled: host copies of the contig table, kernel view
	if (!ctx || !idx) return GD_ERR_ARG;
	cudaSetDevice(ctx->device);
	free(idx->h_seq_len), free(idx->h_seq_off);
	idx->h_seq_len = (uint32_t *)malloc((size_t)idx->n_seq * 4), idx->h_seq_off = (uint64_t *)malloc((size_t)(idx->n_seq + 1) * 8);
	GD_CUDA_OK(ctx, cudaMemcpy(idx->h_seq_len, idx->d_seq_len, (size_t)idx->n_seq * 4, cudaMemcpyDeviceToHost));
	GD_CUDA_OK(ctx, cudaMemcpy(idx->h_seq_off, idx->d_seq_off, (size_t)(idx->n_seq + 1) * 8, cudaMemcpyDeviceToHost));
	idx->d.tab = (const IdxSlot *)idx->d_tab, idx->d.tab_mask = (uint64_t)idx->tab_slots - 1;
	idx->d.pos = (const uint64_t *)idx->d_pos, idx->d.S = (const uint32_t *)idx->d_S;
	idx->d.seq_off = (const uint64_t *)idx->d_seq_off, idx->d.seq_len = (const uint32_t *)idx->d_seq_len;
	return GD_OK;
}

// --------------------------------------------------------------------------------------------
// .mmi -> device index (row F4): the host part parses the file (host/gd_mmi.cpp), the arrays go up and the same table
// insert kernel as in the build makes them searchable
// --------------------------------------------------------------------------------------------
struct GdMmiData {
	int w = 0, k = 0, b = 0, flag = 0;
	std::vector<std::string> names;
	std::vector<int32_t> lens;
	std::vector<uint64_t> keys, positions;
	std::vector<uint32_t> counts, S;
};
int gd_mmi_parse(const char *path, GdMmiData &D);

extern "C" int gd_index_load_mmi(gd_ctx *ctx, const char *path, gd_index **out)
{
	if (!ctx || !path || !out) return GD_ERR_ARG;
	*out = nullptr;
	GdMmiData D;
	if (gd_mmi_parse(path, D) != GD_OK) {
		ctx->err = std::string("gd_index_load_mmi: cannot read ") + path;
		return GD_ERR_ARG;
	}
	gd_index_meta_t m;
	m.n_seq = (int64_t)D.lens.size(), m.total_len = 0, m.n_minimizers = (int64_t)D.positions.size(), m.n_keys = (int64_t)D.keys.size();
	for (int32_t l : D.lens) m.total_len += (uint32_t)l;
	m.table_slots = 1024;
	while (m.table_slots < 2 * m.n_keys) m.table_slots <<= 1;
	m.s_words = (m.total_len + 7) / 8, m.w = D.w, m.k = D.k;
	if (m.n_minimizers > 0xffffffffll) { // the table stores the first position of a minimizer as uint32 (like the build path)
		ctx->err = "gd_index_load_mmi: more than 2^32 minimizers in one index part";
		return GD_ERR_ARG;
	}
	if (m.n_seq <= 0 || (int64_t)D.S.size() != m.s_words) {
		ctx->err = "gd_index_load_mmi: the file has no sequences (index dumped with MM_I_NO_SEQ?)";
		return GD_ERR_ARG;
	}
	gd_index *idx = nullptr;
	int rc = gd_index_alloc(ctx, &m, &idx);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	std::vector<uint64_t> seq_off((size_t)m.n_seq + 1, 0), first((size_t)m.n_keys + 1, 0);
	for (int64_t i = 0; i < m.n_seq; ++i) seq_off[i + 1] = seq_off[i] + (uint32_t)D.lens[i];
	for (int64_t i = 0; i < m.n_keys; ++i) first[i + 1] = first[i] + D.counts[i];
	uint64_t *d_first = nullptr;
	auto up = [&](void *dst, const void *src, size_t bytes) { return bytes == 0 || cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s) == cudaSuccess; };
	bool ok = cudaMalloc(&d_first, (size_t)(m.n_keys + 1) * 8) == cudaSuccess;
	ok = ok && up(idx->d_pos, D.positions.data(), D.positions.size() * 8) && up(idx->d_S, D.S.data(), D.S.size() * 4) &&
	     up(idx->d_seq_off, seq_off.data(), seq_off.size() * 8) && up(idx->d_seq_len, D.lens.data(), D.lens.size() * 4) &&
	     up(idx->d_keys, D.keys.data(), D.keys.size() * 8) && up(idx->d_counts, D.counts.data(), D.counts.size() * 4) &&
	     up(d_first, first.data(), first.size() * 8);
	if (ok) {
		gd_idx_clear_kernel<<<(unsigned)((m.table_slots + 255) / 256), 256, 0, s>>>(m.table_slots, (IdxSlot *)idx->d_tab);
		if (m.n_keys > 0)
			gd_idx_insert_kernel<<<(unsigned)((m.n_keys + 255) / 256), 256, 0, s>>>(m.n_keys, (const uint64_t *)idx->d_keys, (const uint32_t *)idx->d_counts,
			                                                                     d_first, (IdxSlot *)idx->d_tab, (uint64_t)m.table_slots - 1);
		ctx->stat_launches += 2;
		ok = cudaGetLastError() == cudaSuccess && cudaStreamSynchronize(s) == cudaSuccess;
	}
	if (d_first) cudaFree(d_first);
	if (!ok || (rc = gd_index_commit(ctx, idx)) != GD_OK) {
		ctx->err = "gd_index_load_mmi: CUDA failure while uploading the index";
		gd_index_destroy(idx);
		return GD_ERR_CUDA;
	}
	idx->names = D.names;
	*out = idx;
	return GD_OK;
}

extern "C" const char *gd_index_seq_name(const gd_index *idx, int i)
{
	if (!idx || i < 0 || (size_t)i >= idx->names.size()) return "";
	return idx->names[(size_t)i].c_str();
}
