// tools/ksw_kbench.cu -- developer microbenchmark of the DP kernel alone (not part of the product
// library): builds the device code of csrc/gd_ksw.cuh with compile-time variants (-DGD_KSW_*) and
// times pack + DP on synthetic 150x200 pairs, printing GCUPS and a checksum of all result records so
// that variants can be compared for speed and equality in one GPU session.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include "../genome-on-diet_b200/csrc/gd_ksw_host.h"
using namespace gd;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int G, bool RIGHT, int MODE, bool WITH_P>
__global__ void __launch_bounds__(128) dp_kernel(const KswConsts C, const KswBatch B)
{
	extern __shared__ __align__(128) uint8_t gd_smem[];
	const int tid = threadIdx.x;
	ksw_build_lut(gd_smem, tid, blockDim.x);
	__syncthreads();
	uint8_t *warp_smem = gd_smem + GD_KSW_LUT_BYTES + (size_t)(tid >> 5) * (32 / G) * B.group_smem;
	ksw_warp_body<G, RIGHT, MODE, WITH_P>(C, B, warp_smem, gd_smem, tid & 31);
}
__global__ void pack_kernel(int n, const uint8_t *q, int qlen, const uint8_t *t, int tlen, uint8_t *tpk, int t_stride, uint8_t *qpk,
                            int q_stride, const KswConsts C, KswHot *hot)
{
	if (blockIdx.x == 0 && threadIdx.x == 0) *hot = ksw_hot_from_consts(C);
	const int warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps)
		ksw_pack_pair(q + (size_t)i * qlen, qlen, t + (size_t)i * tlen, tlen, tpk + (size_t)i * t_stride, t_stride,
		              qpk + (size_t)i * q_stride, q_stride, lane, 32);
}

typedef void (*kern_t)(const KswConsts, const KswBatch);
template <int G> kern_t pick(int mode)
{
	return mode == 2 ? dp_kernel<G, false, 2, true> : mode == 1 ? dp_kernel<G, false, 1, true> : dp_kernel<G, false, 0, true>;
}

int main(int argc, char **argv)
{
	int n = argc > 1 ? atoi(argv[1]) : 262144, reps = argc > 2 ? atoi(argv[2]) : 5;
	const int qlen = 150, tlen = 200, w = 150;
	std::vector<uint8_t> q((size_t)n * qlen), t((size_t)n * tlen);
	uint64_t s = 88172645463325252ull;
	auto rnd = [&]() { s ^= s << 13, s ^= s >> 7, s ^= s << 17; return (uint32_t)(s >> 11); };
	for (int i = 0; i < n; ++i) {
		uint8_t *tt = &t[(size_t)i * tlen], *qq = &q[(size_t)i * qlen];
		for (int k = 0; k < tlen; ++k) tt[k] = rnd() & 3;
		int src = 0;
		for (int k = 0; k < qlen; ++k) {
			uint32_t x = rnd() % 1000;
			if (x < 30) qq[k] = (tt[src] + 1 + rnd() % 3) & 3, ++src; // substitution
			else if (x < 40) qq[k] = rnd() & 3;                          // insertion
			else if (x < 50) src += 1, qq[k] = tt[src < tlen ? src : tlen - 1], ++src; // deletion
			else qq[k] = tt[src < tlen ? src : tlen - 1], ++src;
			if (src >= tlen) src = tlen - 1;
		}
		if (i % 50 == 49) qq[rnd() % qlen] = 4;
	}
	int8_t mat[25];
	for (int a = 0; a < 5; ++a) for (int b = 0; b < 5; ++b) mat[a * 5 + b] = a == 4 || b == 4 ? 0 : a == b ? 2 : -8;
	cudaDeviceProp prop;
	CK(cudaGetDeviceProperties(&prop, 0));
	uint8_t *dq, *dt, *tpk, *qpk, *p;
	int32_t *dql, *dtl, *ticket;
	KswResult *res;
	CK(cudaMalloc(&dq, q.size())); CK(cudaMalloc(&dt, t.size()));
	CK(cudaMemcpy(dq, q.data(), q.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dt, t.data(), t.size(), cudaMemcpyHostToDevice));
	std::vector<int32_t> ql(n, qlen), tl(n, tlen);
	CK(cudaMalloc(&dql, n * 4)); CK(cudaMalloc(&dtl, n * 4)); CK(cudaMalloc(&ticket, 256)); CK(cudaMalloc(&res, (size_t)n * sizeof(KswResult)));
	CK(cudaMemcpy(dql, ql.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dtl, tl.data(), n * 4, cudaMemcpyHostToDevice));
	long cells = 0;
	for (int r = 0; r < qlen + tlen - 1; ++r) {
		int st0 = std::max(std::max(0, r - qlen + 1), (r - w + 1) >> 1), en0 = std::min(std::min(tlen - 1, r), (r + w) >> 1);
		cells += en0 - st0 + 1;
	}
	printf("variant PREFETCH=%d HOTMEM=%d ST16=%d NBSHFL=%d P32=%d  n=%d cells/pair=%ld\n", GD_KSW_PREFETCH, GD_KSW_HOTMEM, GD_KSW_ST16, GD_KSW_NBSHFL,
	       GD_KSW_P32, n, cells);
	for (int flag : {0x08, 0x00}) {
		const bool exact = !(flag & 8);
		KswConsts C = ksw_make_consts(5, mat, 12, 2, 24, 1, 100, 10, flag);
		KswGeom geo = ksw_geometry(qlen, tlen, w, exact, true, 4);
		CK(cudaMalloc(&tpk, (size_t)n * geo.t_stride + 64)); CK(cudaMalloc(&qpk, (size_t)n * geo.q_stride + 64));
		CK(cudaMalloc(&p, (size_t)n * geo.p_stride + 64));
		for (int G : {4, 8}) {
			geo = ksw_geometry(qlen, tlen, w, exact, true, G);
			kern_t kern = G == 4 ? pick<4>(exact ? 2 : 0) : pick<8>(exact ? 2 : 0);
			const size_t per_warp = (size_t)(32 / G) * geo.group_smem;
			int threads = 0, best = 0;
			for (int wpb = 1; wpb <= 4; ++wpb) {
				size_t blk = GD_KSW_LUT_BYTES + wpb * per_warp;
				if (blk > prop.sharedMemPerBlockOptin) break;
				int resident = (int)std::min<size_t>(32, prop.sharedMemPerMultiprocessor / (blk + 1024)) * wpb;
				if (resident >= best) best = resident, threads = wpb * 32;
			}
			const size_t smem = GD_KSW_LUT_BYTES + (size_t)(threads / G) * geo.group_smem;
			CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			int occ = 0;
			CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
			KswBatch B;
			B.n = n, B.base = 0, B.qlen = dql, B.tlen = dtl, B.w = 0, B.w_all = w, B.tpk = tpk, B.qpk = qpk, B.t_stride = geo.t_stride;
			B.q_stride = geo.q_stride, B.p = p, B.p_stride = geo.p_stride, B.res = res, B.ticket = ticket, B.ring = geo.ring;
			B.group_smem = geo.group_smem, B.hot = (const KswHot *)((uint8_t *)ticket + 64), B.lead64 = nullptr;
			cudaEvent_t e0, e1;
			CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
			float best_ms = 1e9;
			for (int it = 0; it < reps + 2; ++it) {
				CK(cudaMemset(ticket, 0, 4));
				pack_kernel<<<prop.multiProcessorCount * 16, 128>>>(n, dq, qlen, dt, tlen, tpk, geo.t_stride, qpk, geo.q_stride, C,
				                                                     (KswHot *)((uint8_t *)ticket + 64));
				CK(cudaEventRecord(e0));
				kern<<<prop.multiProcessorCount * occ, threads, smem>>>(C, B);
				CK(cudaEventRecord(e1));
				CK(cudaEventSynchronize(e1));
				float ms;
				CK(cudaEventElapsedTime(&ms, e0, e1));
				if (it >= 2) best_ms = std::min(best_ms, ms);
			}
			CK(cudaGetLastError());
			std::vector<KswResult> h(n);
			CK(cudaMemcpy(h.data(), res, (size_t)n * sizeof(KswResult), cudaMemcpyDeviceToHost));
			uint64_t ck = 0;
			for (int i = 0; i < n; ++i) {
				const int32_t *x = (const int32_t *)&h[i];
				for (int k = 0; k < 14; ++k) ck = ck * 1000003ull + (uint32_t)x[k];
			}
			printf("flag %#04x G %d threads %3d blocks/SM %d warps/SM %2d  %.3f ms  %.1f GCUPS  checksum %016llx\n", flag, G, threads, occ,
			       occ * threads / 32, best_ms, (double)n * cells / best_ms / 1e6, (unsigned long long)ck);
		}
		cudaFree(tpk), cudaFree(qpk), cudaFree(p);
	}
	return 0;
}
