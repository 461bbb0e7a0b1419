// gd_ubench.cu -- integer-pipe microbenchmark for the DP roofline ("take L from an
// IADD3/VIMNMX microbenchmark on the box; MEASURED_PEAKS.json has no integer entry").
// Prints one JSON line per instruction class: lane-ops per clock per SM and chip-wide Gop/s.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define ITERS 4096
#define CH 8

template <int OP> __device__ __forceinline__ uint32_t op(uint32_t a, uint32_t b, uint32_t c)
{ // asm volatile: the compiler must issue every instruction of the dependent chains
	uint32_t r;
	if (OP == 0) asm volatile("add.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (OP == 1) asm volatile("add.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (OP == 2) asm volatile("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (OP == 3) asm volatile("{.reg .b32 t; max.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 4) asm volatile("{.reg .b32 t; add.s16x2 t, %1, %2; max.s16x2 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 5) asm volatile("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 6) asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c & 0x7777));
	else if (OP == 7) asm volatile("max.s32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
	else if (OP == 8) asm volatile("mad.lo.s32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 9) asm volatile("shf.r.wrap.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 10) asm volatile("shfl.sync.bfly.b32 %0, %1, 1, 0x1f, 0xffffffff;" : "=r"(r) : "r"(a));
	else if (OP == 11) asm volatile("{.reg .s32 t; max.s32 t, %1, %2; max.s32 %0, t, %3;}" : "=r"(r) : "r"(a), "r"(b), "r"(c));
	else if (OP == 12) asm volatile("shr.u32 %0, %1, 3;" : "=r"(r) : "r"(a));
	else r = a;
	return r;
}

template <int OP> __global__ void __launch_bounds__(256) k(uint32_t *out, long long *cycles, uint32_t seed)
{
	uint32_t v[CH], b = seed + threadIdx.x, c = seed * 3 + 1;
#pragma unroll
	for (int i = 0; i < CH; ++i) v[i] = seed + i * 77 + threadIdx.x;
	long long t0 = clock64();
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll
		for (int i = 0; i < CH; ++i) v[i] = op<OP>(v[i], b, c);
	}
	long long t1 = clock64();
	uint32_t s = 0;
#pragma unroll
	for (int i = 0; i < CH; ++i) s ^= v[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP> static void run(const char *name, int sms, int blocks_per_sm)
{
	int blocks = sms * blocks_per_sm;
	uint32_t *out;
	long long *cyc;
	cudaMalloc(&out, (size_t)blocks * 256 * 4);
	cudaMalloc(&cyc, (size_t)blocks * 8);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	k<OP><<<blocks, 256>>>(out, cyc, 1);
	cudaDeviceSynchronize();
	cudaEventRecord(e0);
	k<OP><<<blocks, 256>>>(out, cyc, 2);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	long long *h = (long long *)malloc((size_t)blocks * 8);
	cudaMemcpy(h, cyc, (size_t)blocks * 8, cudaMemcpyDeviceToHost);
	double mean = 0;
	for (int i = 0; i < blocks; ++i) mean += (double)h[i];
	mean /= blocks;
	double ops_per_sm = (double)blocks_per_sm * 256 * ITERS * CH;
	printf("{\"op\": \"%s\", \"lane_ops_per_clk_per_sm\": %.1f, \"chip_gops\": %.1f, \"ms\": %.3f, \"blocks_per_sm\": %d}\n", name,
	       ops_per_sm / mean, (double)blocks * 256 * ITERS * CH / (ms * 1e6), ms, blocks_per_sm);
	free(h);
	cudaFree(out), cudaFree(cyc);
}

int main()
{
	cudaDeviceProp p;
	if (cudaGetDeviceProperties(&p, 0) != cudaSuccess) {
		fprintf(stderr, "no device\n");
		return 1;
	}
	int sms = p.multiProcessorCount;
	printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
	for (int bps = 4; bps <= 8; bps += 4) {
		run<0>("IADD", sms, bps);
		run<1>("VIADD.16x2", sms, bps);
		run<2>("VIMNMX.S16x2", sms, bps);
		run<3>("VIMNMX3.S16x2", sms, bps);
		run<4>("VIADDMNMX.S16x2", sms, bps);
		run<5>("LOP3", sms, bps);
		run<6>("PRMT", sms, bps);
		run<7>("VIMNMX.S32", sms, bps);
		run<8>("IMAD", sms, bps);
		run<9>("SHF", sms, bps);
		run<10>("SHFL", sms, bps);
		run<11>("VIMNMX3.S32", sms, bps);
		run<12>("SHR.imm", sms, bps);
	}
	return 0;
}
