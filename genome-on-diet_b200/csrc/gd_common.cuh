// gd_common.cuh -- device-side helpers shared by the DP and sketch kernels.
//
// The kernels are written against the tiny SIMT vocabulary below (lane/thread ids, shuffles,
// warp/block barriers, 16x2 packed integer ops).  In the product build (nvcc, sm_100a) every
// item maps 1:1 onto a hardware instruction.  When GD_HOST_EMU is defined (ONLY by
// tests/emu/*, never by the product build) the same vocabulary is provided by a cooperative
// fiber scheduler so the kernel *logic* can be unit-tested in a container without a GPU.
#pragma once
#include <stdint.h>

#ifdef GD_HOST_EMU
#include "simt_emu.h" // tests/emu/simt_emu.h (test infrastructure)
#define GD_DEV static inline
#define GD_MEM inline /* static member functions */
#define GD_HD static inline
#define GD_GLOBAL static
#define GD_RESTRICT
#else
#include <cuda_runtime.h>
#define GD_DEV __device__ __forceinline__
#define GD_MEM __device__ __forceinline__
#define GD_HD __host__ __device__ __forceinline__
#define GD_GLOBAL __global__
#define GD_RESTRICT __restrict__
#endif

namespace gd {

#ifndef GD_HOST_EMU
// ---- hardware mapping (sm_100a: VIADD.16x2 / VIMNMX.S16x2 / VIMNMX3.S16x2 / PRMT / SHF / SHFL) ----
GD_DEV uint32_t vadd2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
GD_DEV uint32_t vmax2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
GD_DEV uint32_t vmin2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
GD_DEV uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
GD_DEV uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); } // max(a+b, c) per half
GD_DEV int iaddmax(int a, int b, int c) { return __viaddmax_s32(a, b, c); }                          // max(a+b, c)
GD_DEV int imax3(int a, int b, int c) { return __vimax3_s32(a, b, c); }
GD_DEV uint32_t fma_add(uint32_t a, uint32_t one, uint32_t c)
{ // a * one + c with `one` == 1 at run time: an IMAD (FMA pipe) where a plain add would be an ALU-pipe VIADD
	uint32_t r;
	asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(one), "r"(c));
	return r;
}
GD_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{ // raw PRMT: unlike __byte_perm (which masks the selector with 0x7777) this keeps bit 3 of every
  // selector nibble = "replicate the sign bit of the selected byte"
	uint32_t r;
	asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
	return r;
}
GD_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh) { return __funnelshift_r(lo, hi, sh); }
// acc + (signed byte `byte` of v): one IDP.4A (dot product with a one-hot byte vector) instead of a sign-extending PRMT and an add
GD_DEV uint32_t add_sbyte(uint32_t acc, uint32_t v, int byte) { return (uint32_t)__dp4a((int)v, (int)(1u << (8 * byte)), (int)acc); }
GD_DEV uint32_t shfl_idx(uint32_t mask, uint32_t v, int src, int width) { return __shfl_sync(mask, v, src, width); }
GD_DEV uint32_t shfl_xor(uint32_t mask, uint32_t v, int lm, int width) { return __shfl_xor_sync(mask, v, lm, width); }
GD_DEV uint32_t shfl_up(uint32_t mask, uint32_t v, int d, int width) { return __shfl_up_sync(mask, v, d, width); }
GD_DEV uint32_t ballot(uint32_t mask, int pred) { return __ballot_sync(mask, pred); }
GD_DEV int reduce_max(uint32_t mask, int v) { return __reduce_max_sync(mask, v); } // REDUX.MAX
GD_DEV void sync_warp(uint32_t mask) { __syncwarp(mask); }
GD_DEV void sync_block() { __syncthreads(); }
GD_DEV int thread_idx() { return (int)threadIdx.x; }
GD_DEV int block_idx() { return (int)blockIdx.x; }
GD_DEV int block_dim() { return (int)blockDim.x; }
GD_DEV int grid_dim() { return (int)gridDim.x; }
GD_DEV int atomic_add(int *p, int v) { return atomicAdd(p, v); }
GD_DEV unsigned long long atomic_add64(unsigned long long *p, unsigned long long v) { return atomicAdd(p, v); }
GD_DEV int popc(uint32_t x) { return __popc(x); }
GD_DEV int clz32(uint32_t x) { return __clz((int)x); }
GD_DEV int ffs32(uint32_t x) { return __ffs((int)x); } // 1-based position of the lowest set bit, 0 if none
GD_DEV uint64_t brev64(uint64_t x) { return __brevll(x); }
GD_DEV void fence() { __threadfence(); }
template <class T> GD_DEV T ld_volatile(const T *p) { return *(const volatile T *)p; }
template <class T> GD_DEV void st_volatile(T *p, T v) { *(volatile T *)p = v; }
#else
// ---- emulation (tests/emu/simt_emu.h supplies the scheduler-backed collectives) ----
GD_DEV uint32_t vadd2(uint32_t a, uint32_t b)
{
	return ((a + b) & 0xffffu) | ((((a >> 16) + (b >> 16)) & 0xffffu) << 16);
}
GD_DEV uint32_t vmax2(uint32_t a, uint32_t b)
{
	int16_t al = (int16_t)a, bl = (int16_t)b, ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
	return (uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
}
GD_DEV uint32_t vmin2(uint32_t a, uint32_t b)
{
	int16_t al = (int16_t)a, bl = (int16_t)b, ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
	return (uint16_t)(al < bl ? al : bl) | ((uint32_t)(uint16_t)(ah < bh ? ah : bh) << 16);
}
GD_DEV uint32_t vmax3(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vmax2(a, b), c); }
GD_DEV uint32_t vaddmax2(uint32_t a, uint32_t b, uint32_t c) { return vmax2(vadd2(a, b), c); }
GD_DEV int iaddmax(int a, int b, int c)
{
	int s = (int)((uint32_t)a + (uint32_t)b);
	return s > c ? s : c;
}
GD_DEV int imax3(int a, int b, int c)
{
	int m = a > b ? a : b;
	return m > c ? m : c;
}
GD_DEV uint32_t fma_add(uint32_t a, uint32_t one, uint32_t c) { return a * one + c; }
GD_DEV uint32_t prmt(uint32_t a, uint32_t b, uint32_t s)
{
	uint64_t v = ((uint64_t)b << 32) | a;
	uint32_t r = 0;
	for (int i = 0; i < 4; ++i) {
		uint32_t sel = (s >> (4 * i)) & 0xf, byte = (uint32_t)(v >> (8 * (sel & 7))) & 0xff;
		if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;
		r |= byte << (8 * i);
	}
	return r;
}
GD_DEV uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
	uint64_t v = ((uint64_t)hi << 32) | lo;
	return (uint32_t)(v >> (sh & 31));
}
GD_DEV uint32_t add_sbyte(uint32_t acc, uint32_t v, int byte) { return acc + (uint32_t)(int32_t)(int8_t)(v >> (8 * byte)); }
GD_DEV uint32_t shfl_idx(uint32_t mask, uint32_t v, int src, int width) { return emu::shfl_idx(mask, v, src, width); }
GD_DEV uint32_t shfl_xor(uint32_t mask, uint32_t v, int lm, int width) { return emu::shfl_xor(mask, v, lm, width); }
GD_DEV uint32_t shfl_up(uint32_t mask, uint32_t v, int d, int width) { return emu::shfl_up(mask, v, d, width); }
GD_DEV uint32_t ballot(uint32_t mask, int pred) { return emu::ballot(mask, pred); }
GD_DEV int reduce_max(uint32_t mask, int v) { return emu::reduce_max(mask, v); }
GD_DEV void sync_warp(uint32_t mask) { emu::sync_warp(mask); }
GD_DEV void sync_block() { emu::sync_block(); }
GD_DEV int thread_idx() { return emu::thread_idx(); }
GD_DEV int block_idx() { return emu::block_idx(); }
GD_DEV int block_dim() { return emu::block_dim(); }
GD_DEV int grid_dim() { return emu::grid_dim(); }
GD_DEV int atomic_add(int *p, int v)
{
	emu::yield(); // a scheduling point: other threads / blocks may draw first
	int o = *p;
	*p = o + v;
	return o;
}
GD_DEV unsigned long long atomic_add64(unsigned long long *p, unsigned long long v)
{
	unsigned long long o = *p;
	*p = o + v;
	return o;
}
GD_DEV int popc(uint32_t x) { return __builtin_popcount(x); }
GD_DEV int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
GD_DEV int ffs32(uint32_t x) { return __builtin_ffs((int)x); }
GD_DEV uint64_t brev64(uint64_t x)
{
	uint64_t r = 0;
	for (int i = 0; i < 64; ++i) r |= ((x >> i) & 1ull) << (63 - i);
	return r;
}
GD_DEV void fence() {}
template <class T> GD_DEV T ld_volatile(const T *p)
{
	emu::yield(); // let the producer fiber run
	return *p;
}
template <class T> GD_DEV void st_volatile(T *p, T v) { *p = v; }
#endif

GD_DEV int imin(int a, int b) { return a < b ? a : b; }
GD_DEV int imax(int a, int b) { return a > b ? a : b; }

} // namespace gd
