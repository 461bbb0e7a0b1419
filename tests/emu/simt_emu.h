// tests/emu/simt_emu.h -- TEST INFRASTRUCTURE ONLY.
// A minimal cooperative-fiber SIMT emulator: every CUDA thread of a block is a ucontext fiber;
// warp shuffles / barriers rendez-vous through per-thread epochs.  It exists so that the
// *logic* of the hand-written kernels (lane ownership, ring buffers, carries, reductions) can be
// checked against the oracle inside a container that has no GPU.  It is never compiled into the
// product library and is not a CPU fallback: nothing outside tests/emu includes it.
#pragma once
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <ucontext.h>
#include <functional>
#include <vector>

// CUDA vector types used by the kernels
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };

namespace emu {

struct Fiber {
	ucontext_t ctx;
	char *stack;
	bool done;
	uint32_t wepoch; // warp-collective epoch
	uint32_t bepoch; // block-barrier epoch
	uint32_t xchg;
};

struct BlockState {
	int bid, nthreads, grid;
	std::vector<Fiber> th;
	ucontext_t sched;
	int cur;
	char *smem;
	std::function<void()> body;
};

inline BlockState *&cur_block()
{
	static BlockState *b = 0;
	return b;
}

inline unsigned &shuffle_seed()
{ // != 0: launch() keeps all blocks resident and schedules blocks and threads in pseudo-random orders (launch_concurrent)
	static unsigned s = 0;
	return s;
}

inline long &rowmax_mismatches()
{ // self-check counter of the DP kernel's exact-max fast path (see gd_ksw.cuh)
	static long n = 0;
	return n;
}

inline void yield()
{
	BlockState *b = cur_block();
	if (!b) return; // device code run as a plain host loop (no fibers): nothing to switch to
	swapcontext(&b->th[b->cur].ctx, &b->sched);
}

inline int thread_idx() { return cur_block()->cur; }
inline int block_idx() { return cur_block()->bid; }
inline int block_dim() { return cur_block()->nthreads; }
inline int grid_dim() { return cur_block()->grid; }
inline char *smem() { return cur_block()->smem; }

inline void sync_warp(uint32_t mask)
{
	BlockState *b = cur_block();
	int tid = b->cur, base = tid & ~31;
	uint32_t e = ++b->th[tid].wepoch;
	for (;;) {
		bool ok = true;
		for (int l = 0; l < 32 && ok; ++l)
			if ((mask >> l & 1) && base + l < b->nthreads && b->th[base + l].wepoch < e) ok = false;
		if (ok) return;
		yield();
	}
}

inline uint32_t xchg(uint32_t mask, uint32_t v, int src_lane)
{
	BlockState *b = cur_block();
	int tid = b->cur, base = tid & ~31;
	b->th[tid].xchg = v;
	sync_warp(mask);
	uint32_t r = b->th[base + src_lane].xchg;
	sync_warp(mask);
	return r;
}

inline uint32_t shfl_idx(uint32_t mask, uint32_t v, int src, int width)
{
	int lane = cur_block()->cur & 31;
	return xchg(mask, v, (lane & ~(width - 1)) | (src & (width - 1)));
}
inline uint32_t shfl_xor(uint32_t mask, uint32_t v, int lm, int width)
{
	int lane = cur_block()->cur & 31, s = lane ^ lm;
	if ((s & ~(width - 1)) != (lane & ~(width - 1))) s = lane;
	return xchg(mask, v, s);
}
inline uint32_t shfl_up(uint32_t mask, uint32_t v, int d, int width)
{
	int lane = cur_block()->cur & 31, s = lane - d;
	if (s < (lane & ~(width - 1))) s = lane;
	return xchg(mask, v, s);
}
inline uint32_t ballot(uint32_t mask, int pred)
{
	BlockState *b = cur_block();
	int tid = b->cur, base = tid & ~31;
	b->th[tid].xchg = pred ? 1u : 0u;
	sync_warp(mask);
	uint32_t r = 0;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l & 1) && base + l < b->nthreads && (b->th[base + l].xchg & 1u)) r |= 1u << l;
	sync_warp(mask);
	return r;
}

inline int reduce_max(uint32_t mask, int v)
{ // __reduce_max_sync: maximum over the lanes named in this thread's mask
	BlockState *b = cur_block();
	int tid = b->cur, base = tid & ~31;
	b->th[tid].xchg = (uint32_t)v;
	sync_warp(mask);
	int r = v;
	for (int l = 0; l < 32; ++l)
		if ((mask >> l & 1) && base + l < b->nthreads && (int)b->th[base + l].xchg > r) r = (int)b->th[base + l].xchg;
	sync_warp(mask);
	return r;
}

inline void sync_block()
{
	BlockState *b = cur_block();
	int tid = b->cur;
	uint32_t e = ++b->th[tid].bepoch;
	for (;;) {
		bool ok = true;
		for (int t = 0; t < b->nthreads && ok; ++t)
			if (b->th[t].bepoch < e) ok = false;
		if (ok) return;
		yield();
	}
}

inline void fiber_entry()
{
	BlockState *b = cur_block();
	b->body();
	b->th[b->cur].done = true;
	swapcontext(&b->th[b->cur].ctx, &b->sched);
}

// Run `body` once per thread for every block of the grid (blocks run one after another, in order).
inline void launch_concurrent(int grid, int block, size_t smem_bytes, std::function<void()> body, unsigned seed);
inline void launch(int grid, int block, size_t smem_bytes, std::function<void()> body)
{
	if (shuffle_seed()) {
		launch_concurrent(grid, block, smem_bytes, body, shuffle_seed());
		return;
	}
	const size_t STK = 256 * 1024;
	for (int bid = 0; bid < grid; ++bid) {
		BlockState b;
		b.bid = bid, b.nthreads = block, b.grid = grid, b.body = body;
		{ // shared memory is uninitialised on the device: poison it so that stale reads show up
			size_t nb = (smem_bytes + 255) / 128 * 128;
			b.smem = (char *)aligned_alloc(128, nb);
			memset(b.smem, 0xCD, nb);
		}
		b.th.resize(block);
		cur_block() = &b;
		for (int t = 0; t < block; ++t) {
			Fiber &f = b.th[t];
			f.done = false, f.wepoch = f.bepoch = 0, f.xchg = 0;
			f.stack = (char *)malloc(STK);
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = f.stack, f.ctx.uc_stack.ss_size = STK, f.ctx.uc_link = &b.sched;
			makecontext(&f.ctx, (void (*)())fiber_entry, 0);
		}
		for (;;) {
			bool any = false;
			for (int t = 0; t < block; ++t) {
				if (b.th[t].done) continue;
				any = true;
				b.cur = t;
				swapcontext(&b.sched, &b.th[t].ctx);
			}
			if (!any) break;
		}
		for (int t = 0; t < block; ++t) free(b.th[t].stack);
		free(b.smem);
		cur_block() = 0;
	}
}

// The same, with all blocks of the grid resident at once: the scheduler visits the blocks in a pseudo-random order and lets
// each one run one to three sweeps (every live thread once -- in a rotated, possibly reversed order -- up to its next barrier /
// collective / volatile load) per visit.
// For protocols BETWEEN blocks -- ordering tickets, decoupled look-back, deferred copy-out -- under many interleavings.
inline void launch_concurrent(int grid, int block, size_t smem_bytes, std::function<void()> body, unsigned seed)
{
	const size_t STK = 256 * 1024;
	std::vector<BlockState> bs(grid);
	for (int bid = 0; bid < grid; ++bid) {
		BlockState &b = bs[bid];
		b.bid = bid, b.nthreads = block, b.grid = grid, b.body = body, b.cur = 0;
		size_t nb = (smem_bytes + 255) / 128 * 128;
		b.smem = (char *)aligned_alloc(128, nb);
		memset(b.smem, 0xCD, nb);
		b.th.resize(block);
		cur_block() = &b;
		for (int t = 0; t < block; ++t) {
			Fiber &f = b.th[t];
			f.done = false, f.wepoch = f.bepoch = 0, f.xchg = 0;
			f.stack = (char *)malloc(STK);
			getcontext(&f.ctx);
			f.ctx.uc_stack.ss_sp = f.stack, f.ctx.uc_stack.ss_size = STK, f.ctx.uc_link = &b.sched;
			makecontext(&f.ctx, (void (*)())fiber_entry, 0);
		}
	}
	unsigned long long rng = 0x9e3779b97f4a7c15ull ^ seed;
	auto next = [&]() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(rng >> 33); };
	std::vector<int> order(grid);
	for (int i = 0; i < grid; ++i) order[i] = i;
	for (;;) {
		bool any = false;
		for (int i = grid - 1; i > 0; --i) { // shuffle
			int j = (int)(next() % (unsigned)(i + 1)), t = order[i];
			order[i] = order[j], order[j] = t;
		}
		for (int oi = 0; oi < grid; ++oi) {
			BlockState &b = bs[order[oi]];
			cur_block() = &b;
			for (int sweeps = 1 + (int)(next() % 3u); sweeps > 0; --sweeps) {
				// the threads of the block in a rotated, possibly reversed order: between two barriers a thread may run before OR
				// after any other one, so a missing barrier (a read of what another thread has yet to write, or has already
				// overwritten) gives wrong output for some seed instead of passing by the accident of index order
				const int start = (int)(next() % (unsigned)block), dir = next() & 1u ? 1 : block - 1;
				for (int i = 0, t = start; i < block; ++i, t = (t + dir) % block) {
					if (b.th[t].done) continue;
					any = true;
					b.cur = t;
					swapcontext(&b.sched, &b.th[t].ctx);
				}
			}
		}
		if (!any) break;
	}
	for (int bid = 0; bid < grid; ++bid) {
		for (int t = 0; t < block; ++t) free(bs[bid].th[t].stack);
		free(bs[bid].smem);
	}
	cur_block() = 0;
}

} // namespace emu
