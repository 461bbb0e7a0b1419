// tests/emu/emu_sketch.cpp -- TEST INFRASTRUCTURE ONLY: runs the device code of gd_sketch.cuh under
// the fiber SIMT emulator (see simt_emu.h).
#define GD_HOST_EMU 1
#include "simt_emu.h"
#include "gd_sketch.cuh"
#include <vector>

using namespace gd;

static int g_ver = 3; // tile body: 3 = sketch_tile_body3 (the product default), 2 = sketch_tile_body
extern "C" void emu_sketch_version(int v) { g_ver = v; }

static unsigned g_conc = 0; // != 0: all blocks resident, scheduled in a pseudo-random interleaving seeded with this value
extern "C" void emu_sketch_concurrency(unsigned seed) { g_conc = seed; }

static void launch_tiles(int grid, int threads, size_t smem, std::function<void()> body)
{
	if (g_conc) emu::launch_concurrent(grid, threads, smem, body, g_conc);
	else emu::launch(grid, threads, smem, body);
}

template <int THREADS>
static void run_tiles(const SketchParams &S, SketchBatch &B, int grid)
{
	if (g_ver == 2)
		launch_tiles(grid, THREADS, sizeof(SketchSmem<THREADS>),
		            [&]() { sketch_tile_body<THREADS>(S, B, (SketchSmem<THREADS> *)emu::smem()); });
	else if (B.pack_jobs > 0)
		launch_tiles(grid, THREADS, sizeof(SketchSmem3<THREADS>),
		            [&]() { sketch_tile_body3<THREADS, true>(S, B, (SketchSmem3<THREADS> *)emu::smem()); });
	else
		launch_tiles(grid, THREADS, sizeof(SketchSmem3<THREADS>),
		            [&]() { sketch_tile_body3<THREADS, false>(S, B, (SketchSmem3<THREADS> *)emu::smem()); });
}

// jobs: n x {seq_off, len, shift, rid}; small != 0 forces the one-tile-per-job configuration
extern "C" long emu_sketch_jobs(int njobs, const int64_t *seq_off, const int32_t *len, const int32_t *shift,
                                const uint32_t *rid, const char *buf, int w, int k, const char *Z, int W, int small,
                                int grid, int64_t *out_off, uint64_t *out, int64_t out_cap)
{
	SketchParams S;
	memset(&S, 0, sizeof(S));
	S.w = w, S.k = k, S.W = W, S.mask = (1ull << 2 * k) - 1;
	for (int g = 0; g < W; ++g)
		if (Z[g] == '1') S.ones_loc[S.ones++] = (uint8_t)g;
	std::vector<SketchJob> jobs(njobs);
	S.TP = sk_tile_emit(small ? 256 : 2048, w, k);
	S.one_tile_per_job = small;
	std::vector<int64_t> tb(njobs + 1, 0);
	for (int i = 0; i < njobs; ++i) {
		// host copy of sk_diet_len
		int64_t dl = 0;
		if (len[i] >= shift[i]) {
			uint32_t rem = (uint32_t)(len[i] - shift[i]) % (uint32_t)W;
			dl = (int64_t)((uint32_t)(len[i] - shift[i]) / (uint32_t)W) * S.ones;
			for (int o = 0; o < S.ones; ++o)
				if (S.ones_loc[o] < rem) ++dl;
		}
		jobs[i] = SketchJob{seq_off[i], len[i], shift[i], rid[i], (int32_t)dl};
		int64_t t = small ? 1 : (dl + S.TP - 1) / S.TP;
		tb[i + 1] = tb[i] + (t < 1 ? 1 : t);
	}
	std::vector<unsigned long long> status(tb[njobs] + 1, 0);
	int ticket = 0;
	SketchBatch B;
	memset(&B, 0, sizeof(B));
	B.njobs = njobs, B.ntiles = tb[njobs], B.jobs = jobs.data(), B.tile_base = small ? nullptr : tb.data();
	B.buf = buf, B.status = status.data(), B.ticket = &ticket, B.out_off = out_off, B.out = out, B.out_cap = out_cap;
	if (small) run_tiles<32>(S, B, grid);
	else run_tiles<256>(S, B, grid);
	return (long)out_off[njobs];
}

// fixed-stride output with `pack` whole jobs per tile of `threads` threads (the short-read configuration of gd_sketch_run_jobs);
// job j's records land at out[j * stride ..], its count in out_cnt[j].  pack = 0: one job per one-warp tile.
extern "C" int emu_sketch_packed(int njobs, const int64_t *seq_off, const int32_t *len, const int32_t *shift, const uint32_t *rid,
                                 const char *buf, int w, int k, const char *Z, int W, int pack, int threads, int grid,
                                 int64_t stride, int32_t *out_cnt, uint64_t *out)
{
	SketchParams S;
	memset(&S, 0, sizeof(S));
	S.w = w, S.k = k, S.W = W, S.mask = (1ull << 2 * k) - 1;
	for (int g = 0; g < W; ++g)
		if (Z[g] == '1') S.ones_loc[S.ones++] = (uint8_t)g;
	std::vector<SketchJob> jobs(njobs);
	for (int i = 0; i < njobs; ++i) {
		int64_t dl = 0;
		if (len[i] >= shift[i]) {
			uint32_t rem = (uint32_t)(len[i] - shift[i]) % (uint32_t)W;
			dl = (int64_t)((uint32_t)(len[i] - shift[i]) / (uint32_t)W) * S.ones;
			for (int o = 0; o < S.ones; ++o)
				if (S.ones_loc[o] < rem) ++dl;
		}
		jobs[i] = SketchJob{seq_off[i], len[i], shift[i], rid[i], (int32_t)dl};
	}
	S.TP = sk_tile_emit(256, w, k), S.one_tile_per_job = 1;
	std::vector<int64_t> oo(njobs + 1, 0);
	std::vector<unsigned long long> status(njobs + 1, 0);
	int ticket = 0;
	SketchBatch B;
	memset(&B, 0, sizeof(B));
	B.njobs = njobs, B.ntiles = pack > 0 ? (njobs + pack - 1) / pack : njobs, B.jobs = jobs.data(), B.buf = buf;
	B.status = status.data(), B.ticket = &ticket, B.out_off = oo.data(), B.out = out, B.out_cap = (int64_t)njobs * stride;
	B.fixed_stride = stride, B.out_cnt = out_cnt, B.pack_jobs = pack;
	if (threads == 32) run_tiles<32>(S, B, grid);
	else if (threads == 64) run_tiles<64>(S, B, grid);
	else if (threads == 128) run_tiles<128>(S, B, grid);
	else return -1;
	for (int i = 0; i <= njobs; ++i)
		if (oo[i] != (int64_t)i * stride) return -2;
	return 0;
}
