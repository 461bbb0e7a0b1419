// gd_sketch.cuh -- sparsified (w,k)-minimizer sketching for sm_100a.
//
// Replaces mm_sketch / mm_sketch2 / mm_sketch3 (GDiet-ShortReads/sketch.c:156,618,1078,1577,1769,
// 1908,2143) with the position-parallel formulation of SURVEY.md 8 A1:
//
//   walk the sparsified sequence  i -> real(i) = (i/ones)*W + ones_loc[i%ones] + shift
//   X(i) = hash64(min(fw,rv))<<8 | k   if the k-mer ending at i has no N and fw != rv, else MAX
//   a window ending at e is full iff the N-free run ending at e is >= w+k-1 long
//   emit i  <=>  X(i) != MAX and X(i) == min X over some full window [e-w+1,e] containing i
//   output ascending in i:  x = X(i),  y = rid<<32 | real(i)<<1 | strand
//
// A *job* is one (sequence, shift, cropped length) triple.  A *tile* is TP consecutive
// sparsified positions of one job, handled by one thread block: every thread rolls the two
// 2-bit k-mers over P consecutive positions (after a k-1 base warm-up), hashes, and leaves X in
// shared memory; the block then takes the sliding minimum over full windows, decides emission
// by a sliding maximum of those minima, and appends its records to the output with a
// decoupled look-back scan over tiles so that the global output is ordered exactly like the
// reference's (jobs in input order, positions ascending).
#pragma once
#include "gd_common.cuh"

namespace gd {

struct SketchJob {
	int64_t seq_off;  // byte offset of the sequence in the ASCII buffer
	int32_t len;      // bytes visible to this job (len or len_crop)
	int32_t shift;
	uint32_t rid;
	int32_t pad;
};

struct SketchParams {
	int32_t w, k, W, ones;
	uint8_t ones_loc[64];
	uint64_t mask;   // 2^(2k)-1
	int32_t TP;      // emit positions per tile = THREADS*P - 2(w-1)
	int32_t one_tile_per_job;
};

struct SketchBatch {
	int32_t njobs;
	int64_t ntiles;
	const SketchJob *jobs;
	const int64_t *tile_base;  // [njobs+1] first tile of each job (NULL when one_tile_per_job)
	const char *buf;
	unsigned long long *status; // [ntiles] look-back words
	int32_t *ticket;
	int64_t *out_off;          // [njobs+1]
	uint64_t *out;             // x,y pairs
	int64_t out_cap;           // entries; records beyond are dropped but still counted
};

GD_DEV uint64_t sk_hash64(uint64_t key, uint64_t mask)
{ // GDiet-ShortReads/sketch.c:25-34
	key = (~key + (key << 21)) & mask;
	key = key ^ key >> 24;
	key = ((key + (key << 3)) + (key << 8)) & mask;
	key = key ^ key >> 14;
	key = ((key + (key << 2)) + (key << 4)) & mask;
	key = key ^ key >> 28;
	key = (key + (key << 31)) & mask;
	return key;
}

GD_DEV int sk_nt4(unsigned c)
{ // seq_nt4_table, GDiet-ShortReads/sketch.c:11-18
	unsigned u = c & 0xdfu; // fold case
	if (c < 4) return (int)c;
	return u == 'A' ? 0 : u == 'C' ? 1 : u == 'G' ? 2 : (u == 'T' || u == 'U') ? 3 : 4;
}

GD_DEV uint32_t sk_diet_len(uint32_t len, uint32_t shift, const SketchParams &S)
{ // GDiet-ShortReads/sketch.c:180-186,1942-1948
	if (len < shift) return 0;
	uint32_t rem = (len - shift) % (uint32_t)S.W, d = ((len - shift) / (uint32_t)S.W) * (uint32_t)S.ones;
	for (int i = 0; i < S.ones; ++i)
		if (S.ones_loc[i] < rem) ++d;
	return d;
}

GD_DEV uint32_t sk_real(uint32_t i, uint32_t shift, const SketchParams &S)
{ // get_real_location, GDiet-ShortReads/sketch.c:20-23
	uint32_t qd = i / (uint32_t)S.ones, rm = i - qd * (uint32_t)S.ones;
	return qd * (uint32_t)S.W + S.ones_loc[rm] + shift;
}

#define GD_SK_MAXU64 0xffffffffffffffffull

// Shared memory of one block: X[NP], M[NP] (uint64), aux[NP] (uint16: run | strand<<15), scan scratch.
template <int THREADS, int P> struct SketchSmem {
	enum { NP = THREADS * P };
	uint64_t X[NP];
	uint64_t M[NP];
	uint16_t aux[NP];
	int32_t warp_cnt[32];
	long long excl;
	int32_t tile;
};

template <int THREADS, int P>
GD_DEV void sketch_tile_body(const SketchParams &S, const SketchBatch &B, SketchSmem<THREADS, P> *sm)
{
	const int NP = THREADS * P;
	const int tid = thread_idx(), lane = tid & 31, wid = tid >> 5;
	const int w = S.w, k = S.k, full_run = w + k - 1;
	for (;;) {
		if (tid == 0) sm->tile = atomic_add(B.ticket, 1);
		sync_block();
		const long long tile = sm->tile;
		if (tile >= B.ntiles) break;
		// tile -> (job, first emit position)
		int job;
		long long chunk;
		if (S.one_tile_per_job) job = (int)tile, chunk = 0;
		else {
			int lo = 0, hi = B.njobs; // last job with tile_base[job] <= tile
			while (hi - lo > 1) {
				int mid = (lo + hi) >> 1;
				if (B.tile_base[mid] <= tile) lo = mid;
				else hi = mid;
			}
			job = lo, chunk = tile - B.tile_base[lo];
		}
		const SketchJob J = B.jobs[job];
		const char *seq = B.buf + J.seq_off;
		const uint32_t shift = (uint32_t)J.shift;
		const long long dl = (long long)sk_diet_len((uint32_t)J.len, shift, S);
		const long long i0 = chunk * S.TP;          // first emit position of the tile
		const long long jbase = i0 - (w - 1);       // position held in slot 0
		// ---- per-thread rolling pass over P positions ----
		{
			const long long j0 = jbase + (long long)tid * P;
			uint64_t fw = 0, rv = 0;
			int run = 0;
			// warm-up: the N-free run that ends just before j0 (only its last w+k-2 bases matter)
			if (j0 > 0 && j0 <= dl) {
				long long back = j0 - 1, stop = j0 - (full_run - 1);
				if (stop < 0) stop = 0;
				while (back >= stop && sk_nt4((unsigned char)seq[sk_real((uint32_t)back, shift, S)]) < 4) --back;
				run = (int)(j0 - 1 - back);
				int take = run < k - 1 ? run : k - 1;
				for (long long j = j0 - take; j < j0; ++j) {
					uint64_t c = (uint64_t)sk_nt4((unsigned char)seq[sk_real((uint32_t)j, shift, S)]);
					fw = (fw << 2 | c) & S.mask;
					rv = (rv >> 2) | (3ull ^ c) << (2 * (k - 1));
				}
			}
			for (int p = 0; p < P; ++p) {
				const long long j = j0 + p;
				uint64_t X = GD_SK_MAXU64;
				uint32_t a = 0;
				if (j >= 0 && j < dl) {
					int c = sk_nt4((unsigned char)seq[sk_real((uint32_t)j, shift, S)]);
					if (c < 4) {
						fw = (fw << 2 | (uint64_t)c) & S.mask;
						rv = (rv >> 2) | (uint64_t)(3 ^ c) << (2 * (k - 1));
						if (run < 0x7fff) ++run;
						if (run >= k && fw != rv) {
							const int z = fw < rv ? 0 : 1;
							X = sk_hash64(z ? rv : fw, S.mask) << 8 | (uint64_t)k;
							a = (uint32_t)z << 15;
						}
					} else run = 0;
					a |= (uint32_t)run;
				}
				sm->X[tid * P + p] = X;
				sm->aux[tid * P + p] = (uint16_t)a;
			}
		}
		sync_block();
		// ---- minimum of every full window ending at e ----
		for (int p = 0; p < P; ++p) {
			const int s = tid * P + p;
			const long long e = jbase + s;
			uint64_t m = 0; // 0 = "no full window ends here" (X >= 1 always)
			if (e >= 0 && e < dl && s >= w - 1 && (sm->aux[s] & 0x7fff) >= full_run) {
				m = GD_SK_MAXU64;
				for (int d = 0; d < w; ++d) {
					uint64_t x = sm->X[s - d];
					m = x < m ? x : m;
				}
				if (m == GD_SK_MAXU64) m = 0;
			}
			sm->M[s] = m;
		}
		sync_block();
		// ---- emission: X(i) equals the largest full-window minimum among windows containing i ----
		uint32_t emit = 0;
		int cnt = 0;
		for (int p = 0; p < P; ++p) {
			const int s = tid * P + p;
			const long long i = jbase + s;
			if (s >= w - 1 && s < w - 1 + S.TP && i < dl) {
				const uint64_t x = sm->X[s];
				if (x != GD_SK_MAXU64) {
					uint64_t mx = 0;
					for (int d = 0; d < w && s + d < NP; ++d) {
						uint64_t m = sm->M[s + d];
						mx = m > mx ? m : mx;
					}
					if (mx == x) emit |= 1u << p, ++cnt;
				}
			}
		}
		// ---- block exclusive scan of cnt ----
		int inc = cnt;
		for (int d = 1; d < 32; d <<= 1) {
			int o = (int)shfl_up(0xffffffffu, (uint32_t)inc, d, 32);
			if (lane >= d) inc += o;
		}
		if (lane == 31) sm->warp_cnt[wid] = inc;
		sync_block();
		int wbase = 0, total = 0;
		for (int i = 0; i < THREADS / 32; ++i) {
			int c = sm->warp_cnt[i];
			if (i < wid) wbase += c;
			total += c;
		}
		const int local = wbase + inc - cnt;
		// ---- decoupled look-back over tiles ----
		if (tid == 0) {
			long long excl = 0;
			if (tile == 0) {
				st_volatile(&B.status[0], (2ull << 62) | (unsigned long long)total);
			} else {
				st_volatile(&B.status[tile], (1ull << 62) | (unsigned long long)total);
				fence();
				long long pt = tile - 1;
				for (;;) {
					unsigned long long sv = ld_volatile(&B.status[pt]);
					unsigned long long state = sv >> 62;
					if (state == 0) continue;
					excl += (long long)(sv & 0x3fffffffffffffffull);
					if (state == 2) break;
					--pt;
				}
				st_volatile(&B.status[tile], (2ull << 62) | (unsigned long long)(excl + total));
			}
			sm->excl = excl;
			if (chunk == 0) B.out_off[job] = excl;
			if (tile == B.ntiles - 1) B.out_off[B.njobs] = excl + total;
		}
		sync_block();
		const long long obase = sm->excl + local;
		int o = 0;
		for (int p = 0; p < P; ++p)
			if (emit >> p & 1) {
				const int s = tid * P + p;
				const long long i = jbase + s, dst = obase + o;
				if (dst < B.out_cap) {
					const uint64_t y = (uint64_t)J.rid << 32 | (uint64_t)sk_real((uint32_t)i, shift, S) << 1 |
					                   (uint64_t)(sm->aux[s] >> 15);
					B.out[2 * dst] = sm->X[s];
					B.out[2 * dst + 1] = y;
				}
				++o;
			}
		sync_block(); // shared memory is reused by the next tile
	}
}

} // namespace gd
