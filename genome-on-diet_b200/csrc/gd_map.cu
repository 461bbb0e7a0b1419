// gd_map.cu -- short-read mapping between the two hot kernels, on the device (SURVEY.md 8 F1 + F2):
// what GDiet-ShortReads/map.c:mm_map_frag does from mm_sketch2 up to the ksw_extz_t of every candidate.
//
//   reads (ASCII, HBM) --sketch kernel--> per (read, shift) minimizer lists
//   K1 gd_sr_seed_kernel   warp per read: mm_get_shift (seed.c:166-194), mm_seed_mz_flt (seed.c:5-29),
//                          index lookups + mm_seed_select (seed.c:36-113,143-164) -> seed table, hit count
//   scan of the hit counts -> one flat hit array for the batch
//   K2 gd_sr_vote_kernel   warp per read: collect_seed_hits (map.c:261-356; the k-way merge is a sort by
//                          target, done as a bitonic network over the warp), vote x2 (map.c:447-584, the
//                          sequential cluster scan runs on lane 0 over 32-hit chunks staged in shared memory),
//                          window arithmetic (map.c:764-839) -> candidates
//   K3 gd_sr_window_kernel warp per candidate: query codes (map.c:737-757), target codes from the 4-bit
//                          reference (index.c:157-166), exact_match_sse (map.c:873-915)
//   DP kernel on the candidates that are not exact (gd_ksw_run_device), flag KSW_EZ_APPROX_MAX
//   K4 scores + CIGARs into the dense, input-ordered output
//
// Only three words per batch cross back to the host before the results do (the totals that size the next
// stage's arrays).
#include "gd_ctx.h"
#include "gd_index.cuh"
#include "gd_sketch.cuh"
#include "gd_sam_core.h"
#include <cub/cub.cuh>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <thread>
#include <time.h>
#include <utility>
#include <vector>

using namespace gd;

#define SR_FLT 0x80000000u // seed filtered (mm_seed_mz_flt / mm_seed_select / max_occ)
#define SR_WARPS 4         // warps per block of the per-read kernels
#define SR_MAX_LOC 32
#define SR_ABORTED (-1000) // internal: this lane only gave up because the other lane of the call failed

struct SrParams {
	int32_t W, JW, crop, k, frag_mode;
	uint32_t max_nb_seeds, bw;
	float bw_frac;          // short reads: per-read band (map.c:624-631) when bw_max > 0
	uint32_t bw_min, bw_max;
	float min_cnt, rec_frac, q_occ_frac;
	int32_t af_max_loc, mid_occ, max_max_occ, occ_dist, for_only, rev_only, a;
	int32_t stride; // bytes per candidate in the query / target code buffers
	// long-read tree (GDiet-LongReads/map.c)
	uint32_t vt_dis, vt_nb_loc, max_max_gap, max_min_gap;
	float vt_cov, vt_df1, vt_df2, vt_f;
	int32_t cnt_table; // the seed kernel was launched with its shared-memory counter table (mm_seed_mz_flt prefilter)
};

// band width / vote distance of a short read of `qlen` bases, map.c:624-631
__host__ __device__ static inline uint32_t sr_read_bw(const SrParams &P, int qlen)
{
	if (P.bw_max == 0) return P.bw;
	uint32_t bw = (uint32_t)((float)qlen * P.bw_frac);
	if (P.bw_min > bw) bw = P.bw_min;
	else if (P.bw_max < bw) bw = P.bw_max;
	return bw;
}

struct SrRead { // per-read state between K1 and K2
	uint32_t shift, n0, n_mv, ext;
};

struct SrVt { // vt_t, map.c:433-440
	uint32_t chrom_id;
	int32_t target_loc;
	uint32_t fq, lq, str, score;
};

// --------------------------------------------------------------------------------------------
// K1: shift, seed filters, lookups
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t warp_sum(uint32_t v)
{
	for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
	return v;
}

// mm_seed_select (seed.c:67-113) over the present seeds of one read, in raw-list index space; lane 0 only.
__device__ void sr_seed_select(const uint64_t *mv, uint32_t *sn, int n0, int qlen, int max_occ, int max_max_occ, int dist)
{
	int present = 0, m = 0;
	for (int e = 0; e < n0; ++e) {
		const uint32_t v = sn[e];
		if (v == 0 || (v & SR_FLT)) continue;
		++present;
		if ((int)v > max_occ) ++m;
	}
	if (present <= 1 || m == 0) return;
	unsigned long long b[128];
	int last0 = -1, st = -1, cnt = 0; // last low-occurrence seed, first seed of the streak, streak length
	for (int e = 0; e <= n0; ++e) {
		if (e < n0) {
			const uint32_t v = sn[e];
			if (v == 0 || (v & SR_FLT)) continue;
			if ((int)v > max_occ) {
				if (cnt++ == 0) st = e;
				continue;
			}
		}
		if (cnt > 0) {
			const int ps = last0 < 0 ? 0 : (int)((uint32_t)mv[2 * last0 + 1] >> 1);
			const int pe = e == n0 ? qlen : (int)((uint32_t)mv[2 * e + 1] >> 1);
			int max_high_occ = (int)((double)(pe - ps) / dist + .499);
			if (max_high_occ > 0) {
				if (max_high_occ > 128) max_high_occ = 128;
				int k = 0, j = st;
				for (; j < e && k < max_high_occ; ++j) {
					const uint32_t v = sn[j];
					if (v == 0 || (v & SR_FLT)) continue;
					b[k++] = (unsigned long long)v << 32 | (uint32_t)j;
				}
				for (; j < e; ++j) { // the heap top of seed.c:94-99 is the maximum of b[]
					const uint32_t v = sn[j];
					if (v == 0 || (v & SR_FLT)) continue;
					int top = 0;
					for (int t = 1; t < k; ++t)
						if (b[t] > b[top]) top = t;
					if ((int)v < (int)(b[top] >> 32)) b[top] = (unsigned long long)v << 32 | (uint32_t)j;
				}
				for (int t = 0; t < k; ++t) sn[(uint32_t)b[t]] |= 0x40000000u; // chosen (flt = 1 before the xor)
			}
			for (int j = st; j < e; ++j) { // flt ^= 1, then the max_max_occ rule
				uint32_t v = sn[j];
				if (v == 0 || (v & SR_FLT)) continue;
				const bool chosen = (v & 0x40000000u) != 0;
				v &= 0x3fffffffu;
				if (!chosen || (int)v > max_max_occ) v |= SR_FLT;
				sn[j] = v;
			}
		}
		last0 = e, cnt = 0;
	}
}

__global__ void __launch_bounds__(SR_WARPS * 32) gd_sr_seed_kernel(IndexDev I, SrParams P, int n, const int32_t *len,
                                                                  const int64_t *job_off, const uint64_t *raw, const int64_t *c3,
                                                                  const int64_t *c2, const uint32_t *ret3, uint32_t *seed_n,
                                                                  uint32_t *seed_first, SrRead *rd, uint32_t *n_hits)
{
	extern __shared__ uint32_t s_cnt_dyn[]; // SR_WARPS x 1024 counters when P.cnt_table (long reads / small mid_occ), else nothing
	const int lane = threadIdx.x & 31, warps = (gridDim.x * blockDim.x) >> 5;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const int64_t *jo = job_off + (size_t)i * P.JW;
		// ---- mm_get_shift: the shift whose mm_sketch2 list has the most index hits (first strict maximum)
		uint32_t best = 0, shift = 0;
		for (int s = 0; s < P.W; ++s) {
			const int64_t base = (P.crop && s == 0) ? jo[P.W] : jo[s];
			const int64_t cnt = c2[(size_t)i * P.W + s];
			uint32_t sum = 0, f;
			for (int64_t e = lane; e < cnt; e += 32) sum += idx_get(I, raw[2 * (base + e)] >> 8, f);
			sum = warp_sum(sum);
			if (sum > best) best = sum, shift = (uint32_t)s;
		}
		// ---- the seeds: mm_sketch3(shift)
		const int64_t base = jo[shift];
		const int n0 = (int)c3[(size_t)i * P.W + shift];
		const uint64_t *mv = raw + 2 * base;
		uint32_t *sn = seed_n + base, *sf = seed_first + base;
		uint32_t removed = 0;
		for (int e = lane; e < n0; e += 32) sn[e] = 0;
		__syncwarp();
		if (P.q_occ_frac > 0.f && P.mid_occ > 0 && n0 > P.mid_occ) {
			// mm_seed_mz_flt (seed.c:5-29): a minimizer value that occurs more than mid_occ times (and in more than q_occ_frac
			// of the seeds) INSIDE the read loses all its seeds.  Every long read comes here (thousands of seeds), but
			// almost no value qualifies: a 1024-counter table in shared memory (an upper bound of every value's count) picks
			// the few seeds that need the exact count, which the warp then does over the whole list.
			uint32_t *cnt = s_cnt_dyn + (size_t)(threadIdx.x >> 5) * 1024;
			if (P.cnt_table) {
				for (int j = lane; j < 1024; j += 32) cnt[j] = 0;
				__syncwarp();
				for (int e = lane; e < n0; e += 32) atomicAdd(&cnt[(uint32_t)((mv[2 * e] * 0x9E3779B97F4A7C15ull) >> 54)], 1u);
				__syncwarp();
			}
			for (int e0 = 0; e0 < n0; e0 += 32) {
				const int e = e0 + lane;
				const uint64_t xe = e < n0 ? mv[2 * e] : 0;
				// (without the table every seed is a candidate: the exact count decides)
				uint32_t cand = __ballot_sync(0xffffffffu, e < n0 && (!P.cnt_table || (int)cnt[(uint32_t)((xe * 0x9E3779B97F4A7C15ull) >> 54)] > P.mid_occ));
				while (cand) {
					const int src = __ffs(cand) - 1;
					cand &= cand - 1;
					const uint64_t x = (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)xe, src) |
					                   (uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(xe >> 32), src) << 32;
					uint32_t c = 0;
					for (int j = lane; j < n0; j += 32) c += mv[2 * j] == x;
					c = warp_sum(c);
					if ((int)c > P.mid_occ && (float)(int)c > (float)n0 * P.q_occ_frac) {
						if (lane == 0) sn[e0 + src] = SR_FLT;
						++removed;
					}
				}
			}
			__syncwarp();
		}
		int high = 0;
		for (int e = lane; e < n0; e += 32) {
			if (sn[e] & SR_FLT) continue;
			uint32_t f;
			const uint32_t c = idx_get(I, mv[2 * e] >> 8, f);
			sn[e] = c, sf[e] = f;
			high |= (int)c > P.mid_occ;
		}
		high = __any_sync(0xffffffffu, high);
		__syncwarp();
		if (high) { // mm_collect_matches2, seed.c:149-155
			if (P.occ_dist > 0 && P.max_max_occ > P.mid_occ) {
				if (lane == 0) sr_seed_select(mv, sn, n0, len[i], P.mid_occ, P.max_max_occ, P.occ_dist);
			} else {
				for (int e = lane; e < n0; e += 32)
					if (!(sn[e] & SR_FLT) && (int)sn[e] > P.mid_occ) sn[e] |= SR_FLT;
			}
			__syncwarp();
		}
		uint32_t na = 0;
		for (int e = lane; e < n0; e += 32)
			if (!(sn[e] & SR_FLT)) na += sn[e];
		na = warp_sum(na);
		if (lane == 0) {
			SrRead r;
			r.shift = shift, r.n0 = (uint32_t)n0, r.n_mv = (uint32_t)n0 - removed, r.ext = ret3[(size_t)i * P.W + shift];
			rd[i] = r;
			n_hits[i] = na;
		}
	}
}

// --------------------------------------------------------------------------------------------
// K2: hits, sort, vote, windows
// --------------------------------------------------------------------------------------------
// Ascending sort of (target, query) by target with one warp.  Ties in target are ordered by descending query
// position: that is the order the reference's merge_sort leaves (map.c:176-255, merge_locations takes the later
// run -- the seed with the larger query position -- on a tie), which the long-read vote_2 depends on.  (The
// short-read vote does not depend on the tie order: tied hits join the same cluster and only the min / max of their
// query positions are kept.)
__device__ __forceinline__ bool hit_less(uint64_t at, uint32_t aq, uint64_t bt, uint32_t bq) { return at < bt || (at == bt && aq > bq); }
__device__ void sr_sort_hits(uint64_t *t, uint32_t *q, int n, int lane)
{
	if (n <= 1) return;
	if (n <= 32) {
		uint64_t key = lane < n ? t[lane] : ~0ull;
		uint32_t val = lane < n ? q[lane] : 0;
		for (int k = 2; k <= 32; k <<= 1)
			for (int j = k >> 1; j > 0; j >>= 1) {
				const uint64_t ok = __shfl_xor_sync(0xffffffffu, key, j);
				const uint32_t ov = __shfl_xor_sync(0xffffffffu, val, j);
				const bool take_min = ((lane & k) == 0) == ((lane & j) == 0);
				if (take_min ? hit_less(ok, ov, key, val) : hit_less(key, val, ok, ov)) key = ok, val = ov;
			}
		if (lane < n) t[lane] = key, q[lane] = val;
		__syncwarp();
		return;
	}
	int P2 = 64;
	while (P2 < n) P2 <<= 1;
	// all comparators ascending (flip network): virtual elements >= n are +inf and never move
	for (int k = 2; k <= P2; k <<= 1) {
		for (int p = lane; p < P2 / 2; p += 32) {
			const int blk = p / (k >> 1), o = p % (k >> 1);
			const int l = blk * k + o, h = blk * k + k - 1 - o;
			if (h < n) {
				const uint64_t a = t[l], b = t[h];
				const uint32_t qa = q[l], qb = q[h];
				if (hit_less(b, qb, a, qa)) t[l] = b, t[h] = a, q[l] = qb, q[h] = qa;
			}
		}
		__syncwarp();
		for (int j = k >> 2; j > 0; j >>= 1) {
			for (int p = lane; p < P2 / 2; p += 32) {
				const int l = (p / j) * 2 * j + (p % j), h = l + j;
				if (h < n) {
					const uint64_t a = t[l], b = t[h];
					const uint32_t qa = q[l], qb = q[h];
					if (hit_less(b, qb, a, qa)) t[l] = b, t[h] = a, q[l] = qb, q[h] = qa;
				}
			}
			__syncwarp();
		}
	}
}

// collect_seed_hits, map.c:284-311: the hits of every unfiltered seed, forward strand from the front of the read's
// slice of the hit array, reverse strand from its back (the slice has room for exactly all of them)
__device__ __forceinline__ void map_fill_hits(const IndexDev &I, const SrParams &P, uint32_t ext, const uint64_t *mv, const uint32_t *sn,
                                              const uint32_t *sf, int n0, uint64_t *t, uint32_t *q, uint32_t na, uint32_t &nf, uint32_t &nr,
                                              int lane)
{
	const uint32_t lt = (1u << lane) - 1;
	auto put = [&](bool valid, uint64_t r, uint32_t qp) {
		const uint32_t qpos = qp >> 1, loc = (uint32_t)r >> 1;
		const uint32_t str = (uint32_t)(r & 1) ^ (qp & 1);
		const bool keep = valid && !(str ? P.for_only : P.rev_only); // skip_seed, map.c:121-127
		const uint32_t fm = __ballot_sync(0xffffffffu, keep && !str), rm = __ballot_sync(0xffffffffu, keep && str);
		if (keep) {
			if (str) {
				const uint32_t slot = na - 1 - (nr + __popc(rm & lt));
				t[slot] = (r >> 32) << 32 | (uint32_t)(loc + qpos), q[slot] = qpos;
			} else {
				const uint32_t slot = nf + __popc(fm & lt);
				t[slot] = (r >> 32) << 32 | (uint32_t)(loc + ext - qpos), q[slot] = qpos;
			}
		}
		nf += __popc(fm), nr += __popc(rm);
	};
	for (int e0 = 0; e0 < n0; e0 += 32) {
		const int e = e0 + lane;
		uint32_t c = e < n0 ? sn[e] : 0;
		if (c & SR_FLT) c = 0;
		const uint32_t qp = e < n0 ? (uint32_t)mv[2 * e + 1] : 0, first = e < n0 ? sf[e] : 0;
		put(c == 1, c == 1 ? I.pos[first] : 0, qp);
		uint32_t mm = __ballot_sync(0xffffffffu, c > 1);
		while (mm) {
			const int src = __ffs(mm) - 1;
			mm &= mm - 1;
			const uint32_t bc = __shfl_sync(0xffffffffu, c, src), bf = __shfl_sync(0xffffffffu, first, src);
			const uint32_t bq = __shfl_sync(0xffffffffu, qp, src);
			for (uint32_t j0 = 0; j0 < bc; j0 += 32) {
				const uint32_t j = j0 + lane;
				put(j < bc, j < bc ? I.pos[bf + j] : 0, bq);
			}
		}
	}
}

struct SrVoteState {
	unsigned out_len;
	SrVt recovery;
};

// the common tail of both branches of vote() (map.c:475-520 and :527-565); lane 0 only
__device__ __forceinline__ void sr_vt_emit(SrVt *pot, SrVoteState &S, uint64_t target_loc, uint32_t fq, uint32_t lq,
                                           unsigned counter, int str, int32_t tmp_ext, unsigned thr, unsigned max_loc,
                                           unsigned rec_thr)
{
	SrVt v;
	v.chrom_id = (uint32_t)(target_loc >> 32);
	v.target_loc = (int32_t)(uint32_t)target_loc + (str ? 0 : -tmp_ext);
	v.fq = fq, v.lq = lq, v.str = (uint32_t)str, v.score = counter;
	if (counter > thr) {
		if (S.out_len == max_loc) {
			if (pot[S.out_len - 1].score >= counter) return;
		} else ++S.out_len;
		pot[S.out_len - 1] = v;
		for (unsigned k = S.out_len - 1; k > 0; k--) {
			if (pot[k].score > pot[k - 1].score) {
				const SrVt t = pot[k];
				pot[k] = pot[k - 1], pot[k - 1] = t;
			} else break;
		}
	} else if (S.out_len == 0 && counter > rec_thr && counter > S.recovery.score) S.recovery = v;
}

__device__ void sr_vote(const uint64_t *ht, const uint32_t *hq, unsigned len, int str, SrVt *pot, SrVoteState &S, uint64_t *s_t,
                        uint32_t *s_q, unsigned dist, int32_t tmp_ext, unsigned thr, unsigned max_loc, unsigned rec_thr, int lane)
{
	if (len == 0) return;
	unsigned counter = 0;
	uint64_t target_loc = 0;
	uint32_t fq = 0, lq = 0;
	for (unsigned c0 = 0; c0 < len; c0 += 32) {
		if (c0 + lane < len) s_t[lane] = ht[c0 + lane], s_q[lane] = hq[c0 + lane];
		__syncwarp();
		if (lane == 0) {
			const unsigned m = len - c0 < 32 ? len - c0 : 32;
			for (unsigned j = 0; j < m; ++j) {
				const uint64_t ct = s_t[j];
				const uint32_t cq = s_q[j];
				if (counter == 0) {
					target_loc = ct, fq = lq = cq, counter = 1;
				} else if (ct - target_loc <= dist) {
					counter++;
					if (cq < fq) target_loc = ct, fq = cq;
					if (cq > lq) lq = cq;
				} else {
					sr_vt_emit(pot, S, target_loc, fq, lq, counter, str, tmp_ext, thr, max_loc, rec_thr);
					target_loc = ct, fq = lq = cq, counter = 1;
				}
			}
		}
		__syncwarp();
	}
	if (lane == 0) sr_vt_emit(pot, S, target_loc, fq, lq, counter, str, tmp_ext, thr, max_loc, rec_thr);
}

__global__ void __launch_bounds__(SR_WARPS * 32) gd_sr_vote_kernel(IndexDev I, SrParams P, int n, const int32_t *len,
                                                                  const int64_t *job_off, const uint64_t *raw,
                                                                  const uint32_t *seed_n, const uint32_t *seed_first,
                                                                  const SrRead *rd, const int64_t *hits_off, uint64_t *ht,
                                                                  uint32_t *hq, gd_sr_cand_t *cand_tmp, uint32_t *n_cand)
{
	__shared__ uint64_t s_t[SR_WARPS][32];
	__shared__ uint32_t s_q[SR_WARPS][32];
	__shared__ SrVt s_pot[SR_WARPS][SR_MAX_LOC];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = (gridDim.x * blockDim.x) >> 5;
	const uint32_t lt = (1u << lane) - 1;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const SrRead R = rd[i];
		const int64_t base = job_off[(size_t)i * P.JW + R.shift];
		const uint64_t *mv = raw + 2 * base;
		const uint32_t *sn = seed_n + base, *sf = seed_first + base;
		const int64_t hoff = hits_off[i];
		const uint32_t na = (uint32_t)(hits_off[i + 1] - hoff);
		const int n0 = (int)R.n0;
		uint64_t *t = ht + hoff;
		uint32_t *q = hq + hoff;
		uint32_t nf = 0, nr = 0;
		map_fill_hits(I, P, R.ext, mv, sn, sf, n0, t, q, na, nf, nr, lane);
		__syncwarp();
		uint64_t *tr = t + (na - nr);
		uint32_t *qr = q + (na - nr);
		sr_sort_hits(t, q, (int)nf, lane);
		sr_sort_hits(tr, qr, (int)nr, lane);
		// ---- voting, map.c:665-699
		const uint32_t qlen_sum = (uint32_t)len[i];
		const bool frag = P.frag_mode && R.ext < qlen_sum;
		unsigned thr = frag ? (unsigned)((float)P.max_nb_seeds * P.min_cnt) : (unsigned)((float)R.n_mv * P.min_cnt);
		const unsigned rec_thr = frag ? (unsigned)((float)P.max_nb_seeds * P.rec_frac) : (unsigned)((float)R.n_mv * P.rec_frac);
		if (thr == 0) thr = 1;
		SrVoteState S;
		S.out_len = 0;
		S.recovery.score = 0, S.recovery.chrom_id = 0, S.recovery.target_loc = 0, S.recovery.fq = S.recovery.lq = S.recovery.str = 0;
		SrVt *pot = s_pot[wib];
		const uint32_t bw = sr_read_bw(P, (int)qlen_sum);
		sr_vote(t, q, nf, 0, pot, S, s_t[wib], s_q[wib], bw, (int32_t)R.ext, thr, (unsigned)P.af_max_loc, rec_thr, lane);
		sr_vote(tr, qr, nr, 1, pot, S, s_t[wib], s_q[wib], bw, (int32_t)R.ext, thr, (unsigned)P.af_max_loc, rec_thr, lane);
		if (lane == 0 && S.out_len == 0 && S.recovery.score != 0) pot[0] = S.recovery, S.out_len = 1; // map.c:692-699
		const unsigned nb = __shfl_sync(0xffffffffu, S.out_len, 0);
		__syncwarp();
		// ---- candidate windows, map.c:764-839 (one lane per candidate)
		gd_sr_cand_t c;
		bool keep = false;
		if (lane < (int)nb) {
			const SrVt v = pot[lane];
			const int str = (int)v.str, k = P.k;
			const int32_t tlen = (int32_t)I.seq_len[v.chrom_id];
			int32_t loc = v.target_loc;
			if (str) loc -= (k - 1);
			int32_t target_start = loc, target_end = loc;
			uint32_t start_offset, end_offset;
			keep = true;
			if (qlen_sum > 300) {
				if (v.fq == v.lq) keep = false;
				start_offset = v.fq - (uint32_t)(k - 1);
				end_offset = v.lq;
				if (str) {
					target_end -= (int32_t)start_offset, target_start -= (int32_t)end_offset;
					if (target_start < 0) end_offset += (uint32_t)target_start, target_start = 0;
				} else {
					target_start += (int32_t)start_offset, target_end += (int32_t)end_offset;
					if (target_end + 1 > tlen) end_offset = (uint32_t)(tlen - 1 - target_start) + start_offset, target_end = tlen - 1;
				}
			} else if (str) {
				if (target_end > tlen - 1) start_offset = (uint32_t)(target_end - (tlen - 1)), target_end = tlen - 1;
				else start_offset = 0;
				if ((uint32_t)target_end < qlen_sum - start_offset - 1) end_offset = start_offset + (uint32_t)target_end, target_start = 0;
				else end_offset = qlen_sum - 1, target_start = target_end - (int32_t)(end_offset - start_offset);
			} else {
				if (target_start < 0) start_offset = (uint32_t)(-target_start), target_start = 0;
				else start_offset = 0;
				if ((uint32_t)(tlen - target_start) < qlen_sum - start_offset)
					end_offset = (uint32_t)(tlen - 1 - target_start) + start_offset, target_end = tlen - 1;
				else end_offset = qlen_sum - 1, target_end = target_start + (int32_t)(end_offset - start_offset);
			}
			c.rid = (int32_t)v.chrom_id, c.rs = target_start, c.re = target_end + 1, c.qs = (int32_t)start_offset;
			c.qe = (int32_t)end_offset + 1, c.rev = str, c.votes = (int32_t)v.score, c.first_q = (int32_t)v.fq, c.last_q = (int32_t)v.lq;
			c.exact = 0, c.score = 0, c.n_cigar = 0, c.cigar_off = 0;
			c.reserved[0] = i, c.reserved[1] = 0, c.reserved[2] = 0;
		}
		const uint32_t km = __ballot_sync(0xffffffffu, keep);
		if (keep) cand_tmp[(size_t)i * P.af_max_loc + __popc(km & lt)] = c;
		if (lane == 0) n_cand[i] = (uint32_t)__popc(km);
		__syncwarp();
	}
}

// --------------------------------------------------------------------------------------------
// K2, long-read tree (GDiet-LongReads/map.c:1052-1590,1654-1713): two voting rounds, density / score filters,
// candidate chaining, windows.  The cluster scans are sequential by definition and run on lane 0 over 32-hit
// chunks the warp stages in shared memory; everything after them works on at most vt_nb_loc + 2 candidates.
// --------------------------------------------------------------------------------------------
struct LrVt { // vt_t, LR/map.c:1032-1045 (next as an index, -1 = NULL)
	uint32_t chrom_id;
	int32_t ft, lt; // first / last target location
	uint32_t fq, lq;
	uint32_t score;
	int32_t next;
	uint32_t str, concat;
};

__device__ __forceinline__ uint64_t lr_loc(int str, uint64_t target, uint32_t query, int32_t ext)
{ // LR/map.c:1064-1065
	return str ? (target - query) : target - (uint64_t)(uint32_t)(ext - (int32_t)query);
}

__device__ __forceinline__ void lr_emit(LrVt *seqs, unsigned &out_len, unsigned max_loc, uint64_t ft, uint64_t lt, uint32_t fq, uint32_t lq,
                                        unsigned counter, int str)
{ // LR/map.c:1095-1131 / :1145-1179
	if (out_len == max_loc) {
		if (seqs[out_len - 1].score >= counter) return;
	} else ++out_len;
	LrVt v;
	v.chrom_id = (uint32_t)(ft >> 32), v.ft = (int32_t)(uint32_t)ft, v.lt = (int32_t)(uint32_t)lt, v.fq = fq, v.lq = lq;
	v.score = counter, v.next = -1, v.str = (uint32_t)str, v.concat = 0;
	seqs[out_len - 1] = v;
	for (unsigned k = out_len - 1; k > 0; k--) {
		if (seqs[k].score > seqs[k - 1].score) {
			const LrVt t = seqs[k];
			seqs[k] = seqs[k - 1], seqs[k - 1] = t;
		} else break;
	}
}

// round == 1: vote (LR/map.c:1052-1182) into seqs / out_len;  round == 2: vote_2 (:1184-1271) into best, query range (qmin, qmax)
template <int ROUND>
__device__ void lr_vote(const uint64_t *ht, const uint32_t *hq, unsigned len, int str, LrVt *seqs, unsigned &out_len, LrVt &best,
                        uint64_t *s_t, uint32_t *s_q, uint32_t dist, int32_t ext, unsigned max_loc, uint32_t cov_thr, uint32_t qmin,
                        uint32_t qmax, int lane)
{
	if (len == 0) return;
	unsigned counter = 0;
	uint64_t ft = 0, lt = 0, ref_loc = 0;
	uint32_t fq = 0, lq = 0;
	auto close = [&]() {
		if (ROUND == 1) {
			if (lq - fq > cov_thr) lr_emit(seqs, out_len, max_loc, ft, lt, fq, lq, counter, str);
		} else if (counter > best.score && lq < qmax && fq > qmin) {
			best.chrom_id = (uint32_t)(ft >> 32), best.ft = (int32_t)(uint32_t)ft, best.lt = (int32_t)(uint32_t)lt;
			best.fq = fq, best.lq = lq, best.score = counter, best.next = -1, best.str = (uint32_t)str, best.concat = 0;
		}
	};
	for (unsigned c0 = 0; c0 < len; c0 += 32) {
		if (c0 + lane < len) s_t[lane] = ht[c0 + lane], s_q[lane] = hq[c0 + lane];
		__syncwarp();
		if (lane == 0) {
			const unsigned m = len - c0 < 32 ? len - c0 : 32;
			for (unsigned j = 0; j < m; ++j) {
				const uint64_t ct = s_t[j];
				const uint32_t cq = s_q[j];
				if (counter != 0 && ct - ref_loc <= dist) {
					if (ROUND == 1 || (cq < qmax && cq > qmin)) {
						const uint64_t l = lr_loc(str, ct, cq, ext);
						counter++;
						if (cq < fq) fq = cq, ref_loc = ct;
						if (cq > lq) lq = cq;
						if (l > lt) lt = l;
						if (l < ft) ft = l;
					}
				} else {
					if (counter != 0) close();
					ft = lt = lr_loc(str, ct, cq, ext), fq = lq = cq, ref_loc = ct, counter = 1;
				}
			}
		}
		__syncwarp();
	}
	if (lane == 0) close();
}

__global__ void __launch_bounds__(SR_WARPS * 32) gd_lr_vote_kernel(IndexDev I, SrParams P, int n, const int32_t *len,
                                                                  const int64_t *job_off, const uint64_t *raw,
                                                                  const uint32_t *seed_n, const uint32_t *seed_first,
                                                                  const SrRead *rd, const int64_t *hits_off, uint64_t *ht,
                                                                  uint32_t *hq, gd_sr_cand_t *cand_tmp, uint32_t *n_cand, int *max_span)
{
	__shared__ uint64_t s_t[SR_WARPS][32];
	__shared__ uint32_t s_q[SR_WARPS][32];
	__shared__ LrVt s_seq[SR_WARPS][SR_MAX_LOC + 2];
	__shared__ unsigned s_nb[SR_WARPS];
	const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, warps = (gridDim.x * blockDim.x) >> 5;
	for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
		const SrRead R = rd[i];
		const int64_t base = job_off[(size_t)i * P.JW + R.shift];
		const uint64_t *mv = raw + 2 * base;
		const uint32_t *sn = seed_n + base, *sf = seed_first + base;
		const int64_t hoff = hits_off[i];
		const uint32_t na = (uint32_t)(hits_off[i + 1] - hoff);
		uint64_t *t = ht + hoff;
		uint32_t *q = hq + hoff;
		uint32_t nf = 0, nr = 0;
		map_fill_hits(I, P, R.ext, mv, sn, sf, (int)R.n0, t, q, na, nf, nr, lane);
		__syncwarp();
		uint64_t *tr = t + (na - nr);
		uint32_t *qr = q + (na - nr);
		sr_sort_hits(t, q, (int)nf, lane);
		sr_sort_hits(tr, qr, (int)nr, lane);
		const uint32_t qlen_sum = (uint32_t)len[i];
		const int32_t ext = (int32_t)R.ext;
		const int k = P.k;
		const unsigned bw = P.bw;
		const uint32_t cov_thr = (uint32_t)((float)qlen_sum * P.vt_cov);
		LrVt *seqs = s_seq[wib];
		LrVt best;
		unsigned nb = 0;
		// ---- first round, LR/map.c:1342-1347
		lr_vote<1>(t, q, nf, 0, seqs, nb, best, s_t[wib], s_q[wib], P.vt_dis, ext, P.vt_nb_loc, cov_thr, 0, 0, lane);
		lr_vote<1>(tr, qr, nr, 1, seqs, nb, best, s_t[wib], s_q[wib], P.vt_dis, ext, P.vt_nb_loc, cov_thr, 0, 0, lane);
		uint32_t qrstart = qlen_sum, qrend = 0;
		int need2a = 0, need2b = 0;
		if (lane == 0) {
			if (nb > 0) { // density filter, LR/map.c:1354-1362 (the assignment is the reference's: seqs[i] = seqs[kept])
				unsigned df = 0;
				for (unsigned c = 0; c < nb; c++)
					if ((float)seqs[c].score > P.vt_df1 * (float)(seqs[c].lt - seqs[c].ft)) seqs[c] = seqs[df], df++;
				nb = df;
			}
			if (nb > 0) { // LR/map.c:1371-1400
				const unsigned filtering_threshold = (unsigned)((float)seqs[0].score * P.vt_f);
				for (unsigned c = 0; c < nb; c++) {
					if (seqs[c].score < filtering_threshold) {
						nb = c;
						break;
					}
					seqs[c].fq -= (uint32_t)(k - 1), seqs[c].ft -= (k - 1);
					seqs[c].next = -1, seqs[c].concat = 0;
					if (seqs[c].lq - seqs[c].fq + 0.5 * bw < seqs[c].lt - seqs[c].ft)
						seqs[c].lt = (int32_t)(seqs[c].ft + seqs[c].lq - seqs[c].fq + 0.5 * bw);
					if (seqs[c].fq < qrstart) qrstart = seqs[c].fq;
					if (seqs[c].lq > qrend) qrend = seqs[c].lq;
				}
				need2a = qrstart > cov_thr, need2b = qlen_sum - qrend > cov_thr;
			}
		}
		nb = __shfl_sync(0xffffffffu, nb, 0);
		if (nb == 0) { // nothing survives the first round: the reference returns before the second one (LR/map.c:1349-1369)
			if (lane == 0) n_cand[i] = 0;
			__syncwarp();
			continue;
		}
		need2a = __shfl_sync(0xffffffffu, need2a, 0), need2b = __shfl_sync(0xffffffffu, need2b, 0);
		qrstart = __shfl_sync(0xffffffffu, qrstart, 0), qrend = __shfl_sync(0xffffffffu, qrend, 0);
		// ---- second round on the uncovered ends of the read, LR/map.c:1402-1445
		for (int pass = 0; pass < 2; ++pass) {
			if (!(pass == 0 ? need2a : need2b)) continue;
			const uint32_t qmin = pass == 0 ? 0u : qrend, qmax = pass == 0 ? qrstart : qlen_sum;
			unsigned dummy = 0;
			best.score = 0, best.chrom_id = 0, best.ft = best.lt = 0, best.fq = best.lq = 0, best.next = -1, best.str = 0, best.concat = 0;
			lr_vote<2>(t, q, nf, 0, seqs, dummy, best, s_t[wib], s_q[wib], P.vt_dis, ext, 0, 0, qmin, qmax, lane);
			lr_vote<2>(tr, qr, nr, 1, seqs, dummy, best, s_t[wib], s_q[wib], P.vt_dis, ext, 0, 0, qmin, qmax, lane);
			if (lane == 0) {
				best.fq -= (uint32_t)(k - 1), best.ft -= (k - 1);
				if ((float)best.score > P.vt_df2 * (float)(best.lt - best.ft)) {
					if (best.lq - best.fq + 0.5 * bw < best.lt - best.ft) best.lt = (int32_t)(best.ft + best.lq - best.fq + 0.5 * bw);
					seqs[nb++] = best;
				}
			}
		}
		if (lane == 0) {
			// ---- which candidates continue each other, LR/map.c:1467-1590
			const unsigned G = P.max_max_gap, g = P.max_min_gap;
			for (unsigned a = 0; a < nb; a++) {
				LrVt &s1 = seqs[a];
				for (unsigned b = 0; b < nb; b++) {
					const LrVt &s2 = seqs[b];
					if (b == a || s2.concat != 0 || s1.str != s2.str || s1.chrom_id != s2.chrom_id) continue;
					bool take = false, better = false;
					if (s1.str) {
						if (s2.lq < s1.fq && s1.lt > s2.ft && s1.ft < s2.ft) {
							take = s2.lq + G > s1.fq;
							better = s1.next >= 0 && s2.lq > seqs[s1.next].lq;
						} else if (s2.lq < s1.fq && s1.lt < s2.ft) {
							take = (s2.lq + g > s1.fq || (uint32_t)s1.lt + g > (uint32_t)s2.ft) && s2.lq + G > s1.fq && (uint32_t)s1.lt + G > (uint32_t)s2.ft;
							better = s1.next >= 0 && s2.lq > seqs[s1.next].lq;
						} else if (s2.lq > s1.fq && s1.lt < s2.ft && s2.lq < s1.lq && s2.fq < s1.fq) {
							take = (uint32_t)s1.lt + G > (uint32_t)s2.ft;
							better = s1.next >= 0 && s2.lq < seqs[s1.next].lq;
						}
					} else {
						if (s1.lq < s2.fq && s1.lt > s2.ft && s1.ft < s2.ft) {
							take = s1.lq + G > s2.fq;
							better = s1.next >= 0 && s2.fq < seqs[s1.next].fq;
						} else if (s1.lq < s2.fq && s1.lt < s2.ft) {
							take = (s1.lq + g > s2.fq || (uint32_t)s1.lt + g > (uint32_t)s2.ft) && (uint32_t)s1.lt + G > (uint32_t)s2.ft && s1.lq + G > s2.fq;
							better = s1.next >= 0 && s2.fq < seqs[s1.next].fq;
						} else if (s1.lq > s2.fq && s1.lt < s2.ft && s1.fq < s2.fq && s1.lq < s2.lq) {
							take = (uint32_t)s1.lt + G > (uint32_t)s2.ft;
							better = s1.next >= 0 && s2.fq < seqs[s1.next].fq;
						}
					}
					if (take && (s1.next < 0 || better)) s1.next = (int32_t)b;
				}
				if (s1.next >= 0) { // adjust the boundaries of the pair, LR/map.c:1557-1589
					LrVt &s2 = seqs[s1.next];
					s2.concat = 1;
					if (s1.str) {
						if (s2.lq < s1.fq && s1.lt < s2.ft) {
							const uint32_t diffq = s1.fq - s2.lq, difft = (uint32_t)(s2.ft - s1.lt), mn = difft > diffq ? diffq : difft;
							s2.lq += mn, s1.lt += (int32_t)mn, s1.fq -= mn, s2.ft -= (int32_t)mn;
						}
					} else if (s1.lq < s2.fq && s1.lt < s2.ft) {
						const uint32_t diffq = s2.fq - s1.lq, difft = (uint32_t)(s2.ft - s1.lt), mn = difft > diffq ? diffq : difft;
						s1.lq += mn, s1.lt += (int32_t)mn, s2.fq -= mn, s2.ft -= (int32_t)mn;
					}
					if (s2.lt < s1.lt) s1.lt = s2.lt - 1;
				}
			}
			s_nb[wib] = nb;
		}
		__syncwarp();
		nb = s_nb[wib];
		// ---- windows, LR/map.c:1654-1713 (one lane per candidate)
		int span = 0;
		if (lane < (int)nb) {
			const LrVt v = seqs[lane];
			const int str = (int)v.str;
			const int32_t chrom_len = (int32_t)I.seq_len[v.chrom_id];
			uint32_t target_start = (uint32_t)v.ft, target_end = (uint32_t)v.lt, query_start, query_end;
			if (str) query_end = qlen_sum - 1 - v.fq, query_start = qlen_sum - 1 - v.lq;
			else query_start = v.fq, query_end = v.lq;
			if (!(qlen_sum > 300)) {
				if (target_start < query_start) query_start -= target_start, target_start = 0;
				else target_start -= query_start, query_start = 0;
				if ((uint32_t)chrom_len + query_end < qlen_sum + target_end) query_end += (uint32_t)chrom_len - target_end - 1, target_end = (uint32_t)chrom_len - 1;
				else target_end += qlen_sum - query_end - 1, query_end = qlen_sum - 1;
			}
			if (str) {
				const uint32_t tmp = qlen_sum - 1 - query_start;
				query_start = qlen_sum - 1 - query_end, query_end = tmp;
			}
			gd_sr_cand_t c;
			c.rid = (int32_t)v.chrom_id, c.rs = (int32_t)target_start, c.re = (int32_t)target_end + 1, c.qs = (int32_t)query_start;
			c.qe = (int32_t)query_end + 1, c.rev = str, c.votes = (int32_t)v.score, c.first_q = (int32_t)v.fq, c.last_q = (int32_t)v.lq;
			c.exact = 0, c.score = 0, c.n_cigar = 0, c.cigar_off = 0;
			c.reserved[0] = i, c.reserved[1] = v.next, c.reserved[2] = (int32_t)v.concat;
			cand_tmp[(size_t)i * (P.vt_nb_loc + 2) + lane] = c;
			const int ql = c.qe - c.qs, tl = c.re - c.rs;
			span = ql > tl ? ql : tl;
		}
		span = __reduce_max_sync(0xffffffffu, span);
		if (lane == 0) {
			n_cand[i] = nb;
			atomicMax(max_span, span);
		}
		__syncwarp();
	}
}

// candidates of all reads, dense and in input order
__global__ void gd_sr_compact_kernel(int n, int af, const gd_sr_cand_t *cand_tmp, const int64_t *cand_off, gd_sr_cand_t *cand)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int64_t o = cand_off[i];
	const int m = (int)(cand_off[i + 1] - o);
	const uint4 *src = (const uint4 *)(cand_tmp + (size_t)i * af);
	uint4 *dst = (uint4 *)(cand + o);
	for (int j = 0; j < m * 4; ++j) dst[j] = src[j];
}

// --------------------------------------------------------------------------------------------
// K3: query / target codes of every candidate + exact match; one warp per candidate
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gd_sr_window_kernel(IndexDev I, SrParams P, int64_t nc, const int64_t *off, const int32_t *len,
                                                          const char *buf, gd_sr_cand_t *cand, uint8_t *qbuf, uint8_t *tbuf,
                                                          uint32_t *need_dp)
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < nc; ci += warps) {
		gd_sr_cand_t *c = cand + ci;
		const int i = c->reserved[0], qs = c->qs, qe = c->qe, rev = c->rev, n = qe - qs;
		const char *rdp = buf + off[i];
		const uint64_t tb = I.seq_off[c->rid] + (uint64_t)c->rs;
		const int tl = c->re - c->rs; // == n for the short-read windows; independent for long reads
		const int t_in = (int)I.seq_len[c->rid] - c->rs; // mm_idx_getseq clips at the contig end (index.c:157-166)
		uint8_t *qo = qbuf + ci * (int64_t)P.stride, *to = tbuf + ci * (int64_t)P.stride;
		int diff = n != tl;
		for (int j = lane; j < n || j < tl; j += 32) {
			// map.c:737-757: forward = nt4, reverse strand = reversed and ^3 (N becomes 7)
			int qc = 0, tc = 0;
			if (j < n) qc = rev ? (sk_nt4((unsigned char)rdp[qe - 1 - j]) ^ 3) : sk_nt4((unsigned char)rdp[qs + j]), qo[j] = (uint8_t)qc;
			if (j < tl) tc = j < t_in ? (int)idx_base(I, tb + j) : 0, to[j] = (uint8_t)tc;
			diff |= qc != tc;
		}
		diff = __any_sync(0xffffffffu, diff);
		if (lane == 0) {
			const int exact = (len[i] < 300 && n > 0 && !diff) ? 1 : 0; // map.c:873
			c->exact = exact;
			need_dp[ci] = exact ? 0u : 1u;
		}
	}
}

// DP work list: the candidates that were not exact matches
__global__ void gd_sr_pairs_kernel(int64_t nc, int stride, const gd_sr_cand_t *cand, const int64_t *pair_off, int32_t *pqlen,
                                   int32_t *ptlen, int64_t *poff, int32_t *pw, SrParams P, const int32_t *len)
{
	const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ci >= nc) return;
	if (pair_off[ci + 1] == pair_off[ci]) return;
	const int64_t p = pair_off[ci];
	pqlen[p] = cand[ci].qe - cand[ci].qs, ptlen[p] = cand[ci].re - cand[ci].rs, poff[p] = ci * (int64_t)stride;
	if (pw) pw[p] = (int32_t)sr_read_bw(P, len[cand[ci].reserved[0]]); // the band of the candidate's read (map.c:624-631,923)
}

// K4a: scores and CIGAR lengths of every candidate
__global__ void gd_sr_scores_kernel(int64_t nc, SrParams P, const int32_t *len, gd_sr_cand_t *cand, const int64_t *pair_off,
                                    const gd_extz_t *ez, uint32_t *ncig)
{
	const int64_t ci = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (ci >= nc) return;
	gd_sr_cand_t *c = cand + ci;
	if (c->exact) c->score = len[c->reserved[0]] * P.a, c->n_cigar = 1; // map.c:901-903
	else {
		const gd_extz_t e = ez[pair_off[ci]];
		c->score = e.score, c->n_cigar = e.n_cigar;
	}
	ncig[ci] = c->n_cigar > 0 ? (uint32_t)c->n_cigar : 0u;
}

// K4b: CIGARs into the dense pool; one warp per candidate
__global__ void __launch_bounds__(128) gd_sr_cigars_kernel(int64_t nc, gd_sr_cand_t *cand, const int64_t *pair_off, const int64_t *cig_off,
                                                          const uint32_t *dp_cigar, int cigar_stride, uint32_t *pool, int64_t pool_cap)
{
	const int lane = threadIdx.x & 31;
	const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
	for (int64_t ci = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; ci < nc; ci += warps) {
		gd_sr_cand_t *c = cand + ci;
		const int64_t o = cig_off[ci];
		const int m = (int)(cig_off[ci + 1] - o);
		if (lane == 0) // reserved[] leaves as {next, concat, 0} (long reads; zeros for short reads)
			c->cigar_off = (int32_t)o, c->reserved[0] = c->reserved[1], c->reserved[1] = c->reserved[2], c->reserved[2] = 0;
		if (o + m > pool_cap) continue;
		if (c->exact) {
			if (lane == 0) pool[o] = (uint32_t)(c->qe - c->qs) << 4; // <len>M
		} else {
			const uint32_t *src = dp_cigar + pair_off[ci] * cigar_stride;
			for (int j = lane; j < m; j += 32) pool[o + j] = src[j];
		}
	}
}

// --------------------------------------------------------------------------------------------
// K5: the post-DP stage ON THE DEVICE (short reads): mm_update_extra ... mm_write_sam3 for every read of the slice, one
// thread per read running gd_sam_core.h.  A count pass (on a scratch copy of the CIGAR pool: mm_fix_cigar edits CIGARs in
// place) gives every read's text length, a scan the offsets, the write pass the dense SAM text -- which is all that crosses
// PCIe back to the host (430 B per 150 bp read) instead of the candidates + CIGARs and 2 us of host CPU per read.
// --------------------------------------------------------------------------------------------
struct SamDev {
	int n;
	const int64_t *off;
	const int32_t *len;
	const char *seq, *qual; // qual may be NULL
	const char *names;
	const int64_t *name_off;
	const int64_t *coff;         // candidates of read i: cand[coff[i] .. coff[i+1])
	const gd_sr_cand_t *cand;
	uint32_t *pool;              // CIGAR pool the cand[].cigar_off index (edited in place)
	const uint8_t *qbuf, *tbuf;  // code strings of candidate c at c * stride
	int64_t stride;
	gdsam::RefNames ref;
	gd_sr_post_opt_t post;
};
GD_DEV gdsam::ReadIn sam_read_in(const SamDev &D, int i)
{
	gdsam::ReadIn R;
	const int64_t c0 = D.coff[i];
	R.name = D.names + D.name_off[i], R.seq = D.seq + D.off[i], R.qual = D.qual ? D.qual + D.off[i] : nullptr;
	R.qlen = D.len[i], R.n_cand = (int)(D.coff[i + 1] - c0);
	R.cand = D.cand + c0, R.cigar = D.pool;
	R.qcodes = D.qbuf + c0 * D.stride, R.tcodes = D.tbuf + c0 * D.stride, R.stride = D.stride;
	return R;
}
__global__ void __launch_bounds__(128) gd_sam_count_kernel(SamDev D, uint32_t *text_len)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= D.n) return;
	const gdsam::ReadIn R = sam_read_in(D, i);
	gdsam::Sink s = {nullptr, 0};
	gdsam::one_read(R, D.post, D.ref, s);
	text_len[i] = (uint32_t)s.n;
}
__global__ void __launch_bounds__(128) gd_sam_write_kernel(SamDev D, const int64_t *text_off, char *text)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= D.n) return;
	const gdsam::ReadIn R = sam_read_in(D, i);
	gdsam::Sink s = {text + text_off[i], 0};
	gdsam::one_read(R, D.post, D.ref, s);
}

// What a call in SAM mode (gd_sr_map_sam_batch) carries through its slices.
struct SamJob {
	const char *const *names; // of the call's reads (host)
	const char *qual;         // same layout as the reads' buffer, or NULL
	gd_sr_post_opt_t post;
	std::string rblob;        // contig names, NUL separated
	std::vector<int32_t> rnoff;
	int gen;                  // which of the two pinned text buffers of every lane this call fills
	struct Piece {
		int lane;
		size_t off, len;
	};
	std::vector<Piece> pieces; // by slice index; resolved to pointers when the call ends (a lane buffer may have grown)
};

// --------------------------------------------------------------------------------------------
// host side
// --------------------------------------------------------------------------------------------
template <class T> static int scan_u32(gd_ctx *ctx, const uint32_t *d_in, T *d_out, int64_t n)
{ // exclusive prefix sums with the total in d_out[n]; the input has n+1 readable entries (the last one is ignored)
	size_t tmp = 0;
	GD_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, tmp, d_in, d_out, n + 1, ctx->stream));
	int rc = gd_reserve(ctx, ctx->mp_tmp, tmp + 16);
	if (rc) return rc;
	GD_CUDA_OK(ctx, cub::DeviceScan::ExclusiveSum(ctx->mp_tmp.p, tmp, d_in, d_out, n + 1, ctx->stream));
	ctx->stat_launches += 2;
	return GD_OK;
}

// GD_MAP_PROFILE=1: host-side wall time of every phase of a slice (each ends with the stream synchronised)
struct SrPhaseClock {
	bool on;
	cudaStream_t s;
	double t0;
	std::vector<std::pair<const char *, double>> marks;
	static double now()
	{
		struct timespec ts;
		clock_gettime(CLOCK_MONOTONIC, &ts);
		return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
	}
	explicit SrPhaseClock(cudaStream_t st) : on(getenv("GD_MAP_PROFILE") != nullptr), s(st), t0(0)
	{
		if (on) t0 = now();
	}
	void mark(const char *what)
	{
		if (!on) return;
		cudaStreamSynchronize(s);
		const double t = now();
		marks.push_back({what, t - t0});
		t0 = t;
	}
	~SrPhaseClock()
	{
		if (!on) return;
		fprintf(stderr, "[gd_sr_map slice]");
		for (auto &m : marks) fprintf(stderr, " %s %.2f ms |", m.first, m.second);
		fprintf(stderr, "\n");
	}
};

// The slices of one call run on several lanes (the caller's context and a chain of peer contexts on helper threads, one stream
// each) so that the transfers, the 8-byte read-backs and the small kernels of one slice overlap the DP kernel of the others.  Results
// land in the caller's arrays in input order: a slice claims its output ranges when its sizes are known, in slice order.
struct SliceOrder {
	std::mutex mu;
	std::condition_variable cv;
	int next = 0;
	bool failed = false;
	int64_t cand_base = 0, cig_base = 0;
	bool claim(int slice, int64_t nc, int64_t ncig, int64_t &cb, int64_t &gb)
	{
		std::unique_lock<std::mutex> lk(mu);
		cv.wait(lk, [&] { return failed || next == slice; });
		if (failed) return false;
		cb = cand_base, gb = cig_base;
		cand_base += nc, cig_base += ncig, ++next;
		cv.notify_all();
		return true;
	}
	void fail()
	{
		std::lock_guard<std::mutex> lk(mu);
		failed = true;
		cv.notify_all();
	}
};
struct SliceClaim { // a slice that returns before claiming (an error) must not leave the other lane waiting
	SliceOrder &ord;
	bool done = false;
	explicit SliceClaim(SliceOrder &o) : ord(o) {}
	~SliceClaim()
	{
		if (!done) ord.fail();
	}
};

// SAM mode: the slice's reads -> SAM text in the lane's pinned buffer (sam->pieces[slice_index]); see K5
static int sam_stage(gd_ctx *ctx, int lane, int n, const int64_t *off, const int32_t *len, int64_t lo, int64_t hi, const int64_t *h_coff,
                     int64_t nc, int64_t ncig, int64_t stride, SamJob *sam, int first_read, int slice_index)
{
	cudaStream_t s = ctx->stream;
	int rc;
	// read names of the slice: one blob + offsets, through pinned staging
	size_t nbytes = 0;
	for (int i = 0; i < n; ++i) nbytes += strlen(sam->names[first_read + i]) + 1;
	const size_t off_bytes = (size_t)(n + 1) * 8;
	if ((rc = gd_reserve_pinned(ctx, ctx->h_names, off_bytes + nbytes + 64))) return rc;
	int64_t *h_noff = (int64_t *)ctx->h_names.p;
	char *h_blob = (char *)ctx->h_names.p + off_bytes;
	size_t at = 0;
	for (int i = 0; i < n; ++i) {
		const char *nm = sam->names[first_read + i];
		const size_t l = strlen(nm) + 1;
		h_noff[i] = (int64_t)at;
		memcpy(h_blob + at, nm, l), at += l;
	}
	h_noff[n] = (int64_t)at;
	if ((rc = gd_reserve(ctx, ctx->mp_names, off_bytes + nbytes + 64))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_names.p, ctx->h_names.p, off_bytes + nbytes, cudaMemcpyHostToDevice, s));
	if (sam->qual) {
		if ((rc = gd_reserve(ctx, ctx->mp_qual, (size_t)(hi - lo) + 16))) return rc;
		GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_qual.p, sam->qual + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s));
	}
	// contig names + the per-read candidate offsets (the device copy was reused for the CIGAR offsets)
	const size_t rn_bytes = sam->rnoff.size() * 4, rb_bytes = sam->rblob.size();
	if ((rc = gd_reserve(ctx, ctx->mp_rnames, rn_bytes + rb_bytes + 64))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_rnames.p, sam->rnoff.data(), rn_bytes, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync((char *)ctx->mp_rnames.p + rn_bytes, sam->rblob.data(), rb_bytes, cudaMemcpyHostToDevice, s));
	if ((rc = gd_reserve(ctx, ctx->mp_rcoff, off_bytes))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_rcoff.p, h_coff, off_bytes, cudaMemcpyHostToDevice, s));
	if ((rc = gd_reserve(ctx, ctx->mp_cpool2, (size_t)(ncig + 1) * 4))) return rc;
	if (ncig) GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_cpool2.p, ctx->mp_cpool.p, (size_t)ncig * 4, cudaMemcpyDeviceToDevice, s));
	if ((rc = gd_reserve(ctx, ctx->mp_slen, (size_t)(n + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_soff, (size_t)(n + 2) * 8))) return rc;
	SamDev D;
	D.n = n, D.off = (const int64_t *)ctx->mp_off.p, D.len = (const int32_t *)ctx->mp_len.p, D.seq = (const char *)ctx->mp_seq.p;
	D.qual = sam->qual ? (const char *)ctx->mp_qual.p : nullptr;
	D.name_off = (const int64_t *)ctx->mp_names.p, D.names = (const char *)ctx->mp_names.p + off_bytes;
	D.coff = (const int64_t *)ctx->mp_rcoff.p, D.cand = (const gd_sr_cand_t *)ctx->mp_cand.p;
	D.pool = (uint32_t *)ctx->mp_cpool2.p; // the count pass edits the copy
	D.qbuf = (const uint8_t *)ctx->mp_qbuf.p, D.tbuf = (const uint8_t *)ctx->mp_tbuf.p, D.stride = stride;
	D.ref.off = (const int32_t *)ctx->mp_rnames.p, D.ref.blob = (const char *)ctx->mp_rnames.p + rn_bytes;
	D.post = sam->post;
	(void)nc;
	const int blocks = (n + 127) / 128;
	gd_sam_count_kernel<<<blocks, 128, 0, s>>>(D, (uint32_t *)ctx->mp_slen.p);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	if ((rc = scan_u32(ctx, (const uint32_t *)ctx->mp_slen.p, (int64_t *)ctx->mp_soff.p, n))) return rc;
	int64_t *h_word = (int64_t *)ctx->h_mp.p;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_word, (int64_t *)ctx->mp_soff.p + n, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const size_t total = (size_t)h_word[0];
	if ((rc = gd_reserve(ctx, ctx->mp_text, total + 16))) return rc;
	D.pool = (uint32_t *)ctx->mp_cpool.p;
	gd_sam_write_kernel<<<blocks, 128, 0, s>>>(D, (const int64_t *)ctx->mp_soff.p, (char *)ctx->mp_text.p);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	// the lane's pinned text buffer of this call's generation: append (grown geometrically, earlier pieces move with it)
	GdPinned &hb = ctx->h_sam[sam->gen];
	size_t &used = ctx->h_sam_used[sam->gen];
	if (used + total + 1 > hb.cap) {
		GdPinned nb;
		const size_t want = std::max<size_t>((used + total) * 2, (size_t)64 << 20);
		GD_CUDA_OK(ctx, cudaHostAlloc(&nb.p, want, cudaHostAllocPortable));
		nb.cap = want;
		if (used) memcpy(nb.p, hb.p, used);
		if (hb.p) cudaFreeHost(hb.p);
		hb = nb;
	}
	if (total) GD_CUDA_OK(ctx, cudaMemcpyAsync((char *)hb.p + used, ctx->mp_text.p, total, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	sam->pieces[slice_index] = {lane, used, total};
	used += total;
	return GD_OK;
}

static int sr_map_slice(gd_ctx *ctx, const gd_index *idx, int n, const int64_t *off, const int32_t *len, const char *buf,
                        const gd_sr_opt_t *o, const gd_lr_opt_t *lr, SliceOrder &ord, int slice_index, int64_t *cand_off, gd_sr_cand_t *cand,
                        int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap, SamJob *sam = nullptr, int lane = 0, int first_read = 0)
{
	SliceClaim claim(ord);
	int64_t cand_base = 0, cig_base = 0;
	cudaStream_t s = ctx->stream;
	SrPhaseClock clk(s);
	int rc;
	int64_t lo = INT64_MAX, hi = 0, sum_len = 0;
	int max_len = 0, min_len = INT32_MAX;
	for (int i = 0; i < n; ++i) {
		min_len = std::min(min_len, len[i]);
		if (len[i] < o->W || off[i] < 0) {
			ctx->err = "gd_sr_map_batch: read shorter than the pattern";
			return GD_ERR_ARG;
		}
		lo = std::min<int64_t>(lo, off[i]), hi = std::max<int64_t>(hi, off[i] + len[i]);
		max_len = std::max(max_len, len[i]), sum_len += len[i];
	}
	SrParams P;
	memset(&P, 0, sizeof(P));
	P.W = o->W, P.k = idx->d.k, P.frag_mode = o->frag_mode;
	P.max_nb_seeds = o->frag_mode ? (o->max_frag_len == 0 ? 800u : (uint32_t)o->max_frag_len) : 0xffffffffu; // map.c:621-622
	P.bw = o->bw, P.bw_frac = lr ? 0.f : o->bw_frac, P.bw_min = lr ? 0u : o->bw_min, P.bw_max = lr ? 0u : o->bw_max;
	P.min_cnt = o->min_cnt, P.rec_frac = o->rec_threshold_frac, P.q_occ_frac = o->q_occ_frac;
	P.af_max_loc = o->af_max_loc, P.mid_occ = o->mid_occ, P.max_max_occ = o->max_max_occ, P.occ_dist = o->occ_dist;
	P.for_only = o->for_only, P.rev_only = o->rev_only, P.a = o->a;
	P.stride = (max_len + 15) / 16 * 16;
	const int per_read = lr ? (int)lr->vt_nb_loc + 2 : o->af_max_loc; // candidate slots per read
	if (lr) P.vt_dis = lr->vt_dis, P.vt_nb_loc = lr->vt_nb_loc, P.max_max_gap = lr->max_max_gap, P.max_min_gap = lr->max_min_gap,
		P.vt_cov = lr->vt_cov, P.vt_df1 = lr->vt_df1, P.vt_df2 = lr->vt_df2, P.vt_f = lr->vt_f;
	// ---- reads up
	if ((rc = gd_reserve(ctx, ctx->mp_seq, (size_t)(hi - lo) + 16))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_off, (size_t)n * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_len, (size_t)n * 4))) return rc;
	if ((rc = gd_reserve_pinned(ctx, ctx->h_mp, (size_t)(2 * n + 2) * 8 + 64))) return rc;
	int64_t *h_off = (int64_t *)ctx->h_mp.p;
	int64_t *h_coff = h_off + n + 4; // this slice's candidate offsets, relative, until its output range is claimed
	for (int i = 0; i < n; ++i) h_off[i] = off[i] - lo;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_seq.p, buf + lo, (size_t)(hi - lo), cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_off.p, h_off, (size_t)n * 8, cudaMemcpyHostToDevice, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(ctx->mp_len.p, len, (size_t)n * 4, cudaMemcpyHostToDevice, s));
	clk.mark("scan+upload");
	const int64_t *d_off = (const int64_t *)ctx->mp_off.p;
	const int32_t *d_len = (const int32_t *)ctx->mp_len.p;
	const char *d_buf = (const char *)ctx->mp_seq.p;
	// ---- sketches of every shift (mm_sketch2 + mm_sketch3)
	GdReadSketch K;
	if ((rc = gd_sketch_reads_device_raw(ctx, n, d_off, d_len, d_buf, max_len, sum_len, idx->d.w, idx->d.k, o->Z, o->W, o->max_seeds,
	                                     P.max_nb_seeds == 0xffffffffu ? 0u : P.max_nb_seeds, &K)))
		return rc;
	P.JW = K.JW, P.crop = K.crop;
	clk.mark("sketch");
	// ---- K1
	if ((rc = gd_reserve(ctx, ctx->mp_seed_n, (size_t)K.raw_cap * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_seed_first, (size_t)K.raw_cap * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_state, (size_t)n * sizeof(SrRead)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_cnt, (size_t)(n + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_hoff, (size_t)(n + 2) * 8))) return rc;
	const int blocks = std::max(1, std::min((n + SR_WARPS - 1) / SR_WARPS, ctx->sms * 16));
	P.cnt_table = (lr || max_len / o->W > o->mid_occ) ? 1 : 0; // only reads that can have more than mid_occ seeds need it
	gd_sr_seed_kernel<<<blocks, SR_WARPS * 32, P.cnt_table ? SR_WARPS * 4096 : 0, s>>>(idx->d, P, n, d_len, K.job_off, K.raw, K.c3, K.c2, K.ret3,
	                                                 (uint32_t *)ctx->mp_seed_n.p, (uint32_t *)ctx->mp_seed_first.p,
	                                                 (SrRead *)ctx->mp_state.p, (uint32_t *)ctx->mp_cnt.p);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	if ((rc = scan_u32(ctx, (const uint32_t *)ctx->mp_cnt.p, (int64_t *)ctx->mp_hoff.p, n))) return rc;
	int64_t *h_word = (int64_t *)ctx->h_mp.p; // (h_off has been consumed by the copy above once the stream reaches here)
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_word, (int64_t *)ctx->mp_hoff.p + n, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const int64_t n_hits = h_word[0];
	clk.mark("seed");
	// ---- K2
	if ((rc = gd_reserve(ctx, ctx->mp_ht, (size_t)(n_hits + 1) * 8))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_hq, (size_t)(n_hits + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_cand_tmp, (size_t)n * per_read * sizeof(gd_sr_cand_t)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_ncand, (size_t)(n + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_coff, (size_t)(n + 4) * 8))) return rc;
	int *d_span = (int *)((int64_t *)ctx->mp_coff.p + n + 1); // widest query / target of any candidate (long reads)
	if (lr) {
		GD_CUDA_OK(ctx, cudaMemsetAsync(d_span, 0, 4, s));
		gd_lr_vote_kernel<<<blocks, SR_WARPS * 32, 0, s>>>(idx->d, P, n, d_len, K.job_off, K.raw, (const uint32_t *)ctx->mp_seed_n.p,
		                                                 (const uint32_t *)ctx->mp_seed_first.p, (const SrRead *)ctx->mp_state.p,
		                                                 (const int64_t *)ctx->mp_hoff.p, (uint64_t *)ctx->mp_ht.p, (uint32_t *)ctx->mp_hq.p,
		                                                 (gd_sr_cand_t *)ctx->mp_cand_tmp.p, (uint32_t *)ctx->mp_ncand.p, d_span);
	} else
		gd_sr_vote_kernel<<<blocks, SR_WARPS * 32, 0, s>>>(idx->d, P, n, d_len, K.job_off, K.raw, (const uint32_t *)ctx->mp_seed_n.p,
		                                                 (const uint32_t *)ctx->mp_seed_first.p, (const SrRead *)ctx->mp_state.p,
		                                                 (const int64_t *)ctx->mp_hoff.p, (uint64_t *)ctx->mp_ht.p, (uint32_t *)ctx->mp_hq.p,
		                                                 (gd_sr_cand_t *)ctx->mp_cand_tmp.p, (uint32_t *)ctx->mp_ncand.p);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	if ((rc = scan_u32(ctx, (const uint32_t *)ctx->mp_ncand.p, (int64_t *)ctx->mp_coff.p, n))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_word, (int64_t *)ctx->mp_coff.p + n, 16, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_coff, ctx->mp_coff.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const int64_t nc = h_word[0];
	int max_q = max_len, max_t = max_len; // upper bounds of the DP shapes
	if (lr) {
		const int span = (int)(h_word[1] & 0xffffffff);
		max_q = std::min(max_len, std::max(span, 1)), max_t = std::max(span, 1);
		P.stride = (std::max(span, 1) + 15) / 16 * 16;
	}
	clk.mark("vote");
	if (nc == 0) {
		if (sam) { // every read of the slice is unmapped: flag-4 records
			claim.done = true;
			return sam_stage(ctx, lane, n, off, len, lo, hi, h_coff, 0, 0, P.stride, sam, first_read, slice_index);
		}
		if (!ord.claim(slice_index, 0, 0, cand_base, cig_base)) return SR_ABORTED;
		claim.done = true;
		for (int i = 0; i <= n; ++i) cand_off[i] = cand_base;
		return GD_OK;
	}
	// ---- K3
	if ((rc = gd_reserve(ctx, ctx->mp_cand, (size_t)nc * sizeof(gd_sr_cand_t)))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_qbuf, (size_t)nc * P.stride + 256))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_tbuf, (size_t)nc * P.stride + 256))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_cnt, (size_t)(std::max<int64_t>(nc, n) + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_hoff, (size_t)(std::max<int64_t>(nc, n) + 2) * 8))) return rc;
	gd_sr_cand_t *d_cand = (gd_sr_cand_t *)ctx->mp_cand.p;
	gd_sr_compact_kernel<<<(n + 127) / 128, 128, 0, s>>>(n, per_read, (const gd_sr_cand_t *)ctx->mp_cand_tmp.p,
	                                                   (const int64_t *)ctx->mp_coff.p, d_cand);
	const int wblocks = (int)std::max<int64_t>(1, std::min<int64_t>((nc + 3) / 4, (int64_t)ctx->sms * 16));
	gd_sr_window_kernel<<<wblocks, 128, 0, s>>>(idx->d, P, nc, d_off, d_len, d_buf, d_cand, (uint8_t *)ctx->mp_qbuf.p,
	                                          (uint8_t *)ctx->mp_tbuf.p, (uint32_t *)ctx->mp_cnt.p);
	ctx->stat_launches += 2;
	GD_CUDA_OK(ctx, cudaGetLastError());
	int64_t *d_pair_off = (int64_t *)ctx->mp_hoff.p;
	if ((rc = scan_u32(ctx, (const uint32_t *)ctx->mp_cnt.p, d_pair_off, nc))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_word, d_pair_off + nc, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const int64_t np = h_word[0];
	clk.mark("window");
	// ---- DP on the candidates that are not exact matches (flag KSW_EZ_APPROX_MAX, map.c:867)
	const int cig_stride = (max_q + max_t + 8 + 31) & ~31; // rows of the DP's CIGAR scratch start on 128-byte lines
	if (np > 0) {
		if (np > 0x7fffffff) {
			ctx->err = "gd_sr_map_batch: too many DP pairs in one slice";
			return GD_ERR_ARG;
		}
		if ((rc = gd_reserve(ctx, ctx->mp_pair, (size_t)np * 20 + 64))) return rc;
		if ((rc = gd_reserve(ctx, ctx->mp_ez, (size_t)np * sizeof(gd_extz_t)))) return rc;
		if ((rc = gd_reserve(ctx, ctx->mp_cig, (size_t)np * cig_stride * 4 + 64))) return rc;
		int64_t *d_poff = (int64_t *)ctx->mp_pair.p;
		int32_t *d_pqlen = (int32_t *)(d_poff + np), *d_ptlen = d_pqlen + np;
		int32_t *d_pw = P.bw_max > 0 ? d_ptlen + np : nullptr; // per-pair band when it depends on the read length
		const int max_bw = (int)std::max(sr_read_bw(P, max_len), sr_read_bw(P, min_len));
		gd_sr_pairs_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(nc, P.stride, d_cand, d_pair_off, d_pqlen, d_ptlen, d_poff, d_pw, P, d_len);
		ctx->stat_launches++;
		int8_t mat[25]; // map.c:861-865
		const int g = o->a, bb = o->b < 0 ? o->b : -o->b;
		for (int x = 0; x < 5; ++x)
			for (int y = 0; y < 5; ++y) mat[x * 5 + y] = (x == 4 || y == 4) ? 0 : (x == y ? g : bb);
		gd_ksw_params_t prm = {5, mat, o->q, o->e, o->q2, o->e2, o->zdrop, o->end_bonus, 0x08};
		if ((rc = gd_ksw_run_device(ctx, (int)np, d_pqlen, d_poff, (const uint8_t *)ctx->mp_qbuf.p, d_ptlen, d_poff,
		                            (const uint8_t *)ctx->mp_tbuf.p, d_pw, (int)o->bw, max_q, max_t, max_bw, &prm,
		                            (gd_extz_t *)ctx->mp_ez.p, (uint32_t *)ctx->mp_cig.p, cig_stride)))
			return rc;
	}
	clk.mark("dp");
	// ---- K4
	if ((rc = gd_reserve(ctx, ctx->mp_ncand, (size_t)(nc + 1) * 4))) return rc;
	if ((rc = gd_reserve(ctx, ctx->mp_coff, (size_t)(std::max<int64_t>(nc, n) + 2) * 8))) return rc;
	gd_sr_scores_kernel<<<(unsigned)((nc + 255) / 256), 256, 0, s>>>(nc, P, d_len, d_cand, d_pair_off, (const gd_extz_t *)ctx->mp_ez.p,
	                                                               (uint32_t *)ctx->mp_ncand.p);
	ctx->stat_launches++;
	int64_t *d_cig_off = (int64_t *)ctx->mp_coff.p;
	if ((rc = scan_u32(ctx, (const uint32_t *)ctx->mp_ncand.p, d_cig_off, nc))) return rc;
	GD_CUDA_OK(ctx, cudaMemcpyAsync(h_word, d_cig_off + nc, 8, cudaMemcpyDeviceToHost, s));
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	const int64_t ncig = h_word[0];
	clk.mark("scores");
	if (sam) { // SAM mode: nothing of this slice depends on the others; the text is produced on the device
		if (ncig > 0x7fffffff) {
			ctx->err = "gd_sr_map_sam_batch: CIGAR pool of one slice exceeds 2^31 entries";
			return GD_ERR_ARG;
		}
		if ((rc = gd_reserve(ctx, ctx->mp_cpool, (size_t)(ncig + 1) * 4))) return rc;
		gd_sr_cigars_kernel<<<wblocks, 128, 0, s>>>(nc, d_cand, d_pair_off, d_cig_off, (const uint32_t *)ctx->mp_cig.p, cig_stride,
		                                          (uint32_t *)ctx->mp_cpool.p, ncig);
		ctx->stat_launches++;
		GD_CUDA_OK(ctx, cudaGetLastError());
		claim.done = true;
		rc = sam_stage(ctx, lane, n, off, len, lo, hi, h_coff, nc, ncig, P.stride, sam, first_read, slice_index);
		clk.mark("sam");
		return rc;
	}
	if (!ord.claim(slice_index, nc, ncig, cand_base, cig_base)) return SR_ABORTED;
	claim.done = true;
	for (int i = 0; i <= n; ++i) cand_off[i] = h_coff[i] + cand_base; // (the shared boundary entry gets the same value from both neighbours)
	const bool fits = cand_base + nc <= cand_cap && cig_base + ncig <= cigar_cap && cand && cigar;
	if ((rc = gd_reserve(ctx, ctx->mp_cpool, (size_t)(ncig + 1) * 4))) return rc;
	gd_sr_cigars_kernel<<<wblocks, 128, 0, s>>>(nc, d_cand, d_pair_off, d_cig_off, (const uint32_t *)ctx->mp_cig.p, cig_stride,
	                                          (uint32_t *)ctx->mp_cpool.p, ncig);
	ctx->stat_launches++;
	GD_CUDA_OK(ctx, cudaGetLastError());
	if (fits) {
		GD_CUDA_OK(ctx, cudaMemcpyAsync(cand + cand_base, d_cand, (size_t)nc * sizeof(gd_sr_cand_t), cudaMemcpyDeviceToHost, s));
		if (ncig) GD_CUDA_OK(ctx, cudaMemcpyAsync(cigar + cig_base, ctx->mp_cpool.p, (size_t)ncig * 4, cudaMemcpyDeviceToHost, s));
	}
	GD_CUDA_OK(ctx, cudaStreamSynchronize(s));
	clk.mark("cigars+download");
	if (fits && cig_base)
		for (int64_t c = 0; c < nc; ++c) cand[cand_base + c].cigar_off += (int32_t)cig_base;
	return GD_OK;
}

extern "C" int gd_init(int device, gd_ctx **ctx);

// lane L of a context: the context itself, its peer, the peer's peer ... (created on first use, destroyed with the context)
static gd_ctx *lane_ctx(gd_ctx *ctx, int L)
{
	while (L-- > 0 && ctx) ctx = ctx->peer;
	return ctx;
}
static int map_lanes(gd_ctx *ctx, int64_t want)
{
	int have = 1;
	for (gd_ctx *c = ctx; have < want; ++have, c = c->peer)
		if (!c->peer && gd_init(ctx->device, &c->peer) != GD_OK) break; // no further context: fewer lanes
	return have;
}

static int run_slices(gd_ctx *ctx, const gd_index *idx, const std::vector<std::pair<int, int>> &slices, const int64_t *off,
                      const int32_t *len, const char *buf, const gd_sr_opt_t *o, const gd_lr_opt_t *lr, int64_t *cand_off,
                      gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cand, int64_t *n_cig,
                      SamJob *sam = nullptr)
{
	SliceOrder ord;
	ctx->err.clear();
	const int ns = (int)slices.size();
	// long reads: two lanes at most (every lane owns a backtrack arena of tens of GB)
	const int lanes = map_lanes(ctx, std::min<int64_t>(ns, lr ? std::min<long>(ctx->opt_map_lanes, 2) : ctx->opt_map_lanes));
	std::vector<int> rcs((size_t)lanes, GD_OK);
	auto lane = [&](int L) {
		gd_ctx *c = lane_ctx(ctx, L);
		cudaSetDevice(c->device);
		for (int k = L; k < ns; k += lanes) {
			const int b = slices[k].first, m = slices[k].second;
			const int rc = sr_map_slice(c, idx, m, off + b, len + b, buf, o, lr, ord, k, cand_off ? cand_off + b : nullptr, cand, cand_cap, cigar,
			                            cigar_cap, sam, L, b);
			if (rc) {
				rcs[L] = rc;
				ord.fail();
				return;
			}
		}
	};
	{
		std::vector<std::thread> helpers;
		for (int L = 1; L < lanes; ++L) helpers.emplace_back(lane, L);
		lane(0);
		for (std::thread &t : helpers) t.join();
	}
	*n_cand = ord.cand_base, *n_cig = ord.cig_base;
	// the code (and message) of the lane that actually failed, not of one that merely bailed out after it
	bool any = false;
	for (int L = 0; L < lanes; ++L) {
		any |= rcs[L] != GD_OK;
		if (rcs[L] && rcs[L] != SR_ABORTED) {
			if (L > 0) ctx->err = lane_ctx(ctx, L)->err;
			return rcs[L];
		}
	}
	return any ? GD_ERR_CUDA : GD_OK;
}

extern "C" int gd_sr_map_batch(gd_ctx *ctx, const gd_index *idx, int n, const int64_t *off, const int32_t *len, const char *buf,
                               const gd_sr_opt_t *o, int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar,
                               int64_t cigar_cap, int64_t *n_cigar)
{
	if (!ctx) return GD_ERR_ARG;
	if (!idx || !o || n < 0 || !cand_off || (n > 0 && (!off || !len || !buf))) {
		ctx->err = "gd_sr_map_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (o->af_max_loc < 1 || o->af_max_loc > SR_MAX_LOC || o->W < 2 || o->W > 63 || idx->device != ctx->device) {
		// (W = 1, `-Z 1`: the reference's mm_sketch2 then walks one shift more than the pattern has, sketch.c:2143-2225, and maps a
		// shifted copy of the read; that path is not pinned against the reference, so it is refused rather than answered differently)
		ctx->err = "gd_sr_map_batch: need 1 <= af_max_loc <= 32, 2 <= W <= 63 and an index built on this device";
		return GD_ERR_ARG;
	}
	cand_off[0] = 0;
	if (n_cigar) *n_cigar = 0;
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	// slices: at most 2^18 reads, at least two per lane when the batch is large enough to share between the lanes
	const int per_call = 2 * (int)std::max<long>(2, std::min<long>(ctx->opt_map_lanes, 8));
	const int slice = std::max(32768, std::min(1 << 18, (n + per_call - 1) / per_call));
	std::vector<std::pair<int, int>> slices;
	for (int b = 0; b < n; b += slice) slices.push_back({b, std::min(slice, n - b)});
	int64_t cand_base = 0, cig_base = 0;
	{
		int rc = run_slices(ctx, idx, slices, off, len, buf, o, nullptr, cand_off, cand, cand_cap, cigar, cigar_cap, &cand_base, &cig_base);
		if (rc) return rc;
	}
	if (n_cigar) *n_cigar = cig_base;
	if (cig_base > 0x7fffffff) {
		ctx->err = "gd_sr_map_batch: CIGAR pool of one call exceeds 2^31 entries; map fewer reads per call";
		return GD_ERR_ARG;
	}
	if (cand_base > cand_cap || cig_base > cigar_cap || (cand_base && !cand) || (cig_base && !cigar)) {
		ctx->err = "gd_sr_map_batch: output buffer too small";
		return GD_ERR_CAPACITY;
	}
	return GD_OK;
}

// Reads in, SAM records out: gd_sr_map_batch with the post-DP stage on the device (K5).  The text comes back as pieces in
// input order (one per slice) that live in pinned buffers OWNED BY THE CONTEXT: they stay valid until the call after the
// next one on this context (two generations alternate, so a host can write batch i out while batch i+1 is mapped); the
// caller frees only the two arrays, with gd_free.
extern "C" int gd_sr_map_sam_batch(gd_ctx *ctx, const gd_index *idx, int n, const char *const *names, const int64_t *off, const int32_t *len,
                                   const char *seq, const char *qual, const gd_sr_opt_t *o, const gd_sr_post_opt_t *post, int n_seq,
                                   const char *const *seq_names, char ***parts, size_t **part_len, int *n_parts)
{
	if (!ctx) return GD_ERR_ARG;
	if (!idx || !o || !post || n < 0 || !parts || !part_len || !n_parts || n_seq < 1 || !seq_names || (n > 0 && (!names || !off || !len || !seq))) {
		ctx->err = "gd_sr_map_sam_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (o->af_max_loc < 1 || o->af_max_loc > SR_MAX_LOC || o->W < 2 || o->W > 63 || idx->device != ctx->device) {
		ctx->err = "gd_sr_map_sam_batch: need 1 <= af_max_loc <= 32, 2 <= W <= 63 and an index built on this device";
		return GD_ERR_ARG;
	}
	if (!post->is_sr) {
		ctx->err = "gd_sr_map_sam_batch: the device SAM stage implements the linear gap cost of MM_F_SR (short reads)";
		return GD_ERR_ARG;
	}
	*parts = nullptr, *part_len = nullptr, *n_parts = 0;
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	SamJob job;
	job.names = names, job.qual = qual, job.post = *post;
	for (int i = 0; i < n_seq; ++i) {
		job.rnoff.push_back((int32_t)job.rblob.size());
		job.rblob += seq_names[i], job.rblob.push_back('\0');
	}
	const bool prof = getenv("GD_MAP_PROFILE") != nullptr;
	const double tp0 = prof ? SrPhaseClock::now() : 0;
	job.gen = (int)(ctx->sam_calls++ & 1);
	const int per_call = 2 * (int)std::max<long>(2, std::min<long>(ctx->opt_map_lanes, 8)); // two slices per lane, as in gd_sr_map_batch
	const int slice = std::max(32768, std::min(1 << 18, (n + per_call - 1) / per_call));
	const int n_slices = (n + slice - 1) / slice;
	const int lanes = map_lanes(ctx, std::min<int64_t>(n_slices, ctx->opt_map_lanes)); // the lanes exist before their text buffers are sized
	for (gd_ctx *c = ctx; c; c = c->peer) c->h_sam_used[job.gen] = 0;
	{ // size this generation's pinned text buffers once, from an estimate of the text (2 x bases + ~200 B per read, a lane's share
	  // of it per lane): page-locking memory is slow, and a buffer that grows is page-locked again and again and copied each time
		size_t est = 0;
		for (int i = 0; i < n; ++i) est += 2 * (size_t)len[i] + 200;
		est = est / (size_t)lanes + est / 8 + ((size_t)1 << 20);
		for (int L = 0; L < lanes; ++L) {
			gd_ctx *c = lane_ctx(ctx, L);
			if (!c || c->h_sam[job.gen].cap >= est) continue;
			if (c->h_sam[job.gen].p) cudaFreeHost(c->h_sam[job.gen].p);
			c->h_sam[job.gen].p = nullptr, c->h_sam[job.gen].cap = 0;
			if (cudaHostAlloc(&c->h_sam[job.gen].p, est, cudaHostAllocPortable) == cudaSuccess) c->h_sam[job.gen].cap = est;
			else cudaGetLastError(), c->h_sam[job.gen].p = nullptr; // (sam_stage grows it on demand)
		}
	}
	std::vector<std::pair<int, int>> slices;
	for (int b = 0; b < n; b += slice) slices.push_back({b, std::min(slice, n - b)});
	job.pieces.assign(slices.size(), SamJob::Piece{0, 0, 0});
	int64_t cb = 0, gb = 0;
	const double tp1 = prof ? SrPhaseClock::now() : 0;
	int rc = run_slices(ctx, idx, slices, off, len, seq, o, nullptr, nullptr, nullptr, 0, nullptr, 0, &cb, &gb, &job);
	if (prof) fprintf(stderr, "[gd_sr_map_sam_batch] %d reads, %zu slices: prologue %.2f ms, slices %.2f ms\n", n, slices.size(), tp1 - tp0, SrPhaseClock::now() - tp1);
	if (rc) return rc;
	const size_t np = job.pieces.size();
	*parts = (char **)malloc((np + 1) * sizeof(char *)), *part_len = (size_t *)malloc((np + 1) * sizeof(size_t));
	if (!*parts || !*part_len) return GD_ERR_ARG;
	for (size_t k = 0; k < np; ++k) {
		gd_ctx *c = lane_ctx(ctx, job.pieces[k].lane);
		(*parts)[k] = (char *)c->h_sam[job.gen].p + job.pieces[k].off, (*part_len)[k] = job.pieces[k].len;
	}
	*n_parts = (int)np;
	return GD_OK;
}

// The lanes of the mapping stage and the pinned text buffers of both generations of gd_sr_map_sam_batch, ahead of the first call
// (a one-shot host program does this beside its index build: creating contexts and page-locking memory is 0.1-0.3 s).
// text_bytes: what one call is expected to write (SAM text of one mini-batch).
extern "C" int gd_sr_map_sam_prepare(gd_ctx *ctx, size_t text_bytes)
{
	if (!ctx) return GD_ERR_ARG;
	cudaSetDevice(ctx->device);
	const int lanes = map_lanes(ctx, ctx->opt_map_lanes);
	const size_t est = text_bytes / (size_t)lanes + text_bytes / 8 + ((size_t)1 << 20);
	for (int L = 0; L < lanes; ++L) {
		gd_ctx *c = lane_ctx(ctx, L);
		for (int gen = 0; gen < 2; ++gen) {
			if (c->h_sam[gen].cap >= est) continue;
			if (c->h_sam[gen].p) cudaFreeHost(c->h_sam[gen].p);
			c->h_sam[gen].p = nullptr, c->h_sam[gen].cap = 0;
			if (cudaHostAlloc(&c->h_sam[gen].p, est, cudaHostAllocPortable) == cudaSuccess) c->h_sam[gen].cap = est;
			else cudaGetLastError(), c->h_sam[gen].p = nullptr; // (the stage grows it on demand)
		}
	}
	return GD_OK;
}

extern "C" int gd_lr_map_batch(gd_ctx *ctx, const gd_index *idx, int n, const int64_t *off, const int32_t *len, const char *buf,
                               const gd_lr_opt_t *lr, int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar,
                               int64_t cigar_cap, int64_t *n_cigar)
{
	if (!ctx) return GD_ERR_ARG;
	if (!idx || !lr || n < 0 || !cand_off || (n > 0 && (!off || !len || !buf))) {
		ctx->err = "gd_lr_map_batch: bad argument";
		return GD_ERR_ARG;
	}
	if (lr->vt_nb_loc < 1 || lr->vt_nb_loc > SR_MAX_LOC || lr->W < 2 || lr->W > 63 || idx->device != ctx->device) {
		ctx->err = "gd_lr_map_batch: need 1 <= vt_nb_loc <= 32, 2 <= W <= 63 (pattern length 1 is not pinned: refused) and an index built on this device";
		return GD_ERR_ARG;
	}
	gd_sr_opt_t o; // the fields both trees share
	memset(&o, 0, sizeof(o));
	o.W = lr->W, memcpy(o.Z, lr->Z, sizeof(o.Z)), o.max_seeds = lr->max_seeds, o.frag_mode = lr->frag_mode, o.max_frag_len = lr->max_frag_len;
	o.bw = lr->bw, o.af_max_loc = (int32_t)lr->vt_nb_loc + 2, o.mid_occ = lr->mid_occ, o.max_max_occ = lr->max_max_occ, o.occ_dist = lr->occ_dist;
	o.q_occ_frac = lr->q_occ_frac, o.for_only = lr->for_only, o.rev_only = lr->rev_only;
	o.a = lr->a, o.b = lr->b, o.q = lr->q, o.e = lr->e, o.q2 = lr->q2, o.e2 = lr->e2, o.zdrop = lr->zdrop, o.end_bonus = lr->end_bonus;
	cand_off[0] = 0;
	if (n_cigar) *n_cigar = 0;
	if (n == 0) return GD_OK;
	cudaSetDevice(ctx->device);
	int64_t cand_base = 0, cig_base = 0;
	std::vector<std::pair<int, int>> slices;
	{ // slices of at most 64 Mbases (the sketch lists and hit arrays scale with the bases).  Long reads are NOT cut finer to
	  // feed the two lanes: their time is the DP, and the DP wants every pair of the batch in one launch (cutting 500 ONT
	  // reads into four slices doubled the time of the stage).
		const int64_t lim = 64ll << 20;
		for (int b = 0; b < n;) {
			int m = 0;
			int64_t bases = 0;
			while (b + m < n && m < (1 << 18) && (m == 0 || bases + len[b + m] <= lim)) bases += len[b + m], ++m;
			slices.push_back({b, m});
			b += m;
		}
		// ... except when one slice is all there is and it is large: its DP would run as several arena-limited launches one
		// after the other on one lane (15 kbp reads, band 1000: ~60 MB of backtrack per pair, ~1,000 pairs per launch = 7 warps
		// per SM); two halves on the two lanes keep twice as many pairs in flight
		if (slices.size() == 1 && n >= 2048) {
			int64_t total = 0, half = 0;
			for (int i = 0; i < n; ++i) total += len[i];
			if (total >= (32ll << 20)) {
				int m = 0;
				while (m < n - 1 && half + len[m] <= total / 2) half += len[m], ++m;
				if (m > 0) slices[0] = {0, m}, slices.push_back({m, n - m});
			}
		}
	}
	{
		int rc = run_slices(ctx, idx, slices, off, len, buf, &o, lr, cand_off, cand, cand_cap, cigar, cigar_cap, &cand_base, &cig_base);
		if (rc) return rc;
	}
	if (n_cigar) *n_cigar = cig_base;
	if (cig_base > 0x7fffffff) {
		ctx->err = "gd_lr_map_batch: CIGAR pool of one call exceeds 2^31 entries; map fewer reads per call";
		return GD_ERR_ARG;
	}
	if (cand_base > cand_cap || cig_base > cigar_cap || (cand_base && !cand) || (cig_base && !cigar)) {
		ctx->err = "gd_lr_map_batch: output buffer too small";
		return GD_ERR_CAPACITY;
	}
	return GD_OK;
}
