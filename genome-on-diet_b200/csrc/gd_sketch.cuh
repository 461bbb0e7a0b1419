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
	int32_t dl;       // sparsified length, sk_diet_len(len, shift) -- filled by the job kernels (read by the v3 tile body)
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
	// Fixed-stride output (one tile per job only): job j writes at most fixed_stride records at out[j * fixed_stride ..]
	// and its count to out_cnt[j].  No tile depends on another one then: no ticket, no look-back.
	int64_t fixed_stride;      // 0 = dense output in job order (decoupled look-back)
	int32_t *out_cnt;          // [njobs], fixed-stride mode
	// v3 tile body only:
	int32_t pack_jobs;         // > 0 (fixed-stride mode): a tile holds this many whole jobs (<= 32), one N slot between them
};

GD_DEV uint64_t sk_hash64(uint64_t key, uint64_t mask)
{ // GDiet-ShortReads/sketch.c:25-34.  The shift-and-add stages are multiplications by constants modulo 2^64
  // (~key + (key << 21) = key * (2^21 - 1) - 1; key + (key << 3) + (key << 8) = key * 265; ... * 21; ... * (2^31 + 1)): IMADs on
  // the FMA pipe instead of shifts and 64-bit adds on the integer pipe.  Measured on a B200 (round 2): no change in kernel time
  // (136.5 vs 133 Gbases/s) -- and neither did comparing the 50-bit records as doubles on the FP64 pipe (sm_100a has no DMNMX:
  // DSETP + 2 selects, 133.6 Gbases/s; reverted).  The kernel waits on its five block-wide phases, not on an ALU pipe.
	key = (key * 0x1fffffull - 1ull) & mask;
	key = key ^ key >> 24;
	key = (key * 265ull) & mask;
	key = key ^ key >> 14;
	key = (key * 21ull) & mask;
	key = key ^ key >> 28;
	key = (key * 0x80000001ull) & mask;
	return key;
}

GD_DEV int sk_nt4(unsigned c)
{ // seq_nt4_table, GDiet-ShortReads/sketch.c:11-18, branch-free: A/a C/c G/g T/t U/u -> 0 1 2 3 3, bytes 0..3 map to
  // themselves, everything else -> 4.  For the letters ((c>>1)^(c>>2))&3 is the code; they are the bytes 0x40..0x7f
  // whose low five bits are 1, 3, 7, 20 or 21.
	const unsigned letter = ((c & 0xc0u) == 0x40u) & ((0x0030008au >> (c & 31u)) & 1u);
	const unsigned code = ((c >> 1) ^ (c >> 2)) & 3u;
	return c < 4 ? (int)c : letter ? (int)code : 4;
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

// ---- tile geometry ----
// A tile loads NP = THREADS*8 consecutive sparsified positions [B0, B0+NP), 8 per thread, and emits for the
// TP positions that follow a left halo of HL = 2w+k-3 positions (w-1 for the windows that contain an emit
// position + w+k-2 for the N-free run that decides whether such a window is full) and precede a right
// halo of w-1 positions.
GD_HD int sk_halo_left(int w, int k) { return 2 * w + k - 3; }
GD_HD int sk_tile_emit(int np, int w, int k) { return np - sk_halo_left(w, k) - (w - 1); }

#define GD_SK_PADW 2 // zero words in front of the forward-packed codes (>= k bases)
#define GD_SK_RAW(THREADS) ((THREADS) * 24) // bytes of original sequence a tile can stage: three per loaded position ("10", "110", "100", "101001" ...); sparser patterns read global memory directly

// Shared memory of one block.  SUF / PREM are indexed [p][t] (position 8t+p at p*THREADS+t) so that
// the threads of a warp touch consecutive 8-byte words.
template <int THREADS> struct SketchSmem {
	enum { NP = THREADS * 8 };
	uint64_t SUF[NP];  // w >= 9: minimum of X over [s, end of s's 8-position chunk]; w <= 8: X itself
	uint64_t PREM[NP]; // w >= 9: maximum of M over [start of chunk, s];               w <= 8: M itself
	uint32_t F2[NP / 16 + GD_SK_PADW + 2]; // 2-bit codes, position s at bit 2s (after GD_SK_PADW zero words)
	uint32_t R2[NP / 16 + 4];              // 2-bit codes, position s at bit 2(NP-1-s); zero words behind
	int32_t warp_val[32];
	int32_t warp_cnt[32];
	long long excl;
	int32_t tile;
	uint8_t ones_loc[64];      // copy of SketchParams::ones_loc (indexed per lane)
	uint32_t raw[GD_SK_RAW(THREADS) / 4 + 2]; // the tile's slice of the ASCII sequence, staged with coalesced word loads
};

GD_DEV uint64_t sk_bits64(const uint32_t *wds, int bit)
{ // 64 bits of a little-endian bit stream starting at bit offset `bit` (3 words are read)
	const int wi = bit >> 5, sh = bit & 31;
	const uint64_t lo = (uint64_t)wds[wi] | (uint64_t)wds[wi + 1] << 32;
	return sh ? (lo >> sh) | ((uint64_t)wds[wi + 2] << (64 - sh)) : lo;
}
GD_DEV uint64_t sk_min64(uint64_t a, uint64_t b) { return a < b ? a : b; }
GD_DEV uint64_t sk_max64(uint64_t a, uint64_t b) { return a > b ? a : b; }

template <int THREADS>
GD_DEV void sketch_tile_body(const SketchParams &S, const SketchBatch &B, SketchSmem<THREADS> *sm)
{
	const int NP = THREADS * 8;
	const int tid = thread_idx(), lane = tid & 31, wid = tid >> 5;
	const int w = S.w, k = S.k, full_run = w + k - 1;
	const int HL = sk_halo_left(w, k);
	uint16_t *const F2h = (uint16_t *)(sm->F2 + GD_SK_PADW), *const R2h = (uint16_t *)sm->R2;
	if (tid < GD_SK_PADW) sm->F2[tid] = 0;
	if (tid < 2) sm->F2[NP / 16 + GD_SK_PADW + tid] = 0;
	if (tid < 4) sm->R2[NP / 16 + tid] = 0;
	for (int i = tid; i < 64; i += THREADS) sm->ones_loc[i] = S.ones_loc[i];
	// Tiles are handed out in order by a global ticket, so every tile's predecessors are held by blocks that are already
	// running: the look-back below can never wait for a block that is not resident.  (Drawing several consecutive tiles
	// per atomic was tried for the one-warp tiles of short reads and dropped: the first tile of a batch then waits for the
	// LAST tile of the previous block's batch, which serialises the blocks.)
	const bool fixed = B.fixed_stride > 0;
	for (long long round = 0;; ++round) {
		if (tid == 0) sm->tile = fixed ? (int32_t)(block_idx() + round * grid_dim()) : atomic_add(B.ticket, 1);
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
		const long long i0 = chunk * S.TP; // first emit position of the tile
		const long long B0 = i0 - HL;      // sparsified position held in slot 0
		const int s0 = tid * 8;
		// A job whose sparsified sequence is shorter than one full window run (w+k-1) cannot emit anything -- the cropped
		// shift-0 job of mm_sketch2 on a 150 bp read is 15 bases long -- so its tile goes straight to the (empty) output scan.
		uint32_t real[8], zbits = 0, emit = 0;
		uint64_t X[8];
		int cnt = 0;
		if (dl >= full_run) {
		// ---- phase 0: stage the original bytes the tile touches, [real(first position), real(last position)] ----
		const long long jlo = B0 > 0 ? B0 : 0, jhi = (B0 + NP < dl ? B0 + NP : dl) - 1; // loaded positions inside the sequence
		long long raw_lo = 0; // offset (relative to seq, may be -1..-3) of the first staged byte
		bool staged = false;
		if (jhi >= jlo) {
			const uint32_t rlo = sk_real((uint32_t)jlo, shift, S), rhi = sk_real((uint32_t)jhi, shift, S);
			// word-aligned window of the global buffer; the last, possibly partial word is fetched bytewise
			const unsigned long long a0 = (unsigned long long)(seq + rlo);
			const uint32_t lead = (uint32_t)(a0 & 3);
			const uint32_t nbytes = rhi - rlo + 1 + lead;
			if (nbytes <= GD_SK_RAW(THREADS)) {
				staged = true, raw_lo = (long long)rlo - lead;
				const uint32_t *gsrc = (const uint32_t *)(seq + raw_lo);
				const uint32_t nfull = nbytes >> 2;
				for (uint32_t i = tid; i < nfull; i += THREADS) sm->raw[i] = gsrc[i];
				if (tid < (int)(nbytes & 3)) ((uint8_t *)sm->raw)[nfull * 4 + tid] = (uint8_t)seq[raw_lo + nfull * 4 + tid];
			}
		}
		sync_block();
		// ---- phase 1: encode 8 positions per thread; positions outside [0,dl) count as N ----
		uint32_t code = 0, nmask = 0;
		{
			const uint32_t raw_lo0 = jhi >= jlo ? sk_real((uint32_t)jlo, shift, S) : 0; // some byte of the sequence that is staged / exists
			const uint8_t *src = staged ? (const uint8_t *)sm->raw - raw_lo : jhi >= jlo ? (const uint8_t *)seq : (const uint8_t *)sm->raw;
			const long long j0 = B0 + s0;
			// real(j) = (j/ones)*W + ones_loc[j%ones] + shift (get_real_location, sketch.c:20-23), stepped incrementally
			uint32_t qd = 0, rm = 0;
			if (j0 > 0) qd = (uint32_t)j0 / (uint32_t)S.ones, rm = (uint32_t)j0 - qd * (uint32_t)S.ones;
			uint32_t base = qd * (uint32_t)S.W + shift;
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const long long j = j0 + p;
				const bool in = j >= 0 && j < dl;
				const uint32_t rp = base + sm->ones_loc[rm];
				const int c = sk_nt4(src[in ? rp : (uint32_t)raw_lo0]); // out-of-range positions read a valid byte, count as N
				real[p] = rp;
				if (in && c < 4) code |= (uint32_t)c << (2 * p);
				else nmask |= 1u << p;
				if (j >= 0 && ++rm == (uint32_t)S.ones) rm = 0, base += (uint32_t)S.W;
			}
			F2h[tid] = (uint16_t)code;
			uint32_t rc = 0; // the same 8 codes in reverse position order
#pragma unroll
			for (int p = 0; p < 8; ++p) rc |= ((code >> (2 * p)) & 3u) << (2 * (7 - p));
			R2h[THREADS - 1 - tid] = (uint16_t)rc;
		}
		// ---- last N position before this thread's chunk: block-wide inclusive max-scan, made exclusive ----
		int lastn;
		{
			int v = nmask ? s0 + 31 - clz32(nmask) : -1; // position 'before the tile' counts as N
			int inc = v;
			for (int d = 1; d < 32; d <<= 1) {
				int o = (int)shfl_up(0xffffffffu, (uint32_t)inc, d, 32);
				if (lane >= d && o > inc) inc = o;
			}
			if (lane == 31) sm->warp_val[wid] = inc;
			sync_block(); // also publishes F2 / R2
			int prev = (int)shfl_up(0xffffffffu, (uint32_t)inc, 1, 32);
			lastn = lane == 0 ? -1 : prev;
			for (int i = 0; i < wid; ++i) lastn = sm->warp_val[i] > lastn ? sm->warp_val[i] : lastn;
		}
		// ---- phase 2: the k-mers ending just before the chunk come out of the packed arrays; roll over the 8
		// positions, hash (sketch.c:1660-1683) ----

		uint32_t fullbits = 0;
		{
			uint64_t rv = (~sk_bits64(sm->F2, 32 * GD_SK_PADW + 2 * (s0 - k))) & S.mask; // bases s0-k .. s0-1, complemented
			uint64_t fw = sk_bits64(sm->R2, 2 * (NP - s0)) & S.mask;                       // bases s0-1 .. s0-k
			// (bases before slot 0 read as zero bits; such k-mers are never valid because slot -1 counts as N)
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const int sp = s0 + p;
				const uint64_t c = (code >> (2 * p)) & 3u;
				fw = (fw << 2 | c) & S.mask;
				rv = (rv >> 2) | (3ull ^ c) << (2 * (k - 1));
				if (nmask >> p & 1) lastn = sp;
				const int run = sp - lastn;
				uint64_t x = GD_SK_MAXU64;
				if (run >= k && fw != rv) {
					const int z = fw < rv ? 0 : 1;
					x = sk_hash64(z ? rv : fw, S.mask) << 8 | (uint64_t)k;
					zbits |= (uint32_t)z << p;
				}
				if (run >= full_run) fullbits |= 1u << p;
				X[p] = x;
			}
		}


		if (w >= 9) {
			// ---- phase 3: minimum of every full window ending at e = s0+p.  The window starts in an earlier
			// chunk (w-1 >= 8): own prefix minimum, whole chunks in between, suffix minimum of the first chunk ----
			uint64_t pre[8], M[8];
			{
				uint64_t suf = GD_SK_MAXU64, pr = GD_SK_MAXU64;
#pragma unroll
				for (int p = 7; p >= 0; --p) suf = sk_min64(suf, X[p]), sm->SUF[p * THREADS + tid] = suf;
#pragma unroll
				for (int p = 0; p < 8; ++p) pr = sk_min64(pr, X[p]), pre[p] = pr;
			}
			sync_block();
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const int a = s0 + p - (w - 1); // first position of the window
				uint64_t m = 0;                 // 0 = "no full window ends here" (X >= 1 always)
				if ((fullbits >> p & 1) && a >= 0) {
					const int ta = a >> 3;
					m = sk_min64(pre[p], sm->SUF[(a & 7) * THREADS + ta]);
#pragma unroll 1
					for (int c = ta + 1; c < tid; ++c) m = sk_min64(m, sm->SUF[c]); // SUF[0][c] = minimum of chunk c
					if (m == GD_SK_MAXU64) m = 0;
				}
				M[p] = m;
			}
			// ---- phase 4: X(i) is emitted iff it equals the largest full-window minimum among the windows that
			// contain i, i.e. the maximum of M over [i, i+w-1]: own suffix maximum, whole chunks, prefix maximum ----
			uint64_t sufm[8];
			{
				uint64_t pm = 0, sx = 0;
#pragma unroll
				for (int p = 0; p < 8; ++p) pm = sk_max64(pm, M[p]), sm->PREM[p * THREADS + tid] = pm;
#pragma unroll
				for (int p = 7; p >= 0; --p) sx = sk_max64(sx, M[p]), sufm[p] = sx;
			}
			sync_block();
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const int sp = s0 + p;
				if (sp >= HL && sp < HL + S.TP && B0 + sp < dl && X[p] != GD_SK_MAXU64) {
					const int b = sp + w - 1, tb = b >> 3; // last position of the last window; b < NP in the emit range
					uint64_t mx = sk_max64(sufm[p], sm->PREM[(b & 7) * THREADS + tb]);
#pragma unroll 1
					for (int c = tid + 1; c < tb; ++c) mx = sk_max64(mx, sm->PREM[7 * THREADS + c]); // PREM[7][c] = maximum of chunk c
					if (mx == X[p]) emit |= 1u << p, ++cnt;
				}
			}
		} else {
			// ---- small windows (w <= 8): direct scans over X and M in shared memory ----
#pragma unroll
			for (int p = 0; p < 8; ++p) sm->SUF[p * THREADS + tid] = X[p];
			sync_block();
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const int sp = s0 + p;
				uint64_t m = 0;
				if ((fullbits >> p & 1) && sp >= w - 1) {
					m = GD_SK_MAXU64;
#pragma unroll 1
					for (int d = 0; d < w; ++d) m = sk_min64(m, sm->SUF[((sp - d) & 7) * THREADS + ((sp - d) >> 3)]);
					if (m == GD_SK_MAXU64) m = 0;
				}
				sm->PREM[p * THREADS + tid] = m;
			}
			sync_block();
#pragma unroll
			for (int p = 0; p < 8; ++p) {
				const int sp = s0 + p;
				if (sp >= HL && sp < HL + S.TP && B0 + sp < dl && X[p] != GD_SK_MAXU64) {
					uint64_t mx = 0;
#pragma unroll 1
					for (int d = 0; d < w; ++d) mx = sk_max64(mx, sm->PREM[((sp + d) & 7) * THREADS + ((sp + d) >> 3)]);
					if (mx == X[p]) emit |= 1u << p, ++cnt;
				}
			}
		}
		} // dl >= full_run
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
		// ---- decoupled look-back over tiles, 32 predecessors per probe (warp 0); fixed-stride mode: the job's own slot ----
		if (fixed) {
			if (tid == 0) {
				sm->excl = (long long)job * B.fixed_stride;
				B.out_off[job] = sm->excl, B.out_cnt[job] = total;
				if (tile == B.ntiles - 1) B.out_off[B.njobs] = (long long)B.njobs * B.fixed_stride;
			}
		} else if (wid == 0) {
			long long excl = 0;
			if (tile == 0) {
				if (lane == 0) st_volatile(&B.status[0], (2ull << 62) | (unsigned long long)total);
			} else {
				if (lane == 0) {
					st_volatile(&B.status[tile], (1ull << 62) | (unsigned long long)total);
					fence();
				}
				long long pt = tile - 1;
				for (;;) {
					const long long idx = pt - lane;
					unsigned long long sv = 2ull << 62; // "before the first tile": an inclusive prefix of 0
					if (idx >= 0) do sv = ld_volatile(&B.status[idx]);
						while ((sv >> 62) == 0);
					const uint32_t incl = ballot(0xffffffffu, (sv >> 62) == 2); // lanes that hold an inclusive prefix
					const int first = incl ? ffs32(incl) - 1 : 32;                // the nearest one ends the walk
					unsigned long long c = lane <= first ? (sv & 0x3fffffffffffffffull) : 0ull;
					for (int d = 16; d >= 1; d >>= 1) {
						const uint32_t lo = shfl_xor(0xffffffffu, (uint32_t)c, d, 32), hi = shfl_xor(0xffffffffu, (uint32_t)(c >> 32), d, 32);
						c += (unsigned long long)hi << 32 | lo;
					}
					excl += (long long)c;
					if (incl) break;
					pt -= 32;
				}
				if (lane == 0) st_volatile(&B.status[tile], (2ull << 62) | (unsigned long long)(excl + total));
			}
			if (lane == 0) {
				sm->excl = excl;
				if (chunk == 0) B.out_off[job] = excl;
				if (tile == B.ntiles - 1) B.out_off[B.njobs] = excl + total;
			}
		}
		sync_block();
		const long long obase = sm->excl + local;
		int o = 0;
#pragma unroll
		for (int p = 0; p < 8; ++p)
			if (emit >> p & 1) {
				const long long dst = obase + o;
				if (fixed ? (local + o < B.fixed_stride) : (dst < B.out_cap)) {
					const uint64_t y = (uint64_t)J.rid << 32 | (uint64_t)real[p] << 1 | (uint64_t)(zbits >> p & 1);
					B.out[2 * dst] = X[p];
					B.out[2 * dst + 1] = y;
				}
				++o;
			}
		sync_block(); // shared memory is reused by the next tile
	}
}


// =============================================================================================
// Tile body, design v3 (round 2): the same tiles, phases and output order as sketch_tile_body above -- about a third of
// its instructions.  What changed, phase by phase (SASS counts of the 256-thread kernel in DESIGN.md 4):
//   * tile header (ticket, tile -> job by a cached job or a 32-ary search, sparsified length from the job record, staged
//     byte range) is made by warp 0 alone, right behind its share of the previous tile, and broadcast through shared memory;
//   * dense output: a tile publishes its count and parks its records in shared memory; the decoupled look-back and the
//     coalesced copy-out run inside the block's next tile, so no barrier waits for a spinning warp;
//   * short reads: PACKED tiles hold several whole jobs, one N slot between them, each with its own fixed-stride output;
//   * encode: one shared-memory table look-up per base (code | N flag << 16), accumulated with a shift-add; the pattern
//     walk has a stride-W form for patterns with a single '1' ("10", "100"); threads whose 8 positions lie inside the
//     sequence skip the range tests;
//   * N-free run lengths come from a bitmap of N flags in shared memory (two 32-bit probes for w+k-1 <= 32) instead of a
//     block-wide max-scan;
//   * both k-mers of all 8 positions are funnel-shifted out of three registers of the forward / reversed 2-bit stream
//     (no rolling dependency chain); hash64 runs on explicit 32-bit halves (the high half of a 2k <= 56 bit key has
//     <= 24 bits, so two of the three xor-shifts touch only the low word; 2k <= 32 is all 32-bit);
//   * window minima / emission maxima run on key = hash | 1 << 56 (0 = "no full window", ~0 = no k-mer); the chunk
//     geometry of a window (which 8-position chunk it starts in, how many whole chunks lie between) is the same for
//     every thread, so the addresses are `thread * 8 + uniform offset` and the in-between loop has a uniform trip count.
// =============================================================================================
template <int THREADS> struct SketchSmem3 {
	enum { NP = THREADS * 8, F2W = NP / 16 + GD_SK_PADW + 4, R2W = NP / 16 + 6,
	       PARK = THREADS >= 64 ? NP / 4 : 8 }; // records the parking area holds (one-warp tiles are used with fixed-stride output, which parks nothing)
	uint64_t SUF[NP];  // w >= 9: minimum of the keys over [s, end of s's chunk]; w <= 8: the key itself   ([p][thread])
	uint64_t PREM[NP]; // w >= 9: maximum of M over [start of chunk, s];          w <= 8: M itself          ([p][thread])
	uint32_t F2[F2W];  // 2-bit codes, position s at bit 2s (after GD_SK_PADW zero words)
	uint32_t R2[R2W];  // 2-bit codes, position s at bit 2(NP-1-s); zero words behind
	uint16_t NB[NP / 8 + 4]; // N flags: position s at bit 2(s & 7) of halfword 2 + s / 8 (two zero halfwords in front)
	uint32_t lut[256]; // seq_nt4_table (sketch.c:11-18) as code, or 1 << 16 for N
	int32_t warp_cnt[32];
	long long excl;
	// tile headers, written by warp 0 while the block finishes the previous tile (two copies, used alternately)
	struct Hdr {
		long long tile, seq_off;
		int32_t job, i0, dl, staged, raw_lo, nbytes, safe;
		uint32_t shift, rid;
	} hdr[2];
	// the job of the block's previous tile: its tile range and record (most tiles of a contig stay in the same job)
	long long c_lo, c_hi;
	// packed tiles: the jobs of the tile -- first slot, slot behind the last position + its N, sparsified length, records in
	// lower slots
	long long pk_seq[32];
	int32_t pk_base[33], pk_end[32], pk_dl[32], pk_R[33], pk_n;
	uint32_t pk_shift[32], pk_rid[32];
	int32_t c_job;
	SketchJob c_J;
	uint8_t ones_loc[64];
	alignas(16) unsigned long long park[2 * PARK]; // dense output: the records of the block's previous tile until their offset is known
	alignas(16) uint32_t raw[GD_SK_RAW(THREADS) / 4 + 4]; // the tile's slice of the ASCII sequence (16-byte aligned in global memory)
};

GD_DEV uint32_t sk_mask_from(int a) { return a <= 0 ? 0xffu : a >= 8 ? 0u : (0xffu << a) & 0xffu; } // bits p >= a of 8

// hash64 (sketch.c:25-34) of a key of 32 < 2k <= 56 bits held as (hi < 2^24, lo); mhi = mask >> 32
GD_DEV void sk_hash_hl(uint32_t &lo, uint32_t &hi, const uint32_t mhi)
{
	uint64_t t = (uint64_t)lo * 0x1fffffu + 0xffffffffffffffffull; // key * (2^21 - 1) - 1 = ~key + (key << 21)
	hi = ((uint32_t)(t >> 32) + hi * 0x1fffffu) & mhi, lo = (uint32_t)t;
	lo ^= funnel_r(lo, hi, 24); // key ^= key >> 24 (hi >> 24 == 0)
	t = (uint64_t)lo * 265u;    // key + (key << 3) + (key << 8)
	hi = ((uint32_t)(t >> 32) + hi * 265u) & mhi, lo = (uint32_t)t;
	lo ^= funnel_r(lo, hi, 14), hi ^= hi >> 14;
	t = (uint64_t)lo * 21u; // key + (key << 2) + (key << 4)
	hi = ((uint32_t)(t >> 32) + hi * 21u) & mhi, lo = (uint32_t)t;
	lo ^= funnel_r(lo, hi, 28); // hi >> 28 == 0
	t = (uint64_t)lo * 0x80000001u; // key + (key << 31)
	hi = ((uint32_t)(t >> 32) + hi * 0x80000001u) & mhi, lo = (uint32_t)t;
}
GD_DEV uint32_t sk_hash_32(uint32_t key, const uint32_t m)
{ // the same for 2k <= 32
	key = (key * 0x1fffffu - 1u) & m;
	key ^= key >> 24;
	key = (key * 265u) & m;
	key ^= key >> 14;
	key = (key * 21u) & m;
	key ^= key >> 28;
	key = (key * 0x80000001u) & m;
	return key;
}

template <int THREADS, bool PACKED>
GD_DEV void sketch_tile_body3(const SketchParams &S, const SketchBatch &B, SketchSmem3<THREADS> *sm)
{
	const int NP = THREADS * 8;
	const int tid = thread_idx(), lane = tid & 31, wid = tid >> 5;
	const int w = S.w, k = S.k, full_run = w + k - 1, wm1 = w - 1;
	const int HL = sk_halo_left(w, k);
	const bool ones1 = S.ones == 1;
	const bool k32 = 2 * k <= 32;
	const uint32_t mlo = (uint32_t)S.mask, mhi = (uint32_t)(S.mask >> 32);
	uint16_t *const F2h = (uint16_t *)(sm->F2 + GD_SK_PADW), *const R2h = (uint16_t *)sm->R2;
	for (int i = tid; i < SketchSmem3<THREADS>::F2W; i += THREADS) sm->F2[i] = 0;
	for (int i = tid; i < SketchSmem3<THREADS>::R2W; i += THREADS) sm->R2[i] = 0;
	if (tid < 4) sm->NB[tid < 2 ? tid : NP / 8 + tid] = 0;
	if (tid == 0) { // with constant indices the kernel parameter stays in the constant bank (a run-time index makes a local copy of S)
		const uint32_t *src = (const uint32_t *)S.ones_loc;
		uint32_t *dst = (uint32_t *)sm->ones_loc;
#pragma unroll
		for (int i = 0; i < 16; ++i) dst[i] = src[i];
	}
	for (int c = tid; c < 256; c += THREADS) sm->lut[c] = (uint32_t)sk_nt4((unsigned)c) < 4u ? (uint32_t)sk_nt4((unsigned)c) : 0x10000u;
	if (tid == 0) sm->c_lo = 0, sm->c_hi = 0, sm->c_job = -1;
	const bool fixed = B.fixed_stride > 0;
	// Tiles are handed out in order by a global ticket, drawn when the block is about to start the tile: every tile's
	// predecessors are then held by blocks that are already computing them.  (Drawing the ticket one tile ahead was measured
	// on a B200 and is twice as slow: a block that is late holds a low ticket hostage and every later tile spins on it in the
	// look-back -- 850 probes per tile; drawing it just one phase ahead, with the look-back deferred: 192 against 198 Gbases/s.  A static tile = block + round * grid assignment of the fixed-stride mode left a third
	// of the resident warps idle: with grid a multiple of the jobs per read, the same blocks got all the 15-base cropped jobs.)
	// The header of a tile -- ticket, tile -> job, byte range to stage -- is made by warp 0 right after it parked the records
	// of the previous tile, so the barrier that ends a tile also publishes the next header.
	auto make_header = [&](typename SketchSmem3<THREADS>::Hdr *h) { // warp 0
		long long tile = 0;
		if (lane == 0) tile = (long long)atomic_add(B.ticket, 1);
		tile = (long long)((uint64_t)shfl_idx(0xffffffffu, (uint32_t)tile, 0, 32) |
		                   (uint64_t)shfl_idx(0xffffffffu, (uint32_t)((uint64_t)tile >> 32), 0, 32) << 32);
		if (lane == 0) h->tile = tile;
		if (tile >= B.ntiles) return;
		if (PACKED) { // jobs tile * G .. : lane g takes job g; a job that cannot emit gets no slots
			const long long jfirst = tile * B.pack_jobs;
			const int n = (int)(B.njobs - jfirst < (long long)B.pack_jobs ? B.njobs - jfirst : (long long)B.pack_jobs);
			SketchJob J;
			J.seq_off = 0, J.len = 0, J.shift = 0, J.rid = 0, J.dl = 0;
			if (lane < n) J = B.jobs[jfirst + lane];
			const int seg = lane < n && J.dl >= full_run ? J.dl + 1 : 0;
			int inc = seg;
			for (int d = 1; d < 32; d <<= 1) {
				const int o = (int)shfl_up(0xffffffffu, (uint32_t)inc, d, 32);
				if (lane >= d) inc += o;
			}
			if (lane < n) {
				sm->pk_seq[lane] = J.seq_off, sm->pk_base[lane] = inc - seg, sm->pk_end[lane] = inc, sm->pk_dl[lane] = seg ? J.dl : 0;
				sm->pk_shift[lane] = (uint32_t)J.shift, sm->pk_rid[lane] = J.rid;
				if (lane == n - 1) sm->pk_base[n] = inc, sm->pk_n = n;
			}
			if (lane == 0) h->job = (int)jfirst, h->i0 = 0, h->dl = full_run, h->shift = 0, h->rid = 0, h->seq_off = 0, h->staged = 0, h->raw_lo = 0, h->nbytes = 0, h->safe = 0;
			return;
		}
		int job;
		long long chunk;
		SketchJob J;
		if (S.one_tile_per_job) job = (int)tile, chunk = 0, J = B.jobs[job];
		else if (tile >= sm->c_lo && tile < sm->c_hi) job = sm->c_job, chunk = tile - sm->c_lo, J = sm->c_J;
		else { // last job with tile_base[job] <= tile: 32 split points per round trip
			long long lo = 0, hi = B.njobs, lo_base = 0;
			while (hi - lo > 1) {
				const long long step = (hi - lo + 31) >> 5, idx = lo + (lane + 1) * step;
				const long long v = idx < hi ? B.tile_base[idx] : 0;
				const uint32_t le = ballot(0xffffffffu, idx < hi && v <= tile);
				const int c = popc(le); // the predicate is monotone: the first c split points are <= tile
				if (c > 0)
					lo_base = (long long)((uint64_t)shfl_idx(0xffffffffu, (uint32_t)v, c - 1, 32) |
					                      (uint64_t)shfl_idx(0xffffffffu, (uint32_t)((uint64_t)v >> 32), c - 1, 32) << 32);
				if (lo + (c + 1) * step < hi) hi = lo + (c + 1) * step;
				lo += c * step;
			}
			job = (int)lo, chunk = tile - lo_base;
			J = B.jobs[job];
			const long long hi_base = B.tile_base[job + 1];
			sync_warp(0xffffffffu); // every lane has read the cache
			if (lane == 0) sm->c_lo = lo_base, sm->c_hi = hi_base, sm->c_job = job, sm->c_J = J;
			sync_warp(0xffffffffu);
		}
		if (lane == 0) {
			const uint32_t shift = (uint32_t)J.shift;
			const int dl = J.dl;
			const int i0 = (int)(chunk * S.TP), B0 = i0 - HL;
			h->job = job, h->i0 = i0, h->dl = dl, h->shift = shift, h->rid = J.rid, h->seq_off = J.seq_off;
			// bytes of the original sequence the tile touches: [real(first loaded position), real(last loaded position)]
			const int jlo = B0 > 0 ? B0 : 0, jhi = (B0 + NP < dl ? B0 + NP : dl) - 1;
			int staged = 0, raw_lo = 0, nbytes = 0, safe = 0;
			if (dl >= full_run && jhi >= jlo) {
				auto real_of = [&](uint32_t i) { // get_real_location, sketch.c:20-23
					const uint32_t qd = i / (uint32_t)S.ones, rm = i - qd * (uint32_t)S.ones;
					return qd * (uint32_t)S.W + sm->ones_loc[rm] + shift;
				};
				const uint32_t rlo = real_of((uint32_t)jlo), rhi = real_of((uint32_t)jhi);
				const uint32_t lead = (uint32_t)((unsigned long long)(B.buf + J.seq_off + rlo) & 15);
				safe = (int)rlo;
				nbytes = (int)(rhi - rlo + 1 + lead);
				if (nbytes <= GD_SK_RAW(THREADS)) staged = 1, raw_lo = (int)rlo - (int)lead;
			}
			h->staged = staged, h->raw_lo = raw_lo, h->nbytes = nbytes, h->safe = safe;
		}
	};
	// Dense output: the records of a tile are parked in shared memory and its count is published; the look-back over the
	// predecessors and the copy-out run in the middle of the block's NEXT tile, when the predecessors have long published
	// theirs: warp 0 issues the first probe when the tile starts and sums up after its hashing phase, the offset travels with
	// the barrier of the window-minimum phase, and nobody waits for warp 0 to spin.  (Measured on a B200, contigs: look-back
	// at the end of the own tile 172 Gbases/s -- 28 % of the stall samples at the barrier behind it --, deferred to the next
	// tile's hashing phase 211, deferred only to the next tile's encode phase 166: too early, 10 probes per tile.)
	// A tile with more records than the parking area (windows of w <= 6) waits for its offset and writes from registers.
	int pend = 0, pend_total = 0, pend_job = 0, pend_first = 0;
	long long pend_tile = 0;
	unsigned long long sv0 = 0; // warp 0: first look-back probe of the pending tile
	auto lookback_issue = [&]() { // warp 0
		const long long idx = pend_tile - 1 - lane;
		sv0 = idx >= 0 ? ld_volatile(&B.status[idx]) : 2ull << 62; // "before the first tile": an inclusive prefix of 0
	};
	auto lookback_finish = [&]() { // warp 0: decoupled look-back over tiles, 32 predecessors per probe
		long long excl = 0, pt = pend_tile - 1;
		unsigned long long sv = sv0;
		for (;;) {
			const long long idx = pt - lane;
			if (idx >= 0)
				while ((sv >> 62) == 0) sv = ld_volatile(&B.status[idx]);
			const uint32_t incl = ballot(0xffffffffu, (sv >> 62) == 2); // lanes that hold an inclusive prefix
			const int first = incl ? ffs32(incl) - 1 : 32;                // the nearest one ends the walk
			unsigned long long c = lane <= first ? (sv & 0x3fffffffffffffffull) : 0ull;
			for (int d = 16; d >= 1; d >>= 1) {
				const uint32_t lo = shfl_xor(0xffffffffu, (uint32_t)c, d, 32), hi = shfl_xor(0xffffffffu, (uint32_t)(c >> 32), d, 32);
				c += (unsigned long long)hi << 32 | lo;
			}
			excl += (long long)c;
			if (incl) break;
			pt -= 32;
			sv = pt - lane >= 0 ? ld_volatile(&B.status[pt - lane]) : 2ull << 62;
		}
		if (lane == 0) {
			st_volatile(&B.status[pend_tile], (2ull << 62) | (unsigned long long)(excl + pend_total));
			sm->excl = excl;
			if (pend_first) B.out_off[pend_job] = excl;
			if (pend_tile == B.ntiles - 1) B.out_off[B.njobs] = excl + pend_total;
		}
	};
	auto copy_out = [&]() { // every thread, after a barrier behind lookback_finish
		const long long excl = sm->excl;
		const ulonglong2 *st = (const ulonglong2 *)sm->park;
#pragma unroll 1
		for (int r = tid; r < pend_total; r += THREADS)
			if (excl + r < B.out_cap) {
				const ulonglong2 v = st[r];
				B.out[2 * (excl + r)] = v.x, B.out[2 * (excl + r) + 1] = v.y;
			}
	};
	sync_block();
	int cur = 0;
	for (;;) {
		if (PACKED && THREADS > 32) sync_block(); // the job table of the previous tile is still being read by the other warps
		if (wid == 0) make_header(&sm->hdr[cur]); // right behind warp 0's share of the previous tile's records
		sync_block(); // header of this tile; the previous tile's parked records
		const typename SketchSmem3<THREADS>::Hdr *const h = &sm->hdr[cur];
		const long long tile = h->tile;
		if (tile >= B.ntiles) break;
		const bool packed = PACKED; // several whole jobs per tile: slot 0 is the first position of the first job (no halo)
		const int job = h->job, dl = h->dl, i0 = h->i0, B0 = packed ? 0 : i0 - HL;
		const uint32_t shift = h->shift;
		const char *seq = B.buf + h->seq_off;
		const bool go = dl >= full_run; // a shorter job cannot emit (the 15-base cropped job of mm_sketch2 on a 150 bp read)
		if (pend && wid == 0) lookback_issue();
		const int s0 = tid * 8, j0 = B0 + s0;
		uint32_t zbits = 0, emit = 0, okbits = 0;
		uint64_t key[8];
		int cnt = 0;
		uint32_t nsp = 0;
		if (go) {
			// ---- phase 0: stage the bytes with 16-byte loads (the tail bytewise) ----
			const int staged = h->staged, raw_lo = h->raw_lo, safe = h->safe;
			if (staged) {
				const int nbytes = h->nbytes, nvec = nbytes >> 4;
				const uint4 *gsrc = (const uint4 *)(seq + raw_lo);
				uint4 *sdst = (uint4 *)sm->raw;
				for (int i = tid; i < nvec; i += THREADS) sdst[i] = gsrc[i];
				if (tid < (nbytes & 15)) ((uint8_t *)sm->raw)[nvec * 16 + tid] = (uint8_t)seq[raw_lo + nvec * 16 + tid];
				sync_block();
			}
			// ---- phase 1: 8 positions per thread -> 2-bit codes (bits 0..15) and N flags (bit 16 + 2p); positions outside
			// [0, dl) count as N.  real(j) = (j / ones) * W + ones_loc[j % ones] + shift (get_real_location, sketch.c:20-23) ----
			uint32_t acc = 0;
			{
				const bool all_in = j0 >= 0 && j0 + 7 < dl;
				uint32_t qd = 0, rm = 0;
				if (j0 > 0) {
					if (ones1) qd = (uint32_t)j0;
					else qd = (uint32_t)j0 / (uint32_t)S.ones, rm = (uint32_t)j0 - qd * (uint32_t)S.ones;
				}
				uint32_t base = qd * (uint32_t)S.W + shift;
				const uint32_t loc0 = sm->ones_loc[0];
				if (packed) { // the slots of job g are [pk_base[g], pk_end[g]): dl positions and one N; bytes straight from global memory
					const int n = sm->pk_n;
					int g = 0;
#pragma unroll
					for (int p = 0; p < 8; ++p) {
						const int sl = s0 + p;
#pragma unroll 1
						while (g < n && sl >= sm->pk_end[g]) ++g;
						uint32_t v = 0x10000u;
						if (g < n) {
							const uint32_t j = (uint32_t)(sl - sm->pk_base[g]);
							if ((int)j < sm->pk_dl[g]) {
								uint32_t real;
								if (ones1) real = j * (uint32_t)S.W + loc0 + sm->pk_shift[g];
								else {
									const uint32_t q = j / (uint32_t)S.ones;
									real = q * (uint32_t)S.W + sm->ones_loc[j - q * (uint32_t)S.ones] + sm->pk_shift[g];
								}
								v = sm->lut[(uint8_t)B.buf[sm->pk_seq[g] + real]];
							}
						}
						acc += v << (2 * p);
					}
				} else if (staged) {
					const uint8_t *rs = (const uint8_t *)sm->raw;
					if (all_in && ones1) {
						uint32_t idx = base + loc0 - (uint32_t)raw_lo;
#pragma unroll
						for (int p = 0; p < 8; ++p) acc += sm->lut[rs[idx]] << (2 * p), idx += (uint32_t)S.W;
					} else if (all_in) {
#pragma unroll
						for (int p = 0; p < 8; ++p) {
							acc += sm->lut[rs[base + sm->ones_loc[rm] - (uint32_t)raw_lo]] << (2 * p);
							if (++rm == (uint32_t)S.ones) rm = 0, base += (uint32_t)S.W;
						}
					} else {
#pragma unroll
						for (int p = 0; p < 8; ++p) {
							const int j = j0 + p;
							const bool in = j >= 0 && j < dl;
							const uint32_t rp = base + sm->ones_loc[rm];
							const uint32_t v = sm->lut[rs[(in ? rp : (uint32_t)safe) - (uint32_t)raw_lo]];
							acc += (in ? v : 0x10000u) << (2 * p);
							if (j >= 0 && ++rm == (uint32_t)S.ones) rm = 0, base += (uint32_t)S.W;
						}
					}
				} else { // sparse patterns (more than three original bytes per position): straight from global memory
					const uint8_t *gs = (const uint8_t *)seq;
#pragma unroll
					for (int p = 0; p < 8; ++p) {
						const int j = j0 + p;
						const bool in = j >= 0 && j < dl;
						const uint32_t rp = base + sm->ones_loc[rm];
						const uint32_t v = sm->lut[gs[in ? rp : (uint32_t)safe]];
						acc += (in ? v : 0x10000u) << (2 * p);
						if (j >= 0 && ++rm == (uint32_t)S.ones) rm = 0, base += (uint32_t)S.W;
					}
				}
			}
			const uint32_t code = acc & 0xffffu;
			nsp = acc >> 16;
			F2h[tid] = (uint16_t)code;
			{ // the same 8 codes in reverse position order: bit reversal, then the two bits of every code swapped back
				const uint32_t br = (uint32_t)(brev64((uint64_t)code) >> 48);
				R2h[THREADS - 1 - tid] = (uint16_t)(((br & 0x5555u) << 1) | ((br >> 1) & 0x5555u));
			}
			sm->NB[2 + tid] = (uint16_t)nsp;
		}
		sync_block(); // codes and N flags of the tile
		if (go) {
			// ---- last N position before the chunk (slot -1 counts as N); only runs up to w+k-1 matter ----
			int L0 = s0 - full_run - 1;
			if (L0 < -1) L0 = -1;
			{
				int hi = tid; // the 16 positions below 8 * hi are the halfwords hi, hi + 1
				for (int rem = full_run; rem > 0 && hi > 0; rem -= 16, hi -= 2) {
					const uint32_t bits = (uint32_t)sm->NB[hi] | (uint32_t)sm->NB[hi + 1] << 16;
					if (bits) {
						const int f = hi * 8 - 16 + ((31 - clz32(bits)) >> 1);
						if (f > L0) L0 = f;
						break;
					}
				}
			}
			// ---- k-mers that have no N (valid) and windows that are full, as 8-bit masks ----
			uint32_t validbits, fullbits;
			if (nsp == 0) {
				const int d = s0 - L0; // run length at p = 0
				validbits = sk_mask_from(k - d), fullbits = sk_mask_from(full_run - d);
			} else {
				validbits = fullbits = 0;
				int ln = L0;
#pragma unroll
				for (int p = 0; p < 8; ++p) {
					if (nsp >> (2 * p) & 1) ln = s0 + p;
					const int run = s0 + p - ln;
					if (run >= k) validbits |= 1u << p;
					if (run >= full_run) fullbits |= 1u << p;
				}
			}
			// ---- phase 2: both k-mers of every position out of the packed streams, canonical k-mer, hash (sketch.c:1660-1683).
			// rv(p): base s0+p-k+1 (complemented) at bit 0 = forward stream at bit 2(s0-k+1) + 2p; fw(p): base s0+p at bit 0 =
			// reversed stream at bit 2(NP-8-s0) + 2(7-p).  Bases before slot 0 read as zero bits (never valid: slot -1 is N).
			{
				const int bf = 32 * GD_SK_PADW + 2 * (s0 - k + 1), br = 2 * (NP - 8 - s0);
				const uint32_t *fp = sm->F2 + (bf >> 5), *rp = sm->R2 + (br >> 5);
				const uint32_t fs = (uint32_t)bf & 31u, rs = (uint32_t)br & 31u;
				const uint32_t a0 = fp[0], a1 = fp[1], a2 = fp[2], c0 = rp[0], c1 = rp[1], c2 = rp[2];
				const uint32_t f0 = funnel_r(a0, a1, fs), f1 = funnel_r(a1, a2, fs), g0 = funnel_r(c0, c1, rs), g1 = funnel_r(c1, c2, rs);
				if (k32) {
#pragma unroll
					for (int p = 0; p < 8; ++p) {
						const uint32_t rv = ~funnel_r(f0, f1, 2 * p) & mlo, fw = funnel_r(g0, g1, 2 * (7 - p)) & mlo;
						uint64_t x = GD_SK_MAXU64;
						if ((validbits >> p & 1) && fw != rv) {
							const uint32_t z = fw < rv ? 0u : 1u;
							x = (1ull << 56) | sk_hash_32(z ? rv : fw, mlo);
							zbits |= z << p, okbits |= 1u << p;
						}
						key[p] = x;
					}
				} else {
					const uint32_t a3 = fp[3], c3 = rp[3];
					const uint32_t f2 = funnel_r(a2, a3, fs), g2 = funnel_r(c2, c3, rs);
#pragma unroll
					for (int p = 0; p < 8; ++p) {
						const uint32_t rl = ~funnel_r(f0, f1, 2 * p), rh = ~funnel_r(f1, f2, 2 * p) & mhi;
						const uint32_t fl = funnel_r(g0, g1, 2 * (7 - p)), fh = funnel_r(g1, g2, 2 * (7 - p)) & mhi;
						const uint64_t rv = (uint64_t)rh << 32 | rl, fw = (uint64_t)fh << 32 | fl;
						uint64_t x = GD_SK_MAXU64;
						if ((validbits >> p & 1) && fw != rv) {
							const bool z = !(fw < rv);
							uint32_t lo = z ? rl : fl, hi = z ? rh : fh;
							sk_hash_hl(lo, hi, mhi);
							x = (uint64_t)(hi | 0x01000000u) << 32 | lo;
							zbits |= (uint32_t)z << p, okbits |= 1u << p;
						}
						key[p] = x;
					}
				}
			}
			// emit candidates: inside the tile's emit range and the sequence, and a k-mer
			uint32_t cand;
			{
				const int elo = (packed ? 0 : HL) - s0, ehi = (packed ? NP - wm1 : imin(HL + S.TP, dl - B0)) - s0; // p in [elo, ehi)
				cand = sk_mask_from(elo) & ~sk_mask_from(ehi) & okbits;
			}
			if (pend && wid == 0) lookback_finish(); // its offset travels with the next barrier
			if (w >= 9) {
				// ---- phase 3: minimum of every full window ending at e = s0+p (0 = no full window ends here).  The window starts
				// at a = e-(w-1) in chunk tid-dt, slot q (the same dt, q for every thread): own prefix minimum, whole chunks in
				// between, suffix minimum of the first chunk.  A full window never starts before slot 0 (full => run <= e+1) ----
				uint64_t pre[8], M[8];
				{
					uint64_t suf = GD_SK_MAXU64, pr = GD_SK_MAXU64;
#pragma unroll
					for (int p = 7; p >= 0; --p) suf = sk_min64(suf, key[p]), sm->SUF[p * THREADS + tid] = suf;
#pragma unroll
					for (int p = 0; p < 8; ++p) pr = sk_min64(pr, key[p]), pre[p] = pr;
				}
				sync_block();
				if (pend) copy_out(), pend = 0;
				// whole chunks between the window's first chunk and the thread's own: tid-dt+1 .. tid-1 with dt = D for the upper
				// positions of the chunk and D + 1 for the lower ones -- two minima per thread, not one loop per position
				const int D = -((7 - wm1) >> 3);
				uint64_t midA = GD_SK_MAXU64, midB = GD_SK_MAXU64;
				if (tid >= D) {
#pragma unroll 1
					for (int c = 1; c < D; ++c) midA = sk_min64(midA, sm->SUF[tid - D + c]); // SUF[0][c] = minimum of chunk c
					midB = sk_min64(midA, sm->SUF[tid - D]);
				}
#pragma unroll
				for (int p = 0; p < 8; ++p) {
					const int d = p - wm1, q = d & 7, dt = -(d >> 3); // d < 0; a = 8 (tid - dt) + q
					uint64_t m = 0;
					if (fullbits >> p & 1) m = sk_min64(sk_min64(pre[p], sm->SUF[q * THREADS + tid - dt]), dt == D ? midA : midB);
					M[p] = m;
				}
				// ---- phase 4: a key is emitted iff it equals the largest full-window minimum among the windows that contain its
				// position, i.e. the maximum of M over [i, i+w-1] (a window of only non-k-mers has minimum ~0, but such a window
				// cannot contain a k-mer's position): own suffix maximum, whole chunks, prefix maximum of the last chunk ----
				uint64_t sufm[8];
				{
					uint64_t pm = 0, sx = 0;
#pragma unroll
					for (int p = 0; p < 8; ++p) pm = sk_max64(pm, M[p]), sm->PREM[p * THREADS + tid] = pm;
#pragma unroll
					for (int p = 7; p >= 0; --p) sx = sk_max64(sx, M[p]), sufm[p] = sx;
				}
				sync_block();
				// whole chunks between the thread's own and the last window's last chunk: tid+1 .. tid+dt-1, dt = E or E + 1
				const int E = wm1 >> 3;
				uint64_t mxA = 0, mxB = 0;
				if (cand) { // indices are clamped: a candidate's own chunks are always inside the tile
#pragma unroll 1
					for (int c = 1; c < E; ++c) mxA = sk_max64(mxA, sm->PREM[7 * THREADS + imin(tid + c, THREADS - 1)]); // PREM[7][c] = maximum of chunk c
					mxB = sk_max64(mxA, sm->PREM[7 * THREADS + imin(tid + E, THREADS - 1)]);
				}
#pragma unroll
				for (int p = 0; p < 8; ++p) {
					if (cand >> p & 1) {
						const int d = p + wm1, q = d & 7, dt = d >> 3; // last position of the last window: 8 (tid + dt) + q < NP in the emit range
						const uint64_t mx = sk_max64(sk_max64(sufm[p], sm->PREM[q * THREADS + tid + dt]), dt == E ? mxA : mxB);
						if (mx == key[p]) emit |= 1u << p, ++cnt;
					}
				}
			} else {
				// ---- small windows (w <= 8): direct scans over the keys and M in shared memory ----
#pragma unroll
				for (int p = 0; p < 8; ++p) sm->SUF[p * THREADS + tid] = key[p];
				sync_block();
				if (pend) copy_out(), pend = 0;
#pragma unroll
				for (int p = 0; p < 8; ++p) {
					const int sp = s0 + p;
					uint64_t m = 0;
					if (fullbits >> p & 1) {
						m = GD_SK_MAXU64;
#pragma unroll 1
						for (int d = 0; d < w; ++d) m = sk_min64(m, sm->SUF[((sp - d) & 7) * THREADS + ((sp - d) >> 3)]);
					}
					sm->PREM[p * THREADS + tid] = m;
				}
				sync_block();
#pragma unroll
				for (int p = 0; p < 8; ++p) {
					const int sp = s0 + p;
					if (cand >> p & 1) {
						uint64_t mx = 0;
#pragma unroll 1
						for (int d = 0; d < w; ++d) mx = sk_max64(mx, sm->PREM[((sp + d) & 7) * THREADS + ((sp + d) >> 3)]);
						if (mx == key[p]) emit |= 1u << p, ++cnt;
					}
				}
			}
		} else if (pend) { // a job too short to emit: the pending tile leaves here
			if (wid == 0) lookback_finish();
			sync_block();
			copy_out(), pend = 0;
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
		// a record: x = hash64 << 8 | k (the marker bit 56 leaves at the top), y = rid << 32 | position << 1 | strand
		auto put_at = [&](uint64_t *dst, int p, uint32_t j, uint32_t sh, uint32_t rid) { // j: sparsified position in its job
			uint32_t real;
			if (ones1) real = j * (uint32_t)S.W + sm->ones_loc[0] + sh;
			else {
				const uint32_t qd = j / (uint32_t)S.ones, rm = j - qd * (uint32_t)S.ones;
				real = qd * (uint32_t)S.W + sm->ones_loc[rm] + sh;
			}
			dst[0] = key[p] << 8 | (uint64_t)k;
			dst[1] = (uint64_t)rid << 32 | (uint64_t)real << 1 | (uint64_t)(zbits >> p & 1);
		};
		// where the records go: 0 = the job's own slot of fixed_stride records, 1 = the slots of the tile's jobs (packed tile),
		// 2 = the parking area, 3 = the dense output (after waiting for the offset)
		int mode;
		long long obase = 0;
		if (packed) {
			// records of job g = ranks [R_g, R_g+1) of the tile's records, R_g = records in slots below pk_base[g]: the thread
			// whose chunk holds that slot knows it
			const int n = sm->pk_n;
			int g = 0;
#pragma unroll 1
			while (g <= n && sm->pk_base[g] < s0) ++g;
#pragma unroll 1
			for (; g <= n && sm->pk_base[g] < s0 + 8; ++g) sm->pk_R[g] = local + popc(emit & ((1u << (sm->pk_base[g] - s0)) - 1u));
			sync_block();
			if (tid < n) {
				const long long jb = (long long)job + tid;
				B.out_off[jb] = jb * B.fixed_stride, B.out_cnt[jb] = sm->pk_R[tid + 1] - sm->pk_R[tid];
				if (jb == B.njobs - 1) B.out_off[B.njobs] = (long long)B.njobs * B.fixed_stride;
			}
			mode = 1;
		} else if (fixed) { // nothing orders the tiles
			if (tid == 0) {
				B.out_off[job] = (long long)job * B.fixed_stride, B.out_cnt[job] = total;
				if (tile == B.ntiles - 1) B.out_off[B.njobs] = (long long)B.njobs * B.fixed_stride;
			}
			mode = 0, obase = (long long)job * B.fixed_stride + local;
		} else { // publish the count; park the records until the look-back inside the next tile has their offset
			if (tid == 0) { // an aggregate for the look-backs of later tiles (tile 0: already its inclusive prefix)
				st_volatile(&B.status[tile], ((tile == 0 ? 2ull : 1ull) << 62) | (unsigned long long)total);
				fence();
			}
			pend = 1, pend_tile = tile, pend_total = total, pend_job = job, pend_first = i0 == 0;
			if (total <= SketchSmem3<THREADS>::PARK) mode = 2;
			else { // too many records to park: wait for the offset, write from registers
				if (wid == 0) lookback_issue(), lookback_finish();
				sync_block();
				mode = 3, obase = sm->excl + local, pend = 0;
			}
		}
		if (emit) {
			int o = 0, g = 0;
#pragma unroll
			for (int p = 0; p < 8; ++p)
				if (emit >> p & 1) {
					uint64_t *dst;
					bool ok = true;
					uint32_t jj = (uint32_t)(j0 + p), sh = shift, rid = h->rid; // j0 + p >= 0: a k-mer
					if (PACKED) {
						const int sl = s0 + p;
#pragma unroll 1
						while (sl >= sm->pk_end[g]) ++g; // a record's slot lies inside a job
						const int r = local + o - sm->pk_R[g];
						ok = r < B.fixed_stride, dst = B.out + 2 * (((long long)job + g) * B.fixed_stride + r);
						jj = (uint32_t)(sl - sm->pk_base[g]), sh = sm->pk_shift[g], rid = sm->pk_rid[g];
					} else if (mode == 0) ok = local + o < B.fixed_stride, dst = B.out + 2 * (obase + o);
					else if (mode == 2) dst = (uint64_t *)sm->park + 2 * (local + o);
					else ok = obase + o < B.out_cap, dst = B.out + 2 * (obase + o);
					if (ok) put_at(dst, p, jj, sh, rid);
					++o;
				}
		}
		cur ^= 1;
	}
	if (pend) { // the block's last tile
		if (wid == 0) lookback_issue(), lookback_finish();
		sync_block();
		copy_out();
	}
}

} // namespace gd
