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
	// Fixed-stride output (one tile per job only): job j writes at most fixed_stride records at out[j * fixed_stride ..]
	// and its count to out_cnt[j].  No tile depends on another one then: no ticket, no look-back.
	int64_t fixed_stride;      // 0 = dense output in job order (decoupled look-back)
	int32_t *out_cnt;          // [njobs], fixed-stride mode
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

} // namespace gd
