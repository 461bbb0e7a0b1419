// gd_index.cuh -- device-resident minimizer index: the layout the kernels read.
//
// The reference keeps 2^14 buckets, each a khash from (minimizer >> 14) to either one position or a
// slice of the bucket's position array (GDiet-ShortReads/index.c:84-100,216-271).  On the GPU the whole
// index is ONE open-addressing table of 16-byte slots {minimizer, first, count} over ONE flat position
// array: a lookup is a single 16-byte load in the common case and returns what mm_idx_get returns -- the
// count and the positions of that minimizer in ascending order.
#pragma once
#include "gd_common.cuh"
#include <string>
#include <vector>

namespace gd {

#define GD_IDX_EMPTY 0xffffffffffffffffull // minimizer values have at most 56 bits

struct IdxSlot {
	unsigned long long key;
	uint32_t first, count;
};

struct IndexDev {
	const IdxSlot *tab;
	uint64_t tab_mask;
	const uint64_t *pos;  // y values (rid<<32 | last_base<<1 | strand), grouped by minimizer
	const uint32_t *S;    // 4-bit packed reference, 8 bases per word (GDiet-ShortReads/mmpriv.h:31-32)
	const uint64_t *seq_off;
	const uint32_t *seq_len;
	int32_t n_seq, w, k;
};

GD_DEV uint64_t idx_mix(uint64_t h)
{ // the minimizer value is already hash64() of the k-mer; one multiply-shift spreads its low bits
	h ^= h >> 29;
	h *= 0x9E3779B97F4A7C15ull;
	return h ^ (h >> 32);
}

// mm_idx_get (index.c:84-100): returns the count, sets first
GD_DEV uint32_t idx_get(const IndexDev &I, uint64_t minier, uint32_t &first)
{
	uint64_t s = idx_mix(minier) & I.tab_mask;
	for (;;) {
		const ulonglong2 raw = *(const ulonglong2 *)(I.tab + s);
		if (raw.x == minier) {
			first = (uint32_t)raw.y;
			return (uint32_t)(raw.y >> 32);
		}
		if (raw.x == GD_IDX_EMPTY) {
			first = 0;
			return 0;
		}
		s = (s + 1) & I.tab_mask;
	}
}

// mm_seq4_get (mmpriv.h:32)
GD_DEV uint32_t idx_base(const IndexDev &I, uint64_t i) { return I.S[i >> 3] >> ((i & 7) << 2) & 0xf; }

} // namespace gd

// host-side object behind the opaque gd_index of include/gdiet_cuda.h
struct gd_index {
	int device = 0;
	gd::IndexDev d;
	int64_t n_seq = 0, total_len = 0, n_min = 0, n_keys = 0, tab_slots = 0, s_words = 0;
	void *d_tab = nullptr, *d_pos = nullptr, *d_S = nullptr, *d_seq_off = nullptr, *d_seq_len = nullptr;
	void *d_keys = nullptr, *d_counts = nullptr; // sorted distinct minimizers + counts (export, max_occ)
	uint32_t *h_seq_len = nullptr;
	uint64_t *h_seq_off = nullptr;
	size_t device_bytes = 0;
	std::vector<std::string> names; // contig names (only known when the index came from a .mmi file)
};
