// gd_mmi.cpp -- row F4 of SURVEY.md section 8: the reference's `.mmi` index file (mm_idx_dump,
// GDiet-ShortReads/index.c:480-517) written from the arrays of the device-built index (gd_index_export), byte for
// byte what `GDiet_avx -d` writes.
//
// The file stores, per bucket (minimizer & (2^b - 1)), the position lists of the minimizers that occur more than once
// and the bucket's khash table in SLOT ORDER, so the table has to be rebuilt exactly as worker_post does
// (index.c:216-271): kh_resize(n_keys), then kh_put in ascending minimizer order, with klib's triangular probing,
// its 0.77 load bound and its in-place rehash (khash.h:232-334).  Host C++; no GPU work.
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include "../../include/gdiet_cuda.h"

namespace {

// klib khash (khash.h) for key = uint64, value = uint64, hash(a) = a >> 1, equal(a, b) = (a >> 1 == b >> 1), restricted
// to what worker_post uses: resize + put of distinct keys, no deletions by the caller.
struct KHash {
	uint32_t n_buckets = 0, size = 0, n_occupied = 0, upper_bound = 0;
	std::vector<uint8_t> flag; // 2 = empty, 1 = deleted, 0 = occupied (the two flag bits of khash.h:166-172)
	std::vector<uint64_t> keys, vals;

	static uint32_t roundup32(uint32_t x)
	{
		--x, x |= x >> 1, x |= x >> 2, x |= x >> 4, x |= x >> 8, x |= x >> 16, ++x;
		return x;
	}
	void resize(uint32_t new_n)
	{ // khash.h:232-288
		new_n = roundup32(new_n);
		if (new_n < 4) new_n = 4;
		if (size >= (uint32_t)(new_n * 0.77 + 0.5)) return; // requested size is too small
		std::vector<uint8_t> new_flag(new_n, 2);
		if (n_buckets < new_n) keys.resize(new_n), vals.resize(new_n);
		const uint32_t new_mask = new_n - 1;
		for (uint32_t j = 0; j != n_buckets; ++j) {
			if (flag[j] != 0) continue;
			uint64_t key = keys[j], val = vals[j];
			flag[j] = 1;
			for (;;) { // kick-out process
				uint32_t step = 0, i = (uint32_t)(key >> 1) & new_mask;
				while (new_flag[i] != 2) i = (i + (++step)) & new_mask;
				new_flag[i] = 0;
				if (i < n_buckets && flag[i] == 0) {
					std::swap(keys[i], key), std::swap(vals[i], val);
					flag[i] = 1;
				} else {
					keys[i] = key, vals[i] = val;
					break;
				}
			}
		}
		if (n_buckets > new_n) keys.resize(new_n), vals.resize(new_n);
		flag.swap(new_flag);
		n_buckets = new_n, n_occupied = size, upper_bound = (uint32_t)(n_buckets * 0.77 + 0.5);
	}
	uint32_t put(uint64_t key)
	{ // khash.h:289-334, key not present
		if (n_occupied >= upper_bound) {
			if (n_buckets > (size << 1)) resize(n_buckets - 1);
			else resize(n_buckets + 1);
		}
		const uint32_t mask = n_buckets - 1;
		uint32_t step = 0, i = (uint32_t)(key >> 1) & mask, site = n_buckets, x = n_buckets;
		if (flag[i] == 2) x = i;
		else {
			const uint32_t last = i;
			while (flag[i] != 2 && (flag[i] == 1 || (keys[i] >> 1) != (key >> 1))) {
				if (flag[i] == 1) site = i;
				i = (i + (++step)) & mask;
				if (i == last) {
					x = site;
					break;
				}
			}
			if (x == n_buckets) x = (flag[i] == 2 && site != n_buckets) ? site : i;
		}
		if (flag[x] == 2) keys[x] = key, flag[x] = 0, ++size, ++n_occupied;
		else if (flag[x] == 1) keys[x] = key, flag[x] = 0, ++size;
		return x;
	}
};

} // namespace

// keys[n_keys] ascending distinct minimizer values, counts[n_keys], positions grouped by key in key order (what
// gd_index_export returns), S = the 4-bit reference; names / lens describe the n_seq contigs.
extern "C" int gd_mmi_write(const char *path, int w, int k, int bucket_bits, int flag, int n_seq, const char *const *names,
                            const int32_t *lens, int64_t n_keys, const uint64_t *keys, const uint32_t *counts,
                            const uint64_t *positions, const uint32_t *S)
{
	if (!path || n_seq < 0 || n_keys < 0 || (n_keys > 0 && (!keys || !counts || !positions))) return GD_ERR_ARG;
	int b = bucket_bits;
	if (k * 2 < b) b = k * 2; // mm_idx_init, index.c:48
	// mm_idx_dump writes S iff !(flag & MM_I_NO_SEQ) (index.c:515) and mm_idx_load reads it by the same test: keep the two in step
	if (S) flag &= ~0x2;
	else flag |= 0x2;
	FILE *fp = fopen(path, "wb");
	if (!fp) return GD_ERR_ARG;
	const uint32_t hdr[5] = {(uint32_t)w, (uint32_t)k, (uint32_t)b, (uint32_t)n_seq, (uint32_t)flag};
	fwrite("MMI\2", 1, 4, fp);
	fwrite(hdr, 4, 5, fp);
	uint64_t sum_len = 0;
	for (int i = 0; i < n_seq; ++i) {
		const uint8_t l = names && names[i] ? (uint8_t)strlen(names[i]) : 0;
		fwrite(&l, 1, 1, fp);
		if (l) fwrite(names[i], 1, l, fp);
		fwrite(&lens[i], 4, 1, fp);
		sum_len += (uint32_t)lens[i];
	}
	// keys of every bucket, ascending (a counting sort by the low b bits keeps the global order inside a bucket)
	const uint32_t nb = 1u << b, mask = nb - 1;
	std::vector<int64_t> start((size_t)nb + 1, 0);
	for (int64_t i = 0; i < n_keys; ++i) ++start[(keys[i] & mask) + 1];
	for (uint32_t i = 0; i < nb; ++i) start[i + 1] += start[i];
	std::vector<int64_t> order((size_t)n_keys), first((size_t)n_keys + 1, 0);
	{
		std::vector<int64_t> at(start.begin(), start.end() - 1);
		for (int64_t i = 0; i < n_keys; ++i) order[at[keys[i] & mask]++] = i;
	}
	for (int64_t i = 0; i < n_keys; ++i) first[i + 1] = first[i] + counts[i];
	std::vector<uint64_t> p;
	for (uint32_t bk = 0; bk < nb; ++bk) { // worker_post, index.c:216-271 + the dump loop, index.c:500-514
		const int64_t lo = start[bk], hi = start[bk + 1];
		p.clear();
		KHash h;
		if (hi > lo) {
			h.resize((uint32_t)(hi - lo));
			for (int64_t j = lo; j < hi; ++j) {
				const int64_t ki = order[j];
				const uint64_t key = keys[ki] >> b << 1;
				const uint32_t itr = h.put(key);
				if (counts[ki] == 1) h.keys[itr] |= 1, h.vals[itr] = positions[first[ki]];
				else {
					h.vals[itr] = (uint64_t)p.size() << 32 | counts[ki];
					p.insert(p.end(), positions + first[ki], positions + first[ki] + counts[ki]);
				}
			}
		}
		const int32_t n = (int32_t)p.size();
		const uint32_t size = h.size;
		fwrite(&n, 4, 1, fp);
		if (n) fwrite(p.data(), 8, (size_t)n, fp);
		fwrite(&size, 4, 1, fp);
		if (size == 0) continue;
		for (uint32_t s = 0; s < h.n_buckets; ++s) {
			if (h.flag[s] != 0) continue;
			const uint64_t x[2] = {h.keys[s], h.vals[s]};
			fwrite(x, 8, 2, fp);
		}
	}
	if (S) fwrite(S, 4, (size_t)((sum_len + 7) / 8), fp);
	const int bad = ferror(fp);
	fclose(fp);
	return bad ? GD_ERR_ARG : GD_OK;
}

// --------------------------------------------------------------------------------------------
// reading (mm_idx_load, GDiet-ShortReads/index.c:519-571): the file back into the flat arrays the device index is
// uploaded from -- keys ascending, counts, positions grouped by key.  Implemented in the host part; gd_index_load_mmi
// (gd_index.cu) puts them into HBM.
// --------------------------------------------------------------------------------------------
#include <algorithm>

struct GdMmiData {
	int w = 0, k = 0, b = 0, flag = 0;
	std::vector<std::string> names;
	std::vector<int32_t> lens;
	std::vector<uint64_t> keys, positions;
	std::vector<uint32_t> counts, S;
};

int gd_mmi_parse(const char *path, GdMmiData &D)
{
	FILE *fp = fopen(path, "rb");
	if (!fp) return GD_ERR_ARG;
	char magic[4];
	uint32_t x[5];
	bool ok = fread(magic, 1, 4, fp) == 4 && !memcmp(magic, "MMI\2", 4) && fread(x, 4, 5, fp) == 5;
	if (!ok) {
		fclose(fp);
		return GD_ERR_ARG;
	}
	D.w = (int)x[0], D.k = (int)x[1], D.b = (int)x[2], D.flag = (int)x[4];
	if (D.k < 1 || D.k > 28 || D.w < 1 || D.w > 255 || D.b < 0 || D.b > 28 || D.b > 2 * D.k) { // a corrupt header must not size an allocation
		fclose(fp);
		return GD_ERR_ARG;
	}
	uint64_t sum_len = 0;
	for (uint32_t i = 0; ok && i < x[3]; ++i) {
		uint8_t l = 0;
		char nm[256];
		int32_t len = 0;
		ok = fread(&l, 1, 1, fp) == 1 && (l == 0 || fread(nm, 1, l, fp) == l) && fread(&len, 4, 1, fp) == 1;
		D.names.emplace_back(nm, nm + l), D.lens.push_back(len), sum_len += (uint32_t)len;
	}
	struct Ent {
		uint64_t key, a, n; // minimizer, first position (within its bucket's list, or the position itself), count
		uint32_t bucket;
	};
	std::vector<Ent> ents;
	std::vector<std::vector<uint64_t>> plists((size_t)1 << D.b);
	for (uint32_t bk = 0; ok && bk < (1u << D.b); ++bk) {
		int32_t n = 0;
		uint32_t size = 0;
		ok = fread(&n, 4, 1, fp) == 1;
		if (ok && n > 0) plists[bk].resize((size_t)n), ok = fread(plists[bk].data(), 8, (size_t)n, fp) == (size_t)n;
		ok = ok && fread(&size, 4, 1, fp) == 1;
		for (uint32_t j = 0; ok && j < size; ++j) {
			uint64_t kv[2];
			ok = fread(kv, 8, 2, fp) == 2;
			Ent e;
			e.key = (kv[0] >> 1) << D.b | bk, e.bucket = bk; // index.c:246-258 inverted
			if (kv[0] & 1) e.a = kv[1], e.n = 0;              // singleton: the value is the position
			else e.a = kv[1] >> 32, e.n = (uint32_t)kv[1];
			ents.push_back(e);
		}
	}
	if (ok && !(D.flag & 0x2 /* MM_I_NO_SEQ */)) {
		D.S.resize((size_t)((sum_len + 7) / 8));
		ok = D.S.empty() || fread(D.S.data(), 4, D.S.size(), fp) == D.S.size();
	}
	fclose(fp);
	if (!ok) return GD_ERR_ARG;
	std::sort(ents.begin(), ents.end(), [](const Ent &p, const Ent &q) { return p.key < q.key; });
	D.keys.reserve(ents.size()), D.counts.reserve(ents.size());
	for (const Ent &e : ents) {
		D.keys.push_back(e.key);
		if (e.n == 0) D.counts.push_back(1), D.positions.push_back(e.a);
		else {
			D.counts.push_back((uint32_t)e.n);
			const std::vector<uint64_t> &pl = plists[e.bucket];
			if (e.a + e.n > pl.size()) return GD_ERR_ARG;
			D.positions.insert(D.positions.end(), pl.begin() + (size_t)e.a, pl.begin() + (size_t)(e.a + e.n));
		}
	}
	return GD_OK;
}
