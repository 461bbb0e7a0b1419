// tests/emu/emu_sam.cpp -- TEST INFRASTRUCTURE ONLY.
// Runs csrc/gd_sam_core.h -- the per-read SAM stage the GPU executes with one thread per read -- as plain host code, so its
// output can be compared with the threaded host implementation (host/gd_sr_post.cpp, itself pinned against the reference
// program's SAM) on a machine without a GPU.
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include "gd_sam_core.h"

static inline uint8_t nt4(unsigned char c)
{
	switch (c) {
	case 'A': case 'a': case 0: return 0;
	case 'C': case 'c': case 1: return 1;
	case 'G': case 'g': case 2: return 2;
	case 'T': case 't': case 'U': case 'u': case 3: return 3;
	default: return 4;
	}
}

extern "C" int emu_sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq, const char *qual,
                             const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int64_t n_cigar, int n_seq,
                             const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                             const gd_sr_post_opt_t *opt, char **sam, size_t *sam_len)
{
	std::string blob;
	std::vector<int32_t> noff;
	size_t max_name = 0;
	for (int i = 0; i < n_seq; ++i) {
		noff.push_back((int32_t)blob.size());
		blob += seq_names[i], blob.push_back('\0');
		max_name = std::max(max_name, strlen(seq_names[i]));
	}
	gdsam::RefNames N = {blob.data(), noff.data()};
	std::vector<uint32_t> pool(cigar, cigar + n_cigar + 1); // edited in place, like the device copy
	std::string out;
	std::vector<uint8_t> qc, tc;
	std::vector<gd_sr_cand_t> local;
	for (int i = 0; i < n; ++i) {
		const int nc = (int)(cand_off[i + 1] - cand_off[i]);
		int64_t stride = 16, ncig = 0;
		for (int j = 0; j < nc; ++j) {
			const gd_sr_cand_t &c = cand[cand_off[i] + j];
			stride = std::max<int64_t>(stride, std::max(c.qe - c.qs, c.re - c.rs) + 16);
			ncig += c.n_cigar > 0 ? c.n_cigar : 0;
		}
		qc.assign((size_t)(stride * std::max(nc, 1)), 0), tc.assign((size_t)(stride * std::max(nc, 1)), 0);
		const char *rd = seq + off[i];
		for (int j = 0; j < nc; ++j) { // the code strings of map.c:737-757 (what gd_sr_window_kernel leaves on the device)
			const gd_sr_cand_t &c = cand[cand_off[i] + j];
			const int nq = c.qe - c.qs, tl = c.re - c.rs;
			uint8_t *qd = qc.data() + j * stride, *td = tc.data() + j * stride;
			if (c.rev)
				for (int k = 0; k < nq; ++k) qd[k] = nt4((unsigned char)rd[c.qe - 1 - k]) ^ 3;
			else
				for (int k = 0; k < nq; ++k) qd[k] = nt4((unsigned char)rd[c.qs + k]);
			const unsigned char *tp = (const unsigned char *)ref + ref_off[c.rid] + c.rs;
			const int t_in = std::min(tl, ref_len[c.rid] - c.rs);
			for (int k = 0; k < t_in; ++k) td[k] = nt4(tp[k]);
		}
		gdsam::ReadIn R = {names[i], rd, qual ? qual + off[i] : nullptr, len[i], nc, cand + cand_off[i], pool.data(), qc.data(), tc.data(), stride};
		// count on a scratch copy of the pool (one_read edits CIGARs in place), then write
		std::vector<uint32_t> save(pool);
		gdsam::Sink cnt = {nullptr, 0};
		gdsam::one_read(R, *opt, N, cnt);
		pool = save;
		R.cigar = pool.data();
		const size_t bound = gdsam::text_bound(strlen(names[i]), len[i], nc, ncig, max_name);
		if (cnt.n > bound) return -2; // the slot bound must hold
		const size_t at = out.size();
		out.resize(at + cnt.n);
		gdsam::Sink w = {&out[0] + at, 0};
		gdsam::one_read(R, *opt, N, w);
		if (w.n != cnt.n) return -3;
	}
	*sam = (char *)malloc(out.size() + 1);
	memcpy(*sam, out.data(), out.size());
	*sam_len = out.size();
	return 0;
}

// printf("%.4f") against put_fixed4 for every (mlen, den) pair of a range; returns the number of differing pairs
extern "C" long emu_sam_check_fixed4(int max_den)
{
	long bad = 0;
	char a[64], b[64];
	for (int den = 1; den <= max_den; ++den)
		for (int mlen = 0; mlen <= den; ++mlen) {
			const double div = 1.0 - (double)mlen / den;
			gdsam::Sink s = {b, 0};
			gdsam::put_fixed4(s, div);
			b[s.n] = 0;
			snprintf(a, sizeof a, "%.4f", div);
			if (strcmp(a, b)) ++bad;
		}
	return bad;
}
