// tests/emu/asan_ksw_driver.cpp -- TEST INFRASTRUCTURE ONLY: drives the emulated DP + traceback kernels (emu_ksw.cpp) over random
// pairs, bands, flags and gang sizes with exact-size heap buffers, to be built with -fsanitize=address: any access of the device
// code outside the packed sequences, the backtrack arena, the result records or the blocks' shared memory aborts the program.
// (tests/test_emu_logic.py::test_emu_ksw_address_sanitizer builds and runs it.)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

struct Res { int32_t v[16]; }; // gd::KswResult: 16 x int32
extern "C" int emu_ksw_batch(int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf, const int32_t *tlen,
                             const int64_t *toff, const uint8_t *tbuf, const int32_t *w, int m, const int8_t *mat, int q, int e,
                             int q2, int e2, int zdrop, int end_bonus, int flag, int G, int threads, Res *res, uint32_t *cigar,
                             int cigar_stride, int *lead64_count);

int main()
{
	srand(11);
	const int flags[] = {0x08, 0x00, 0x40, 0x48, 0x42, 0x18, 0x0a, 0x09, 0x80, 0xc2};
	const int gangs[] = {4, 8, 16, 32, 64, 128};
	int8_t mat[25];
	for (int i = 0; i < 25; ++i) mat[i] = (i / 5 == 4 || i % 5 == 4) ? 0 : (i / 5 == i % 5 ? 2 : -8);
	for (int it = 0; it < 10; ++it) {
		const int flag = flags[it % 10], G = gangs[it % 6], n = 5;
		const int maxlen = G >= 64 ? 420 : G >= 16 ? 300 : 200;
		std::vector<int32_t> ql(n), tl(n), w(n);
		std::vector<int64_t> qo(n), to(n);
		size_t qs = 0, ts = 0;
		for (int i = 0; i < n; ++i) {
			ql[i] = 1 + rand() % maxlen, tl[i] = 1 + rand() % maxlen;
			w[i] = rand() % 4 == 0 ? rand() % 6 : rand() % (maxlen + 40);
			qo[i] = qs, to[i] = ts, qs += ql[i], ts += tl[i];
		}
		uint8_t *qb = (uint8_t *)malloc(qs), *tb = (uint8_t *)malloc(ts); // exact size
		for (size_t i = 0; i < qs; ++i) qb[i] = rand() % 40 == 0 ? 4 : rand() & 3;
		for (int i = 0; i < n; ++i) // the target: the query with edits, so that alignments are long
			for (int j = 0; j < tl[i]; ++j) tb[to[i] + j] = (j < ql[i] && rand() % 10) ? qb[qo[i] + j] : (uint8_t)(rand() & 3);
		int stride = 2 * maxlen + 8, l64 = 0;
		std::vector<Res> res(n);
		std::vector<uint32_t> cig((size_t)n * stride);
		int rc = emu_ksw_batch(n, ql.data(), qo.data(), qb, tl.data(), to.data(), tb, w.data(), 5, mat, 12, 2, 24, 1, it % 3 ? 100 : 30, 10, flag, G, 64,
		                       res.data(), cig.data(), stride, &l64);
		printf("it %d flag %#x G %d rc %d lead64 %d score0 %d\n", it, flag, G, rc, l64, res[0].v[8]);
		if (rc) return 1;
		free(qb), free(tb);
	}
	return 0;
}
