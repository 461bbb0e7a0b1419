/* oracle/gd_oracle.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C, scalar, one-cell-at-a-time restatement of the two reference kernels on the
 * Genome-on-Diet per-read mapping hot path.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load this; the product (libgdiet_cuda.so) never does.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py checks every function below against the
 * unmodified reference compiled into oracle/_ref/ (ksw_extd2_sse, ksw_extd2_avx512, mm_sketch,
 * mm_sketch2, mm_sketch3 of both the scalar and the AVX-512 build), and tests/golden/ holds
 * vectors generated from those reference builds (tests/golden/make_golden.py).
 */
#ifndef GD_ORACLE_H
#define GD_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* flat image of ksw_extz_t (GDiet-ShortReads/ksw2.h:31-40) without the heap pointer */
typedef struct {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
} gdo_extz_t;

#define GDO_NEG_INF (-0x40000000)

/* score_rule: 0 = compare rule of ksw2_extd2_sse.c:166-180; 1 = xor-table rule of
 * ksw2_extd2_avx.c:187-208,312-313 (the parity target, GDiet_avx). They differ only for codes > 4. */
int gdo_ksw_extd2(int qlen, const uint8_t *query, int tlen, const uint8_t *target, int m, const int8_t *mat, int q,
                  int e, int q2, int e2, int w, int zdrop, int end_bonus, int flag, int score_rule, gdo_extz_t *ez,
                  uint32_t *cigar, int cigar_cap);

/* number of banded cells the reference visits: sum over executed rows of en0-st0+1.
 * rows_done = number of anti-diagonals executed before a band-closure break (Z-drop breaks are data
 * dependent and not modelled here; pass the row count from a DP run if needed). */
int64_t gdo_band_cells(int qlen, int tlen, int w);

int gdo_exact_match(int qlen, const uint8_t *query, int tlen, const uint8_t *target);

/* Position-parallel minimizer model, the AVX-512 build's N rule (>=).
 * Generic worker: sketch str[0..len_crop) with pattern Z/W at the given shift; keep at most cap
 * entries (cap==0 => uncapped). Returns the number of entries the reference would have pushed
 * (== min(n, cap)); writes min(n, out_cap) of them. *last_y receives y of the last kept entry. */
long gdo_sketch_core(const char *str, unsigned len_crop, int w, int k, uint32_t rid, const char *Z, int W,
                     unsigned shift, uint64_t cap, uint64_t *out_xy, long out_cap, uint64_t *last_y);

long gdo_mm_sketch(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, uint64_t *out_xy,
                   long cap);
long gdo_mm_sketch3(const char *str, unsigned len, int w, int k, uint32_t rid, const char *Z, int W, int shift,
                    uint32_t max_nb_seeds, uint64_t *out_xy, long cap, uint32_t *ret);
long gdo_mm_sketch2(const char *str, int len, int w, int k, uint32_t rid, const char *Z, int W, float max_seeds,
                    uint64_t *out_xy, long cap, uint32_t *counts);

uint64_t gdo_hash64(uint64_t key, uint64_t mask);

#ifdef __cplusplus
}
#endif
#endif
