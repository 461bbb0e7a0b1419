/* oracle/gd_oracle_map.h -- TEST INFRASTRUCTURE ONLY (see oracle/README.md and gd_oracle_map.c). */
#ifndef GD_ORACLE_MAP_H
#define GD_ORACLE_MAP_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* the reference's mm_idx_t (GDiet-ShortReads/minimap.h:86-96) reduced to what a lookup returns */
typedef struct {
	int32_t n_seq, w, k;
	uint32_t *len;    /* contig lengths */
	uint64_t *offset; /* contig offsets into codes[] */
	uint8_t *codes;   /* nt4 codes 0..4 of all contigs (what mm_idx_getseq yields, index.c:157-166) */
	int64_t n_keys;   /* distinct minimizer values (x >> 8), ascending */
	uint64_t *key, *start;
	uint32_t *cnt;
	int64_t n_pos; /* y values ordered by (key, y) */
	uint64_t *pos;
} gdo_index_t;

/* mapping options the stage reads (GDiet-ShortReads/minimap.h:142-205, main.c:166-182, options.c:130-150);
 * the SAME layout as gd_sr_opt_t of include/gdiet_cuda.h */
typedef struct {
	int32_t W;
	char Z[64];
	float max_seeds;
	int32_t frag_mode, max_frag_len;
	uint32_t bw; /* used when bw_max == 0 */
	float min_cnt, rec_threshold_frac;
	int32_t af_max_loc, mid_occ, max_max_occ, occ_dist;
	float q_occ_frac;
	int32_t for_only, rev_only;
	int32_t a, b, q, e, q2, e2, zdrop, end_bonus;
	float bw_frac; /* -r bw_frac,bw_min,bw_max: per-read band, map.c:624-631 (when bw_max > 0) */
	uint32_t bw_min, bw_max;
} gdo_sr_opt_t;

/* one candidate location of one read, in the order of the reference's candidate loop (map.c:764);
 * the SAME layout as gd_sr_cand_t of include/gdiet_cuda.h */
typedef struct {
	int32_t rid, rs, re, qs, qe, rev; /* the mm_reg1_t fields set at map.c:932-938 */
	int32_t votes, first_q, last_q;   /* vt_t of map.c:433-440 */
	int32_t exact;                    /* exact_match_sse said equal: no DP (map.c:873-915) */
	int32_t score, n_cigar;           /* ez.score, ez.n_cigar */
	int32_t cigar_off;                /* first entry in the read's / the batch's CIGAR pool */
	int32_t reserved[3];
} gdo_sr_cand_t;

/* long-read tree: the mm_mapopt_t fields GDiet-LongReads/map.c:1273-1853 reads (defaults LR/main.c:170-182);
 * the SAME layout as gd_lr_opt_t of include/gdiet_cuda.h */
typedef struct {
	int32_t W;
	char Z[64];
	float max_seeds;
	int32_t frag_mode, max_frag_len;
	uint32_t bw; /* opt->bw (-r) */
	int32_t mid_occ, max_max_occ, occ_dist;
	float q_occ_frac;
	int32_t for_only, rev_only;
	int32_t a, b, q, e, q2, e2, zdrop, end_bonus;
	uint32_t vt_dis, vt_nb_loc;
	float vt_cov, vt_df1, vt_df2, vt_f;
	uint32_t max_max_gap, max_min_gap;
} gdo_lr_opt_t;

typedef struct {
	uint32_t shift, tmp_extracted_len, n_mv, n_a_for, n_a_rev, vt_threshold, nb_potentials, reserved;
} gdo_sr_dbg_t;

gdo_index_t *gdo_index_build(int n_seq, const char *buf, const int64_t *off, const int32_t *len, int w, int k,
                             const char *Z, int W);
void gdo_index_destroy(gdo_index_t *mi);
const uint64_t *gdo_index_get(const gdo_index_t *mi, uint64_t minier, int *n);
int32_t gdo_index_cal_max_occ(const gdo_index_t *mi, float f);

/* returns the number of candidates written to out[] (<= out_cap); CIGARs are appended to cigar[] */
int gdo_sr_map_read(const gdo_index_t *mi, const char *seq, int qlen, const gdo_sr_opt_t *o, gdo_sr_cand_t *out,
                    int out_cap, uint32_t *cigar, int cigar_cap, gdo_sr_dbg_t *dbg);

/* long reads: candidates in the order of the loop at LR/map.c:1654; reserved[0] = index of the candidate this one is
 * continued by (vt_t::next, -1 = none), reserved[1] = vt_t::concat */
int gdo_lr_map_read(const gdo_index_t *mi, const char *seq, int qlen, const gdo_lr_opt_t *o, gdo_sr_cand_t *out, int out_cap,
                    uint32_t *cigar, int cigar_cap, gdo_sr_dbg_t *dbg);

#ifdef __cplusplus
}
#endif
#endif
