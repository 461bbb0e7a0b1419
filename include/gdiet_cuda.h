/* gdiet_cuda.h -- C ABI of libgdiet_cuda.so: the B200 (sm_100a) replacement for the two kernels on
 * Genome-on-Diet's per-read mapping hot path.
 *
 *   (1) ksw2 dual-affine banded extension DP      reference: GDiet-ShortReads/ksw2.h:68-69,
 *       ksw_extd2_sse / ksw_extd2_avx512          GDiet-ShortReads/ksw2_extd2_avx.h:38,
 *                                                 dispatch GDiet-ShortReads/ksw2_dispatch.c:79-92
 *   (2) sparsified minimizer sketching            reference: GDiet-ShortReads/mmpriv.h:63-68
 *       mm_sketch / mm_sketch2 / mm_sketch3       (GDiet-ShortReads/sketch.c:156,618,1078,2143)
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  Two layers:
 *   - drop-in entry points with EXACTLY the reference's names and signatures (batch of one);
 *   - batched entry points (gd_*) taking either host buffers or device-resident buffers.
 * There is no CPU fallback: every entry point fails loudly (error code + gd_strerror, or abort()
 * for the void drop-in functions) when no CUDA device is usable.
 */
#ifndef GDIET_CUDA_H
#define GDIET_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* types                                                                                       */
/* ------------------------------------------------------------------------------------------ */

/* ksw_extz_t exactly as GDiet-ShortReads/ksw2.h:31-40 (the drop-in functions fill this one).
 * Guarded so that a host program that already includes the reference's ksw2.h can include us. */
#ifndef KSW2_H_
typedef struct {
	uint32_t max : 31, zdropped : 1;
	int max_q, max_t;
	int mqe, mqe_t;
	int mte, mte_q;
	int score;
	int m_cigar, n_cigar;
	int reach_end;
	uint32_t *cigar;
} ksw_extz_t;
#endif

/* mm128_t / mm128_v / mm_pattern_t exactly as GDiet-ShortReads/minimap.h:69-76,99-102 */
#ifndef MINIMAP2_H
typedef struct {
	uint64_t x, y;
} mm128_t;
typedef struct {
	size_t n, m;
	mm128_t *a;
} mm128_v;
typedef struct {
	uint32_t n;
	uint32_t *shift_seeds_number;
} mm_pattern_t;
#endif

/* Flat per-pair result of the batched DP (64 bytes). The first 11 fields are the fields of
 * ksw_extz_t; n_cigar < 0 means the CIGAR needed -n_cigar entries and did not fit. */
typedef struct {
	int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
	int32_t tb_i, tb_j; /* traceback start cell (target, query), -1 = none */
	int32_t rows_done;  /* anti-diagonals executed */
	int32_t lead64;     /* 1: the walk read cells the AVX-512 build computes left of the 16-aligned row start
	                       (ksw2_extd2_avx.c:242,442); the CIGAR comes from the literal 64-lane model */
	int32_t reserved;
} gd_extz_t;

/* Scoring / control arguments of ksw_extd2_sse(), batch-uniform (GDiet-ShortReads/ksw2.h:42-59). */
typedef struct {
	int32_t m;         /* alphabet size (5 at every call site) */
	const int8_t *mat; /* m*m scores */
	int32_t q, e, q2, e2;
	int32_t zdrop, end_bonus, flag; /* KSW_EZ_* */
} gd_ksw_params_t;

typedef struct gd_ctx gd_ctx; /* one per host thread / stream; owns device and pinned staging memory */

enum {
	GD_OK = 0,
	GD_ERR_NO_DEVICE = 1,   /* no usable sm_100 device / CUDA failure at init */
	GD_ERR_CUDA = 2,        /* a CUDA call failed, see gd_strerror */
	GD_ERR_ARG = 3,         /* invalid argument (also: unsupported flag KSW_EZ_GENERIC_SC) */
	GD_ERR_CAPACITY = 4     /* caller-provided output buffer too small; required size reported */
};

/* ------------------------------------------------------------------------------------------ */
/* context                                                                                     */
/* ------------------------------------------------------------------------------------------ */
int gd_init(int device, gd_ctx **ctx);
void gd_destroy(gd_ctx *ctx);
const char *gd_strerror(const gd_ctx *ctx); /* ctx may be NULL: last init error */
/* options: "ksw_group" (lanes per pair: 0=auto,4,8,16,32), "p_budget_mb" (backtrack arena),
 * "ksw_blocks_per_sm" (0=auto), "map_lanes" (mapping stage: the slices of a call are dealt to this many contexts, each driven by its own
 * host thread: 1..8, default 4 or the environment variable GDIET_MAP_LANES; long reads use two at most), "time_kernels" (1: bracket the DP / sketch kernel launches with CUDA
 * events).  stats: "kernel_launches", "ksw_ring", "ksw_group", "ksw_chunks", "device_sms"; with
 * "time_kernels": "ksw_dp_us", "ksw_dp_launches", "sketch_us", "sketch_launches" (cumulative device
 * time of the hot kernels; reading one waits for the recorded launches), "ksw_dp_reset" */
int gd_set_option(gd_ctx *ctx, const char *key, long value);
long gd_get_stat(const gd_ctx *ctx, const char *key);
void *gd_stream(gd_ctx *ctx); /* the cudaStream_t all work of this context is issued on */
void *gd_pinned_alloc(size_t bytes); /* page-locked host memory for read / reference staging (NULL on failure) */
void gd_pinned_free(void *p);
long gd_thread_ctx_pool_size(int device); /* contexts of exited drop-in threads parked for reuse on `device` */

/* ------------------------------------------------------------------------------------------ */
/* (1) DP                                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* Drop-in replacements: same names, arguments, result fields and CIGAR ownership as
 * GDiet-ShortReads/ksw2.h:68-69 and GDiet-ShortReads/ksw2_extd2_avx.h:38.  ez->cigar is (re)allocated
 * with krealloc(km, ...) when the host program provides kalloc, else realloc (km must then be NULL);
 * the caller frees it with kfree(km, ez->cigar) exactly as at GDiet-ShortReads/map.c:952.
 * Both names run the same kernel and BOTH answer like ksw_extd2_avx512 (the GDiet_avx build, the parity target):
 * codes > 4 (the 7 a reverse-complemented N becomes, map.c:737-757) score by its xor table
 * (ksw2_extd2_avx.c:187-208; the SSE file's compare rule, ksw2_extd2_sse.c:166-180, scores them as mismatches), and the
 * walk reads the cells that build computes left of the 16-aligned row start (off[r] rounded down to 64, :242,442).  A
 * host built WITHOUT AVX-512 that links this library therefore gets GDiet_avx's output, not its own CPU build's,
 * on reads that hold N and on walks that leave a narrow band.
 * The calling thread's context comes from a per-device pool and goes back to it when the thread exits (kt_for starts
 * fresh threads per mini-batch): gd_thread_ctx_pool_size(). */
void ksw_extd2_sse(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                   const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus,
                   int flag, ksw_extz_t *ez);
void ksw_extd2_avx512(void *km, int qlen, const uint8_t *query, int tlen, const uint8_t *target, int8_t m,
                      const int8_t *mat, int8_t q, int8_t e, int8_t q2, int8_t e2, int w, int zdrop, int end_bonus,
                      int flag, ksw_extz_t *ez);

/* Batched DP over host buffers: pair i is query qbuf[qoff[i] .. +qlen[i]) against target
 * tbuf[toff[i] .. +tlen[i]) (byte codes as at GDiet-ShortReads/map.c:737-757,840), band w[i] (or w_all if
 * w == NULL).  Stages through pinned memory, runs pack + DP + traceback on the context's stream and
 * returns after the results are on the host:
 *   ez[i]                       per-pair result
 *   cigar_off[0..n]             prefix offsets (entries) into cigar[]
 *   cigar[cigar_off[i] .. cigar_off[i+1])   BAM-encoded CIGAR of pair i (len<<4|op)
 * cigar may be NULL (then only ez and cigar_off are produced).  Returns GD_ERR_CAPACITY and sets
 * cigar_off[n] to the required number of entries if cigar_cap is too small. */
int gd_ksw_extd2_batch(gd_ctx *ctx, int n, const int32_t *qlen, const int64_t *qoff, const uint8_t *qbuf,
                       const int32_t *tlen, const int64_t *toff, const uint8_t *tbuf, const int32_t *w, int w_all,
                       const gd_ksw_params_t *prm, gd_extz_t *ez, int64_t *cigar_off, uint32_t *cigar,
                       int64_t cigar_cap);

/* Same, but every array argument is a DEVICE pointer and nothing is copied or synchronised: the
 * work is enqueued on gd_stream(ctx).  max_qlen/max_tlen/max_w are upper bounds over the batch
 * (they size the staging arenas).  d_cigar receives n * cigar_stride entries (pair i at
 * i*cigar_stride, d_ez[i].n_cigar of them valid); pass NULL / KSW_EZ_SCORE_ONLY for scores only. */
int gd_ksw_extd2_batch_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff, const uint8_t *d_qbuf,
                              const int32_t *d_tlen, const int64_t *d_toff, const uint8_t *d_tbuf,
                              const int32_t *d_w, int w_all, int max_qlen, int max_tlen, int max_w,
                              const gd_ksw_params_t *prm, gd_extz_t *d_ez, uint32_t *d_cigar, int cigar_stride);

/* exact_match_sse (GDiet-ShortReads/exact_match_sse.c:27-88) for a batch, device-resident:
 * d_equal[i] = (memcmp(query_i, target_i, qlen[i]) == 0). */
int gd_exact_match_batch_device(gd_ctx *ctx, int n, const int32_t *d_qlen, const int64_t *d_qoff,
                                const uint8_t *d_qbuf, const int64_t *d_toff, const uint8_t *d_tbuf,
                                uint8_t *d_equal);

/* ------------------------------------------------------------------------------------------ */
/* (2) sketching                                                                               */
/* ------------------------------------------------------------------------------------------ */

/* Drop-in replacements (GDiet-ShortReads/mmpriv.h:63-68): same append / allocation behaviour
 * (kv_push semantics on *p with krealloc(km), kmalloc'd shift_seeds_number). */
void mm_sketch(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p, const char *Z,
               int W);
mm_pattern_t mm_sketch2(void *km, const char *str, int len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p,
                        const char *Z, int W, const float max_seeds);
unsigned mm_sketch3(void *km, const char *str, const unsigned len, int w, int k, uint32_t rid, int is_hpc, mm128_v *p,
                    const char *Z, int W, int shift2, uint32_t MAX_NB_SEEDS);

/* the data symbol sketch.c exports next to its functions (GDiet-ShortReads/sketch.c:11-18) */
extern unsigned char seq_nt4_table[256];

/* Index-build sketching of n reference sequences (mm_sketch semantics, rid[i] stamped into y):
 * seqs are ASCII in buf at off[i], len[i].  Output minimizers are concatenated in input order:
 * out[out_off[i] .. out_off[i+1]).  Host buffers; returns GD_ERR_CAPACITY with out_off[n] set when
 * out_cap (entries) is too small. */
int gd_sketch_ref_batch(gd_ctx *ctx, int n, const int64_t *off, const int32_t *len, const uint32_t *rid,
                        const char *buf, int w, int k, const char *Z, int W, int64_t *out_off, mm128_t *out,
                        int64_t out_cap);

/* Device-resident variant: everything is a device pointer, nothing is synchronised.
 * d_out_count[0] receives the total; d_out_off[0..n] the per-sequence offsets. */
int gd_sketch_ref_batch_device(gd_ctx *ctx, int n, const int64_t *d_off, const int32_t *d_len, const uint32_t *d_rid,
                               const char *d_buf, int64_t total_len, int w, int k, const char *Z, int W,
                               int64_t *d_out_off, mm128_t *d_out, int64_t out_cap);

/* Read sketching for the mapping pipeline: for every read the two calls the reference makes per
 * read (GDiet-ShortReads/map.c:74-99), in one launch:
 *   mm_sketch2(max_seeds)       -> s2_counts[i*W + shift], entries s2[s2_off[i] .. s2_off[i+1])
 *   mm_sketch3(shift, cap) for EVERY shift 0..W-1
 *                               -> s3[s3_off[i*W+shift] .. s3_off[i*W+shift+1]), s3_ret[i*W+shift]
 * so the host can pick the shift (mm_get_shift) without a second device round trip. */
int gd_sketch_reads_batch(gd_ctx *ctx, int n, const int64_t *off, const int32_t *len, const char *buf, int w, int k,
                          const char *Z, int W, float max_seeds, uint32_t max_nb_seeds, uint32_t *s2_counts,
                          int64_t *s2_off, mm128_t *s2, int64_t s2_cap, int64_t *s3_off, uint32_t *s3_ret,
                          mm128_t *s3, int64_t s3_cap);

/* ------------------------------------------------------------------------------------------ */
/* (3) the stages either side of the two kernels: device-resident index + short-read mapping   */
/*     (SURVEY.md section 8 rows F1 "index lookup" and F2 "seed-hit sort + location voting")    */
/* ------------------------------------------------------------------------------------------ */

/* Device-resident minimizer index: what mm_idx_gen builds (GDiet-ShortReads/index.c:389-420) -- every
 * contig sketched with mm_sketch, the (minimizer, position) records grouped by minimizer with the
 * positions of one minimizer ascending (index.c:216-271), the 4-bit packed reference mi->S
 * (index.c:351-356) -- laid out for the GPU: one open-addressing table {minimizer -> first, count} over a
 * flat position array instead of 2^14 khash buckets.  A lookup returns exactly what mm_idx_get returns
 * (index.c:84-100): the count and the positions in ascending order. */
typedef struct gd_index gd_index;

/* Sketches n_seq contigs (ASCII, host buffers; contig i is rid i) on the device and builds the index
 * there.  (The reference's bucket_bits only shapes its host hash tables; no lookup result depends on it.) */
int gd_index_build(gd_ctx *ctx, int n_seq, const int64_t *off, const int32_t *len, const char *buf, int w, int k,
                   const char *Z, int W, gd_index **out);
/* Same with the ASCII sequence already in HBM (d_buf is a device pointer and is not modified; off / len are host arrays). */
int gd_index_build_device(gd_ctx *ctx, int n_seq, const int64_t *off, const int32_t *len, const char *d_buf, int w, int k,
                          const char *Z, int W, gd_index **out);
void gd_index_destroy(gd_index *idx);
/* "n_seq", "total_len", "n_minimizers" (records), "n_keys" (distinct minimizers), "table_slots",
 * "device_bytes", "s_words" (uint32 words of the 4-bit reference) */
int64_t gd_index_stat(const gd_index *idx, const char *key);
/* mm_idx_get for a batch (host arrays): count[i] = occurrences of minier[i] (= mm128_t.x >> 8),
 * first[i] = offset of its first position in the exported position array (-1 if absent). */
int gd_index_get_batch(gd_ctx *ctx, const gd_index *idx, int64_t n, const uint64_t *minier, uint32_t *count,
                       int64_t *first);
/* Copies the index to host arrays (any pointer may be NULL): keys[n_keys] ascending, counts[n_keys],
 * positions[n_minimizers] grouped by key in key order (each group ascending), S[s_words] = mi->S. */
int gd_index_export(gd_ctx *ctx, const gd_index *idx, uint64_t *keys, uint32_t *counts, uint64_t *positions,
                    uint32_t *S);
/* mm_idx_cal_max_occ (GDiet-ShortReads/index.c:182-201): the value mm_mapopt_update stores in mid_occ. */
int gd_index_cal_max_occ(gd_ctx *ctx, const gd_index *idx, float frac, int32_t *max_occ);

/* Replication onto the other GPUs of a node: the index is built once; the caller broadcasts
 * its GD_INDEX_NBUF device buffers with its own communicator (ncclBroadcast over NVLink) into an index allocated
 * with the same meta data on every other GPU, then commits it.  Buffers: table, positions, 4-bit reference,
 * contig offsets, contig lengths, distinct minimizers, their counts. */
#define GD_INDEX_NBUF 7
typedef struct {
	int64_t n_seq, total_len, n_minimizers, n_keys, table_slots, s_words;
	int32_t w, k;
} gd_index_meta_t;
int gd_index_meta(const gd_index *idx, gd_index_meta_t *meta);
int gd_index_alloc(gd_ctx *ctx, const gd_index_meta_t *meta, gd_index **out);
int gd_index_buffers(gd_index *idx, void **ptrs /* [GD_INDEX_NBUF] */, size_t *bytes /* [GD_INDEX_NBUF] */);
int gd_index_commit(gd_ctx *ctx, gd_index *idx);

/* The mm_mapopt_t fields the short-read path reads between mm_sketch2 and ksw_extd2
 * (GDiet-ShortReads/minimap.h:142-205; defaults main.c:166-182, preset options.c:130-150). */
typedef struct {
	int32_t W;     /* pattern length (-W) */
	char Z[64];    /* pattern (-Z), NUL padded */
	float max_seeds; /* opt->max_seeds (mm_sketch2) */
	int32_t frag_mode, max_frag_len; /* MM_F_FRAG_MODE, opt->max_frag_len: cap of mm_sketch3 (map.c:621-622) */
	uint32_t bw;   /* band width / vote distance of every read of the batch (used when bw_max == 0) */
	float min_cnt, rec_threshold_frac; /* -n */
	int32_t af_max_loc;                /* --AF_max_loc (<= 32) */
	int32_t mid_occ, max_max_occ, occ_dist;
	float q_occ_frac;
	int32_t for_only, rev_only; /* MM_F_FOR_ONLY / MM_F_REV_ONLY (map.c:121-127) */
	int32_t a, b, q, e, q2, e2, zdrop, end_bonus;
	/* -r bw_frac,bw_min,bw_max: with bw_max > 0 the band and vote distance are computed PER READ as at map.c:624-631,
	 * bw = (unsigned)(len * bw_frac) clamped to [bw_min, bw_max], and `bw` above is ignored */
	float bw_frac;
	uint32_t bw_min, bw_max;
} gd_sr_opt_t;

/* One candidate location of one read, in the order of the reference's candidate loop (map.c:764), with
 * the outcome of exact_match_sse / ksw_extd2 (flag KSW_EZ_APPROX_MAX) for it: everything mm_map_frag has
 * in hand when it calls mm_update_extra (map.c:932-954). 64 bytes. */
typedef struct {
	int32_t rid, rs, re, qs, qe, rev; /* mm_reg1_t fields set at map.c:932-938 */
	int32_t votes, first_q, last_q;   /* vt_t (map.c:433-440) */
	int32_t exact;                    /* exact_match_sse said equal: no DP, CIGAR = <len>M (map.c:873-915) */
	int32_t score, n_cigar;           /* ez.score, ez.n_cigar */
	int32_t cigar_off;                /* first entry of this candidate in cigar[] */
	int32_t reserved[3];
} gd_sr_cand_t;

/* mm_map_frag (GDiet-ShortReads/map.c:586-952) for a batch of single-segment short reads, from the
 * read's ASCII up to the ksw_extz_t of every candidate: mm_sketch2 -> mm_get_shift -> mm_sketch3 ->
 * mm_seed_mz_flt -> mm_collect_matches2 -> collect_seed_hits -> vote x2 -> window arithmetic ->
 * exact_match_sse | ksw_extd2.  Read i is buf[off[i] .. off[i]+len[i]).  Results in input order:
 *   cand[cand_off[i] .. cand_off[i+1])   candidates of read i
 *   cigar[c.cigar_off .. +c.n_cigar)     BAM CIGAR of candidate c
 * Returns GD_ERR_CAPACITY (with cand_off[n] / *n_cigar = required sizes) if an output is too small. */
int gd_sr_map_batch(gd_ctx *ctx, const gd_index *idx, int n, const int64_t *off, const int32_t *len, const char *buf,
                    const gd_sr_opt_t *opt, int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar,
                    int64_t cigar_cap, int64_t *n_cigar);

/* Long-read tree (GDiet-LongReads/map.c:1273-1853; defaults LR/main.c:170-182, presets LR/options.c:86-111): the
 * mm_mapopt_t fields read between mm_sketch2 and ksw_extd2.  bw is opt->bw (-r); mid_occ is what mm_mapopt_update
 * derives from the index (gd_index_cal_max_occ clamped to [min_mid_occ, max_mid_occ], LR/options.c:62-71). */
typedef struct {
	int32_t W;
	char Z[64];
	float max_seeds;
	int32_t frag_mode, max_frag_len;
	uint32_t bw;
	int32_t mid_occ, max_max_occ, occ_dist;
	float q_occ_frac;
	int32_t for_only, rev_only;
	int32_t a, b, q, e, q2, e2, zdrop, end_bonus;
	uint32_t vt_dis, vt_nb_loc;          /* --vt_dis, --vt_nb_loc (<= 32) */
	float vt_cov, vt_df1, vt_df2, vt_f;  /* --vt_cov, --vt_df1, --vt_df2, --vt_f */
	uint32_t max_max_gap, max_min_gap;   /* --max_max_gap, --max_min_gap */
} gd_lr_opt_t;

/* mm_map_frag of the long-read tree for a batch of reads, from the read's ASCII up to the ksw_extz_t of every
 * candidate (everything before mm_update_extra / concatenate_cigars, LR/map.c:1293-1805): both sketch calls,
 * mm_get_shift, seed filters, collect_seed_hits with the merge-sort hit order (--sort=merge, the default of the
 * long-read presets), vote x2, density and score filters, vote_2 x2 on the uncovered read ends, candidate chaining,
 * window arithmetic, ksw_extd2 (flag KSW_EZ_APPROX_MAX, band bw).  Output as gd_sr_map_batch, candidates in the
 * order of the loop at LR/map.c:1654; cand.reserved[0] = index (within the read) of the candidate that continues
 * this one (vt_t::next, -1 = none), cand.reserved[1] = vt_t::concat; score == KSW_NEG_INF marks the candidates the
 * reference drops at LR/map.c:1812. */
int gd_lr_map_batch(gd_ctx *ctx, const gd_index *idx, int n, const int64_t *off, const int32_t *len, const char *buf,
                    const gd_lr_opt_t *opt, int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar,
                    int64_t cigar_cap, int64_t *n_cigar);

/* ------------------------------------------------------------------------------------------ */
/* (4) host side after the DP -- plain C++, no GPU work            */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
	int32_t a, b, q, e;   /* scoring (mm_update_extra, map.c:954) */
	int32_t min_dp_max;   /* opt->min_dp_max (map.c:961) */
	int32_t best_n;       /* opt->best_n (map.c:980) */
	int32_t no_print_2nd; /* MM_F_NO_PRINT_2ND */
	int32_t is_sr;        /* MM_F_SR: linear instead of log gap cost in dp_max (map.c:954) */
	int32_t sam_hit_only; /* MM_F_SAM_HIT_ONLY: no record for unmapped reads */
	int32_t softclip;     /* MM_F_SOFTCLIP */
	int32_t n_threads;    /* host threads (0 = all cores) */
	int32_t q2, e2;       /* second gap piece: the junction gaps of concatenate_cigars (long reads, LR/map.c:1864-1866) */
} gd_sr_post_opt_t;

/* For every read of a batch: mm_update_extra (+ mm_fix_cigar), the clip / min_dp_max filter and the ordering of
 * map.c:956-978, mm_set_sam_params (hit.c:494-557) and mm_write_sam3 (format.c:412-603, single-segment reads):
 * the SAM records of the batch, in input order, as one malloc'ed text (free with gd_free).
 * names[i] / seq / qual (qual may be NULL) describe the reads like gd_sr_map_batch's off/len/buf; cand_off, cand,
 * cigar are its outputs; seq_names / ref_off / ref_len / ref describe the contigs (ASCII) the index was built from. */
int gd_sr_sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                    const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int n_seq,
                    const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                    const gd_sr_post_opt_t *opt, char **sam, size_t *sam_len);
/* Same records, handed over as the pieces the worker threads produced (in input order) instead of one concatenated
 * text, so that a host that writes them out (fwrite / writev, map.c:1208-1256) saves the copy: parts[0..n_parts) with
 * part_len[] bytes each; free every piece and both arrays with gd_free. */
int gd_sr_sam_batch_parts(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                          const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int n_seq,
                          const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                          const gd_sr_post_opt_t *opt, char ***parts, size_t **part_len, int *n_parts);
/* Reads in, SAM records out: gd_sr_map_batch followed ON THE DEVICE by what gd_sr_sam_batch does on the host
 * (mm_update_extra + mm_fix_cigar, the filter and ordering of map.c:956-978, mm_set_sam_params, mm_write_sam3 for
 * single-segment reads; post->is_sr must be set: linear gap cost).  One thread per read runs csrc/gd_sam_core.h; only the
 * text crosses PCIe.  The pieces (input order) live in pinned buffers OWNED BY THE CONTEXT and stay valid until the call
 * after the next one on this context; free only the two arrays with gd_free.  Byte-identical to gd_sr_sam_batch. */
int gd_sr_map_sam_batch(gd_ctx *ctx, const gd_index *idx, int n, const char *const *names, const int64_t *off, const int32_t *len,
                        const char *seq, const char *qual, const gd_sr_opt_t *opt, const gd_sr_post_opt_t *post, int n_seq,
                        const char *const *seq_names, char ***parts, size_t **part_len, int *n_parts);

/* The same for the long-read tree (LR/map.c:1807-1912): candidates with score == KSW_NEG_INF are dropped,
 * mm_update_extra uses the logarithmic gap cost, a valid candidate that is continued by a valid candidate
 * (cand.reserved[0] >= 0) absorbs it (concatenate_cigars, LR/map.c:41-640), then the min_dp_max filter and the ordering.
 * sam_off[0..n] (optional) gives every read's byte range in the text; needs_stitch (optional) is kept for callers of the
 * earlier interface and is always 0 now. */
int gd_lr_sam_batch(int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                    const char *qual, const int64_t *cand_off, const gd_sr_cand_t *cand, const uint32_t *cigar, int n_seq,
                    const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                    const gd_sr_post_opt_t *opt, char **sam, size_t *sam_len, int64_t *sam_off, uint8_t *needs_stitch);
/* the @SQ lines of mm_write_sam_hdr (format.c:128-137); the @PG line (command line) is the caller's */
int gd_sam_header(int n_seq, const char *const *seq_names, const int32_t *ref_len, char **sam, size_t *sam_len);
void gd_free(void *p);

/* Row F4: the reference's `.mmi` index file (mm_idx_dump, GDiet-ShortReads/index.c:480-517) from the arrays
 * gd_index_export returns -- byte for byte what `GDiet_avx -d` writes, including the slot order of the per-bucket
 * khash tables (rebuilt like worker_post, index.c:216-271).  bucket_bits = mm_idxopt_t::bucket_bits (14), flag =
 * mm_idx_t::flag (0 without -H). */
int gd_mmi_write(const char *path, int w, int k, int bucket_bits, int flag, int n_seq, const char *const *names,
                 const int32_t *lens, int64_t n_keys, const uint64_t *keys, const uint32_t *counts, const uint64_t *positions,
                 const uint32_t *S);
/* ... and the way back (mm_idx_load, index.c:519-571): a `.mmi` file written by the reference (`-d`) or by gd_mmi_write
 * becomes a device-resident index; gd_index_seq_name returns the contig names stored in the file ("" for built indexes). */
int gd_index_load_mmi(gd_ctx *ctx, const char *path, gd_index **out);
const char *gd_index_seq_name(const gd_index *idx, int i);

/* ------------------------------------------------------------------------------------------ */
/* (5) several GPUs of one box from one host process                     */
/* ------------------------------------------------------------------------------------------ */
/* What kt_for does over host cores (GDiet-ShortReads/map.c:1045-1092,1206) with the order kt_pipeline restores
 * (kthread.c:101-115): one context per device, the index built or loaded ONCE on the first device and broadcast to the
 * others, every mini-batch cut into contiguous read shards (one per device, equal bases), each shard mapped by its own
 * host thread, results handed back in input order.  No data-path collective besides the broadcast. */
typedef struct gd_multi gd_multi;
int gd_multi_init(int n_dev, const int *devices /* NULL: 0..n_dev-1 */, gd_multi **out);
void gd_multi_destroy(gd_multi *m);
int gd_multi_size(const gd_multi *m);
gd_ctx *gd_multi_ctx(gd_multi *m, int i);              /* context of device i (build / load the index on i = 0) */
const gd_index *gd_multi_index(const gd_multi *m, int i);
const char *gd_multi_strerror(const gd_multi *m);
/* root lives on gd_multi_ctx(m, 0); replicas are created on every other device and filled with ncclBroadcast over NVLink
 * (libnccl.so.2 is dlopen'ed at run time -- no link-time dependency; GDIET_NO_NCCL=1 or a missing library selects direct
 * peer copies, cudaMemcpyPeerAsync).  take_ownership != 0: gd_multi_destroy frees root as well.
 * gd_multi_stat: "bcast_seconds", "bcast_bytes", "bcast_path" (1 = NCCL, 2 = peer copies) of the last broadcast. */
int gd_multi_index_bcast(gd_multi *m, gd_index *root, int take_ownership);
double gd_multi_stat(const gd_multi *m, const char *key);
/* Optional, ahead of the first gd_multi_sr_map_sam: create the lanes of every device and page-lock the SAM text buffers for
 * calls that write about text_bytes each (a one-shot host runs this beside its index build; otherwise the first two calls pay).
 * gd_sr_map_sam_prepare is the same for one context. */
int gd_multi_prepare_sam(gd_multi *m, size_t text_bytes);
int gd_sr_map_sam_prepare(gd_ctx *ctx, size_t text_bytes);
/* gd_sr_map_batch / gd_lr_map_batch over all devices: same arguments, same results (input order, dense arrays) */
int gd_multi_sr_map_batch(gd_multi *m, int n, const int64_t *off, const int32_t *len, const char *buf, const gd_sr_opt_t *opt,
                          int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap,
                          int64_t *n_cigar);
int gd_multi_lr_map_batch(gd_multi *m, int n, const int64_t *off, const int32_t *len, const char *buf, const gd_lr_opt_t *opt,
                          int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap,
                          int64_t *n_cigar);
/* mapping + the post-DP stage per shard: the SAM records of the mini-batch as text pieces in input order (what pipeline
 * step 2, map.c:1208-1256, writes out).  Short reads: gd_sr_map_sam_batch on every device (the text is made on the GPU);
 * long reads: gd_lr_map_batch + gd_lr_sam_batch.  The pieces BELONG TO THE HANDLE and stay valid until the call after the
 * next one; free only the two arrays with gd_free.  ref_off / ref_len / ref are read by the long-read host stage only. */
int gd_multi_sr_map_sam(gd_multi *m, int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                        const char *qual, const gd_sr_opt_t *opt, const gd_sr_post_opt_t *post, int n_seq,
                        const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                        char ***parts, size_t **part_len, int *n_parts);
int gd_multi_lr_map_sam(gd_multi *m, int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                        const char *qual, const gd_lr_opt_t *opt, const gd_sr_post_opt_t *post, int n_seq,
                        const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                        char ***parts, size_t **part_len, int *n_parts);

#ifdef __cplusplus
}
#endif
#endif
