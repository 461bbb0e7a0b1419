/* gd_batched_host.c -- the batched host of INTEGRATION.md level 2: ONE file that, compiled next to the UNMODIFIED
 * reference sources of either tree (GDiet-ShortReads or, with -DGD_HOST_LR, GDiet-LongReads) and linked against
 * libgdiet_cuda.so, replaces the mapping pipeline of map.c:
 *
 *   mm_map_file_frag()    map.c:1301-1322 (LR/map.c:2185-2206)   -- this file provides the symbol; map.c is compiled
 *                                                                    with -Dmm_map_file_frag=gdref_cpu_mm_map_file_frag
 *   worker_pipeline()     map.c:1165-1281   step 0 read a mini-batch, step 1 map it, step 2 write it out
 *   worker_for()          map.c:1045-1092   one mm_map_frag per read on -t host threads
 *
 * Here step 0 reads the mini-batch with the reference's own reader (mm_bseq_read3, bseq.c) and flattens it into one
 * pinned buffer; step 1 hands the WHOLE mini-batch to the device path on every GPU named by GDIET_GPUS (gd_multi_*:
 * contiguous read shards, one per device, SAM pieces handed back in input order) -- sketching, index lookups, voting,
 * windows, ksw_extd2, CIGARs and, for short reads, the post-DP phase too (mm_update_extra ... mm_write_sam3 run on the
 * device, gd_sr_map_sam_batch: only SAM text comes back); the long-read tree finishes with the library's threaded host
 * restatement of that phase (gd_lr_sam_batch); step 2 writes the pieces.  The three steps run under the
 * reference's kt_pipeline (kthread.c:71-160), so reading batch i+1, mapping batch i and writing batch i-1 overlap
 * exactly as in the reference, and the output order is the input order.
 *
 * and, on the index side (index.c is compiled with these four names renamed away in the same manner):
 *
 *   mm_idx_reader_read()  index.c:624-640   a FASTA reference is read with the reference's reader, but the minimizer index
 *                                           is built ON THE DEVICE (gd_index_build, ~1 s for 3.1 Gbp) instead of mm_idx_gen's
 *                                           sketch + sort + 2^b host hash tables (19 s on 16 threads), which the device path
 *                                           never looks at; the mm_idx_t that main.c gets holds names, lengths, w, k, flags
 *   mm_idx_cal_max_occ()  index.c:190-210   answered by the device index (mm_mapopt_update calls it for mid_occ)
 *   mm_idx_stat(), mm_idx_destroy()         the statistics line comes from the device index; destroy also drops the device side
 *   (an .mmi file, -d, and homopolymer-compressed indexing keep the reference's code: the host tables ARE the file format)
 *
 * Everything else is the reference's: option parsing (main.c), the SAM header (format.c), FASTA/FASTQ parsing of gzip / stdin
 * input (bseq.c).  With an index that came from an .mmi the device index is built from the sequences it holds (mm_idx_getseq).
 *
 * What the device path does not cover is refused with an error, never mapped differently: paired / multi-segment
 * input, PAF output, --split-prefix, --cs / --MD / --eqx / -y, splice mode, and (long reads) --sort=radix|heap.
 */
#include <errno.h>
#include <pthread.h>
#include <fcntl.h>
#include <stdio.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include "minimap.h"
#include "mmpriv.h"
#include "bseq.h"
#include "kthread.h"
#include "kalloc.h"
#include "gdiet_cuda.h" /* after minimap.h: its guards skip the types the reference already defines */

typedef struct {
	const mm_mapopt_t *opt;
	const mm_idx_t *mi;
	mm_bseq_file_t *fp;        /* the reference's reader (gzip / stdin input) */
	const char *map;           /* or: the whole plain-text file mapped read-only, parsed in place */
	size_t map_len, map_pos;
	int map_last;              /* kseq's last_char: a header character already consumed (0 = none) */
	int n_threads, n_processed;
	int64_t mini_batch_size;
	int plain_buffers;    /* the batch's bases go to plain memory (the reference reader of the index side), not to pinned buffers */
	size_t par_min_bytes; /* the several-thread reader needs at least this many bytes per batch (smaller batches: one thread) */
	long n_par_batches;   /* mini-batches read by the several-thread reader */
	int64_t sub_bases; /* long reads: a mini-batch also ends at >= sub_bases bases and >= sub_reads reads (0: only at -K) */
	int sub_reads;
	gd_multi *gm;
	/* contigs as ASCII (the host stage counts mismatches against them, align.c:259-318) */
	int n_ref;
	const char **ref_names;
	int64_t *ref_off;
	int32_t *ref_len;
	char *ref;
#ifdef GD_HOST_LR
	gd_lr_opt_t mo;
#else
	gd_sr_opt_t mo;
#endif
	gd_sr_post_opt_t po;
	double t_map, t_read, t_write;
	int64_t n_reads, n_bases;
	/* The text pieces of a mini-batch stay valid until the call after the next one (two alternating buffers per device), and
	 * kt_pipeline lets batch b+2 enter the mapping step as soon as batch b+1 has LEFT it -- batch b may still be in the middle of
	 * its write.  So the mapping step of batch b waits until batch b-2 has been written. */
	pthread_mutex_t order_mu;
	pthread_cond_t order_cv;
	long n_batches, n_written;
} gdh_pipeline_t;

typedef struct {
	int n;
	mm_bseq1_t *seq;
	const char **names;
	char *name_blob;           /* mapped-file reader: the names of the batch, NUL separated */
	int64_t *off;
	int32_t *len;
	char *buf, *qual; /* flattened reads / qualities */
	size_t buf_cap, qual_cap;
	char **parts;
	size_t *part_len;
	int n_parts;
	long batch; /* 0, 1, 2 ... in input order */
} gdh_step_t;

/* Page-locking memory is slow (tens of ms per 100 MB): the read buffers of the mini-batches in flight (at most three under
 * kt_pipeline, a read and a quality buffer each) come from a small pool instead of a fresh gd_pinned_alloc per batch. */
#include <pthread.h>
static struct {
	pthread_mutex_t mu;
	char *p[8];
	size_t cap[8];
} gdh_pool = {PTHREAD_MUTEX_INITIALIZER, {0}, {0}};
static char *gdh_buf_get(size_t bytes, size_t *cap)
{
	int i, best = -1;
	char *p = 0;
	pthread_mutex_lock(&gdh_pool.mu);
	for (i = 0; i < 8; ++i)
		if (gdh_pool.p[i] && gdh_pool.cap[i] >= bytes && (best < 0 || gdh_pool.cap[i] < gdh_pool.cap[best])) best = i;
	if (best >= 0) p = gdh_pool.p[best], *cap = gdh_pool.cap[best], gdh_pool.p[best] = 0;
	pthread_mutex_unlock(&gdh_pool.mu);
	if (p) return p;
	*cap = bytes + bytes / 4 + 4096;
	return (char *)gd_pinned_alloc(*cap);
}
static void gdh_buf_put(char *p, size_t cap)
{
	int i;
	pthread_mutex_lock(&gdh_pool.mu);
	for (i = 0; i < 8 && p; ++i)
		if (!gdh_pool.p[i]) gdh_pool.p[i] = p, gdh_pool.cap[i] = cap, p = 0;
	pthread_mutex_unlock(&gdh_pool.mu);
	if (p) gd_pinned_free(p);
}

#ifdef GD_HOST_LR /* the long-read host stage reads the qualities on the CPU only: plain memory, nothing to page-lock */
#define GDH_QUAL_PINNED 0
static char *gdh_qbuf_get(size_t bytes, size_t *cap) { *cap = bytes; return (char *)malloc(bytes); }
static void gdh_qbuf_put(char *p, size_t cap) { (void)cap; free(p); }
#else /* short reads: the device SAM stage reads them over PCIe */
#define GDH_QUAL_PINNED 1
#define gdh_qbuf_get gdh_buf_get
#define gdh_qbuf_put gdh_buf_put
#endif

static void gdh_die(const char *what)
{
	fprintf(stderr, "[gdiet_cuda] ERROR: %s\n", what);
	exit(1);
}

/* ---- mini-batch reader over a memory-mapped plain-text FASTA/FASTQ file -------------------------------------------------
 * The reference reads with kseq.h over gzread: three mallocs and a byte-wise state machine per record, about 1.2 M short
 * reads/s on one thread -- which bounds the whole pipeline once mapping runs on a GPU.  This reader applies the SAME rules
 * (kseq_read, kseq.h:193-232; mm_bseq_read3, bseq.c:73-113; kseq2bseq, bseq.c:58-71) to the mapped file with memchr and
 * writes bases and qualities straight into the pinned mini-batch buffers:
 *   a record starts at the next '>' or '@'; the name ends at the first white space; sequence lines run until a line that
 *   starts with '>', '+' or '@' (empty lines skipped, a trailing CR dropped); after '+' quality lines are appended until
 *   they are as long as the sequence; U/u become T/t; a batch ends with the record that brings it to >= -K bases.
 * A record whose quality length differs from its sequence length ends the batch like kseq's -2 (with the reference's
 * warning).  gzip or stdin input keeps the reference's reader. */
typedef struct {
	char *p;
	size_t n, cap;
} gdh_str_t;
static void gdh_str_need(gdh_str_t *s, size_t extra, int pinned, size_t *pool_cap)
{
	if (s->n + extra + 1 <= s->cap) return;
	{
		size_t want = (s->n + extra + 1) * 2;
		char *np;
		if (pinned) {
			size_t cap2 = 0;
			np = gdh_buf_get(want, &cap2);
			if (!np) { fprintf(stderr, "[gdiet_cuda] ERROR: out of page-locked memory\n"); exit(1); }
			if (s->n) memcpy(np, s->p, s->n);
			if (s->p) gdh_buf_put(s->p, *pool_cap);
			*pool_cap = cap2, s->cap = cap2;
		} else {
			np = (char *)realloc(s->p, want);
			s->cap = want;
		}
		s->p = np;
	}
}
static inline int gdh_isspace(int c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

/* appends the rest of the current line (without the newline, a trailing CR dropped when the string is longer than one
 * character: ks_getuntil2, kseq.h:146) to dst; returns 0 at end of file with nothing read */
static int gdh_line(gdh_pipeline_t *p, gdh_str_t *dst, size_t base, int pinned, size_t *pool_cap)
{ /* base: where the current record's string starts in dst (the CR rule looks at that string's length) */
	const char *b = p->map + p->map_pos, *e;
	size_t l;
	if (p->map_pos >= p->map_len) return 0;
	e = (const char *)memchr(b, '\n', p->map_len - p->map_pos);
	l = e ? (size_t)(e - b) : p->map_len - p->map_pos;
	gdh_str_need(dst, l, pinned, pool_cap);
	memcpy(dst->p + dst->n, b, l), dst->n += l;
	p->map_pos += l + (e ? 1 : 0);
	if (dst->n - base > 1 && dst->p[dst->n - 1] == '\r') --dst->n;
	return 1;
}

static int gdh_read_fastq_par(gdh_pipeline_t *p, gdh_step_t *s, int with_qual);

static int gdh_read_mapped(gdh_pipeline_t *p, gdh_step_t *s, int with_qual)
{
	gdh_str_t seq = {0, 0, 0}, qual = {0, 0, 0}, names = {0, 0, 0};
	size_t seq_pool = 0, qual_pool = 0, dummy = 0;
	int64_t size = 0, *off = 0;
	int32_t *len = 0;
	int64_t *noff = 0;
	int n = 0, m = 0, any_qual = 0, all_qual = 1, bad = 0, i;
	const int64_t left = (int64_t)(p->map_len - p->map_pos);
	int64_t want = p->mini_batch_size < left ? p->mini_batch_size : left; /* bases of the batch are at most the bytes left in the file */
	size_t guess;
	if (p->sub_bases > 0 && p->sub_bases + p->sub_bases / 4 < want) want = p->sub_bases + p->sub_bases / 4; /* the usual end of a long-read batch */
	guess = (size_t)want + 65536;
	if (!p->plain_buffers && (i = gdh_read_fastq_par(p, s, with_qual)) >= 0) return i; /* strict four-line FASTQ: several threads */
	gdh_str_need(&seq, guess, !p->plain_buffers, &seq_pool);
	if (with_qual) gdh_str_need(&qual, guess, GDH_QUAL_PINNED, &qual_pool);
	while (1) {
		size_t l0, q0;
		int c = p->map_last;
		if (c == 0) { /* jump to the next header line (kseq.h:197-201) */
			const char *b = p->map + p->map_pos, *e = p->map + p->map_len;
			while (b < e && *b != '>' && *b != '@') ++b;
			if (b >= e) { p->map_pos = p->map_len; break; }
			p->map_pos = (size_t)(b - p->map) + 1;
		}
		p->map_last = 0;
		if (p->map_pos >= p->map_len) break; /* ks_getuntil on an exhausted stream: normal exit */
		if (n == m) {
			m = m ? m * 2 : 4096;
			off = (int64_t *)realloc(off, sizeof(int64_t) * m), len = (int32_t *)realloc(len, sizeof(int32_t) * m);
			noff = (int64_t *)realloc(noff, sizeof(int64_t) * (m + 1));
		}
		{ /* name: up to the first white space; the rest of the header line is the comment (not copied) */
			const char *b = p->map + p->map_pos, *e = p->map + p->map_len, *q = b;
			while (q < e && !gdh_isspace((unsigned char)*q)) ++q;
			if (q == b) fprintf(stderr, "[WARNING]\033[1;31m empty sequence name in the input.\033[0m\n");
			gdh_str_need(&names, (size_t)(q - b) + 1, 0, &dummy);
			noff[n] = (int64_t)names.n;
			memcpy(names.p + names.n, b, (size_t)(q - b)), names.n += (size_t)(q - b), names.p[names.n++] = 0;
			p->map_pos = (size_t)(q - p->map) + (q < e ? 1 : 0);
			if (q < e && *q != '\n') { /* skip the comment */
				const char *nl = (const char *)memchr(q, '\n', (size_t)(e - q));
				p->map_pos = nl ? (size_t)(nl - p->map) + 1 : p->map_len;
			}
		}
		l0 = seq.n, q0 = qual.n;
		c = -1;
		while (p->map_pos < p->map_len) { /* sequence lines (kseq.h:209-213) */
			c = (unsigned char)p->map[p->map_pos++];
			if (c == '>' || c == '+' || c == '@') break;
			if (c == '\n') { c = -1; continue; }
			--p->map_pos; /* the character belongs to the line: first character + rest of the line, as kseq appends them */
			gdh_line(p, &seq, l0, !p->plain_buffers, &seq_pool);
			c = -1;
		}
		if (c == '>' || c == '@') p->map_last = c;
		off[n] = (int64_t)l0, len[n] = (int32_t)(seq.n - l0);
		for (i = 0; i < len[n]; ++i) /* U -> T, bseq.c:66-68 */
			if (seq.p[l0 + i] == 'u' || seq.p[l0 + i] == 'U') --seq.p[l0 + i];
		if (c == '+') { /* FASTQ: skip the '+' line, then quality lines until they cover the sequence (kseq.h:226-230) */
			const char *nl = p->map_pos < p->map_len ? (const char *)memchr(p->map + p->map_pos, '\n', p->map_len - p->map_pos) : 0;
			gdh_str_t *qd = with_qual ? &qual : 0;
			gdh_str_t scratch = {0, 0, 0};
			if (!nl) { bad = 1; break; } /* no quality string */
			p->map_pos = (size_t)(nl - p->map) + 1;
			if (!qd) qd = &scratch;
			{
				const size_t qstart = qd->n;
				while (gdh_line(p, qd, qstart, qd == &qual && GDH_QUAL_PINNED, &qual_pool) && qd->n - qstart < (size_t)len[n]);
				if (qd->n - qstart != (size_t)len[n]) bad = 1;
			}
			free(scratch.p);
			if (bad) break;
			any_qual = 1;
		} else all_qual = 0;
		if (with_qual && qual.n - q0 != (size_t)len[n]) { /* a FASTA record among FASTQ ones: keep the two buffers in step */
			gdh_str_need(&qual, (size_t)len[n], GDH_QUAL_PINNED, &qual_pool);
			qual.n = q0 + (size_t)len[n];
		}
		size += len[n], ++n;
		if (size >= p->mini_batch_size) break;
		if (p->sub_bases > 0 && size >= p->sub_bases && n >= p->sub_reads) break;
	}
	if (bad) { /* kseq returns -2: the record is dropped, the batch ends here, the next call resynchronises on a header */
		seq.n = n < m && off ? (size_t)off[n] : seq.n;
		if (n) fprintf(stderr, "[WARNING]\033[1;31m failed to parse the FASTA/FASTQ record next to '%s'. Continue anyway.\033[0m\n", names.p + noff[n - 1]);
		else fprintf(stderr, "[WARNING]\033[1;31m failed to parse the first FASTA/FASTQ record. Continue anyway.\033[0m\n");
		p->map_last = 0;
	}
	if (n == 0) {
		if (seq.p && !p->plain_buffers) gdh_buf_put(seq.p, seq_pool);
		else free(seq.p);
		if (qual.p) gdh_qbuf_put(qual.p, qual_pool);
		free(names.p), free(off), free(len), free(noff);
		return 0;
	}
	s->n = n, s->off = off, s->len = len, s->buf = seq.p, s->buf_cap = seq_pool;
	s->name_blob = names.p;
	s->names = (const char **)malloc(sizeof(char *) * n);
	for (i = 0; i < n; ++i) s->names[i] = names.p + noff[i];
	free(noff);
	if (with_qual && any_qual && all_qual) s->qual = qual.p, s->qual_cap = qual_pool; /* mm_write_sam3 prints '*' for reads without qualities */
	else if (qual.p) gdh_qbuf_put(qual.p, qual_pool);
	if (with_qual && any_qual && !all_qual)
		gdh_die("a mini-batch mixes FASTA and FASTQ records: not covered by the batched device path");
	return n;
}

/* ---- the same reader on several threads, for the common case: strict four-line FASTQ ------------------------------------
 * One thread parses 2.6 M short reads/s, the device stage maps five times that.  When the next batch is made of records of
 * exactly this shape
 *     @name[ comment]\n  SEQ\n  +[anything]\n  QUAL(\n | end of file)        SEQ not empty, not starting with '>', '+', '@',
 *                                                                          QUAL as long as SEQ, no CR before a newline,
 *                                                                          the next record right behind
 * the rules above reduce to "four lines per record" (kseq's line-joining, blank-line and resynchronisation rules never
 * fire), so the byte range of the batch is cut into one segment per thread at record starts, every thread VALIDATES and
 * measures its records (and must arrive exactly at the next segment's start), the batch end is found on the per-record
 * lengths by the same rule as above, and the threads copy names, bases and qualities to their final places.  Where a segment
 * starts: a line that starts with '@' and whose next-but-one line starts with '+' is a header -- were it a quality line, the
 * next-but-one line would be a sequence, which cannot start with '+' in a file of this shape; a wrong guess cannot survive
 * the arrive-exactly check.  Anything else (multi-line records, FASTA, CRLF, blank lines, a truncated tail ...) returns -1
 * before any state changes, and the one-thread reader above parses the batch by the full rules. */
typedef struct {
	size_t beg, end;               /* the records of the segment start at beg; the walk has to arrive exactly at end */
	int64_t *pos;                  /* per record: file offset of its '@' */
	int32_t *slen, *nlen;          /* per record: sequence and name length */
	int n, m, ok;
	int take;                      /* pass 2: the first `take` records belong to the batch */
	int64_t rec0, seq0, name0;     /* pass 2: index of the first record in the batch, offsets into the read and name buffers */
} gdh_seg_t;
typedef struct {
	const char *map;
	size_t len;
	gdh_seg_t *seg;
	char *seq, *qual, *names;
	int64_t *off;
	int32_t *rlen;
	const char **name_ptr;
} gdh_par_t;

/* the strict record at q: its sequence / name length and where the next record starts; 0 when the bytes at q are not of that shape */
static inline size_t gdh_strict_record(const char *map, size_t len, size_t q, int32_t *slen, int32_t *nlen)
{
	const char *b = map + q, *e = map + len, *l1, *l2, *l3, *l4, *t;
	size_t L;
	if (b >= e || *b != '@') return 0;
	if (!(l1 = (const char *)memchr(b, '\n', (size_t)(e - b)))) return 0;
	for (t = b + 1; t < l1 && !gdh_isspace((unsigned char)*t); ++t) {}
	if (t == b + 1) return 0; /* empty name: the full reader warns */
	if (l1 + 1 >= e || l1[1] == '>' || l1[1] == '+' || l1[1] == '@' || l1[1] == '\n') return 0;
	if (!(l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(e - (l1 + 1))))) return 0;
	if (l2[-1] == '\r' || l2 + 1 >= e || l2[1] != '+') return 0;
	if (!(l3 = (const char *)memchr(l2 + 1, '\n', (size_t)(e - (l2 + 1))))) return 0;
	L = (size_t)(l2 - (l1 + 1));
	if (L > 0x7fffffff || (size_t)(e - (l3 + 1)) < L) return 0;
	l4 = (const char *)memchr(l3 + 1, '\n', (size_t)(e - (l3 + 1)));
	if (l4 ? (size_t)(l4 - (l3 + 1)) != L : (size_t)(e - (l3 + 1)) != L) return 0; /* the quality line is exactly as long as the sequence */
	if (l3[L] == '\r') return 0; /* (l3 + 1)[L - 1] */
	*slen = (int32_t)L, *nlen = (int32_t)(t - (b + 1));
	if (!l4) return len;
	if (l4 + 1 < e && l4[1] != '@') return 0; /* a blank line, FASTA, garbage: the full rules decide */
	return (size_t)(l4 + 1 - map);
}

/* first record start at or behind `from` (see above); 0 if none within a few lines */
static size_t gdh_find_record_start(const char *map, size_t len, size_t from)
{
	const char *e = map + len, *ls = map + from;
	int tries;
	if (from > 0 && map[from - 1] != '\n') {
		if (!(ls = (const char *)memchr(ls, '\n', (size_t)(e - ls)))) return 0;
		++ls;
	}
	for (tries = 0; tries < 16 && ls < e; ++tries) {
		const char *l1 = (const char *)memchr(ls, '\n', (size_t)(e - ls)), *l2;
		if (!l1 || l1 + 1 >= e) return 0;
		if (*ls == '@' && (l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(e - (l1 + 1)))) && l2 + 1 < e && l2[1] == '+') return (size_t)(ls - map);
		ls = l1 + 1;
	}
	return 0;
}

static void gdh_par_scan(void *data, long i, int tid)
{
	gdh_par_t *P = (gdh_par_t *)data;
	gdh_seg_t *g = &P->seg[i];
	size_t q = g->beg;
	(void)tid;
	g->ok = 1;
	while (q < g->end) {
		int32_t sl, nl;
		const size_t nq = gdh_strict_record(P->map, P->len, q, &sl, &nl);
		if (!nq) { g->ok = 0; return; }
		if (g->n == g->m) {
			g->m = g->m ? g->m * 2 : 4096;
			g->pos = (int64_t *)realloc(g->pos, sizeof(int64_t) * g->m);
			g->slen = (int32_t *)realloc(g->slen, sizeof(int32_t) * g->m), g->nlen = (int32_t *)realloc(g->nlen, sizeof(int32_t) * g->m);
		}
		g->pos[g->n] = (int64_t)q, g->slen[g->n] = sl, g->nlen[g->n] = nl, ++g->n;
		q = nq;
	}
	if (q != g->end) g->ok = 0;
}

static void gdh_par_copy(void *data, long i, int tid)
{
	gdh_par_t *P = (gdh_par_t *)data;
	const gdh_seg_t *g = &P->seg[i];
	int64_t so = g->seq0, no = g->name0;
	int k, j;
	(void)tid;
	for (k = 0; k < g->take; ++k) {
		const char *b = P->map + g->pos[k];
		const int32_t L = g->slen[k], nl = g->nlen[k];
		const char *sq = (const char *)memchr(b + 1 + nl, '\n', (size_t)(P->map + P->len - (b + 1 + nl))) + 1; /* (the scan found this newline) */
		char *d = P->seq + so;
		memcpy(P->names + no, b + 1, (size_t)nl), P->names[no + nl] = 0;
		P->name_ptr[g->rec0 + k] = P->names + no;
		memcpy(d, sq, (size_t)L);
		for (j = 0; j < L; ++j) /* U -> T, bseq.c:66-68 */
			if (d[j] == 'u' || d[j] == 'U') --d[j];
		if (P->qual) {
			const char *pl = (const char *)memchr(sq + L + 1, '\n', (size_t)(P->map + P->len - (sq + L + 1))) + 1; /* behind the '+' line */
			memcpy(P->qual + so, pl, (size_t)L);
		}
		P->off[g->rec0 + k] = so, P->rlen[g->rec0 + k] = L;
		so += L, no += nl + 1;
	}
}

/* returns the number of reads, or -1 when this batch is not for the several-thread reader (nothing has changed then) */
static int gdh_read_fastq_par(gdh_pipeline_t *p, gdh_step_t *s, int with_qual)
{
	int T = p->n_threads < 16 ? p->n_threads : 16, i, k, n = 0, reached = 0;
	gdh_par_t P;
	gdh_seg_t seg[16];
	int32_t sl, nl;
	size_t n1, rend, next_pos = 0;
	int64_t budget, size = 0, name_bytes = 0;
	if (T < 2 || p->map_last != 0 || p->map_pos >= p->map_len || p->map[p->map_pos] != '@') return -1;
	if (!(n1 = gdh_strict_record(p->map, p->map_len, p->map_pos, &sl, &nl))) return -1;
	budget = p->mini_batch_size;
	if (p->sub_bases > 0) {
		int64_t b2 = (int64_t)p->sub_reads * sl > p->sub_bases ? (int64_t)p->sub_reads * sl : p->sub_bases;
		if (b2 < budget) budget = b2;
	}
	{ /* the byte range that should hold the batch: bytes per base of the first record, a quarter more, and some slack */
		const double bytes = (double)(n1 - p->map_pos) / (double)sl * (double)budget * 1.25 + 65536.0;
		if (bytes >= (double)(p->map_len - p->map_pos)) rend = p->map_len;
		else if (!(rend = gdh_find_record_start(p->map, p->map_len, p->map_pos + (size_t)bytes))) return -1;
	}
	if (rend - p->map_pos < p->par_min_bytes) return -1;
	memset(seg, 0, sizeof(seg)), memset(&P, 0, sizeof(P));
	P.map = p->map, P.len = p->map_len, P.seg = seg;
	for (i = 0; i < T; ++i) { /* segment starts: record starts near the equal cuts (a cut that finds none makes an empty segment) */
		size_t b = i == 0 ? p->map_pos : gdh_find_record_start(p->map, p->map_len, p->map_pos + (rend - p->map_pos) / (size_t)T * (size_t)i);
		if (i > 0 && (b == 0 || b > rend || b < seg[i - 1].beg)) b = seg[i - 1].beg;
		seg[i].beg = b;
		if (i > 0) seg[i - 1].end = b;
	}
	seg[T - 1].end = rend;
	kt_for(T, gdh_par_scan, &P, T);
	for (i = 0; i < T; ++i)
		if (!seg[i].ok) goto not_strict;
	for (i = 0; i < T && !reached; ++i) { /* the end of the batch, by the rule of the one-thread reader */
		seg[i].rec0 = n, seg[i].seq0 = size, seg[i].name0 = name_bytes;
		for (k = 0; k < seg[i].n && !reached; ++k) {
			size += seg[i].slen[k], name_bytes += seg[i].nlen[k] + 1, ++n;
			if (size >= p->mini_batch_size || (p->sub_bases > 0 && size >= p->sub_bases && n >= p->sub_reads)) reached = 1;
		}
		seg[i].take = k;
		if (reached) next_pos = k < seg[i].n ? (size_t)seg[i].pos[k] : seg[i].end;
	}
	if (!reached) {
		if (rend < p->map_len) goto not_strict; /* the estimate fell short (ragged records): the one-thread reader takes the batch */
		next_pos = p->map_len;
	}
	if (n == 0) goto not_strict;
	P.seq = gdh_buf_get((size_t)size + 16, &s->buf_cap);
	P.qual = with_qual ? gdh_qbuf_get((size_t)size + 16, &s->qual_cap) : 0;
	if (!P.seq || (with_qual && !P.qual)) gdh_die("out of page-locked memory");
	P.names = (char *)malloc((size_t)name_bytes + 1);
	P.off = (int64_t *)malloc(sizeof(int64_t) * n), P.rlen = (int32_t *)malloc(sizeof(int32_t) * n);
	P.name_ptr = (const char **)malloc(sizeof(char *) * n);
	kt_for(T, gdh_par_copy, &P, T);
	for (i = 0; i < T; ++i) free(seg[i].pos), free(seg[i].slen), free(seg[i].nlen);
	s->n = n, s->off = P.off, s->len = P.rlen, s->buf = P.seq, s->qual = P.qual, s->name_blob = P.names, s->names = P.name_ptr;
	p->map_pos = next_pos, p->map_last = 0, ++p->n_par_batches;
	return n;
not_strict:
	for (i = 0; i < T; ++i) free(seg[i].pos), free(seg[i].slen), free(seg[i].nlen);
	return -1;
}

static void *gdh_worker(void *shared, int step, void *in)
{
	gdh_pipeline_t *p = (gdh_pipeline_t *)shared;
	if (step == 0) { /* read a mini-batch (map.c:1168-1199) and flatten it */
		const int with_qual = !!(p->opt->flag & MM_F_OUT_SAM) && !(p->opt->flag & MM_F_NO_QUAL);
		double t0 = realtime();
		gdh_step_t *s = (gdh_step_t *)calloc(1, sizeof(gdh_step_t));
		int i;
		int64_t tot = 0, o = 0;
		if (p->map) {
			if (!gdh_read_mapped(p, s, with_qual)) {
				free(s);
				return 0;
			}
			for (i = 0; i < s->n; ++i) {
				tot += s->len[i];
				if (i > 0 && mm_qname_same(s->names[i - 1], s->names[i]))
					gdh_die("paired / multi-segment reads are not covered by the batched device path (use the GDiet_cuda build)");
			}
			p->n_processed += s->n, p->n_reads += s->n, p->n_bases += tot, p->t_read += realtime() - t0;
			s->batch = p->n_batches++;
			return s;
		}
		s->seq = mm_bseq_read3(p->fp, p->mini_batch_size, with_qual, 0, 0, &s->n);
		if (!s->seq) {
			free(s);
			return 0;
		}
		for (i = 0; i < s->n; ++i) s->seq[i].rid = p->n_processed++, tot += s->seq[i].l_seq;
		s->names = (const char **)malloc(sizeof(char *) * s->n);
		s->off = (int64_t *)malloc(sizeof(int64_t) * s->n), s->len = (int32_t *)malloc(sizeof(int32_t) * s->n);
		s->buf = gdh_buf_get((size_t)tot + 16, &s->buf_cap); /* pinned: the upload of step 1 runs at PCIe speed */
		if (!s->buf) gdh_die("out of page-locked memory");
		s->qual = with_qual ? gdh_qbuf_get((size_t)tot + 16, &s->qual_cap) : 0;
		for (i = 0; i < s->n && s->qual; ++i)
			if (!s->seq[i].qual) gdh_qbuf_put(s->qual, s->qual_cap), s->qual = 0; /* FASTA input: no quality strings ('*' in SAM) */
		for (i = 0; i < s->n; ++i) {
			const mm_bseq1_t *t = &s->seq[i];
			if (i > 0 && mm_qname_same(s->seq[i - 1].name, t->name))
				gdh_die("paired / multi-segment reads are not covered by the batched device path (use the GDiet_cuda build)");
			s->names[i] = t->name, s->off[i] = o, s->len[i] = t->l_seq;
			memcpy(s->buf + o, t->seq, t->l_seq);
			if (s->qual) memcpy(s->qual + o, t->qual, t->l_seq);
			o += t->l_seq;
		}
		p->n_reads += s->n, p->n_bases += tot, p->t_read += realtime() - t0;
		s->batch = p->n_batches++;
		return s;
	} else if (step == 1) { /* map + post-process the whole mini-batch on the GPUs (replaces kt_for(worker_for), map.c:1206) */
		gdh_step_t *s = (gdh_step_t *)in;
		double t0;
		int rc;
		pthread_mutex_lock(&p->order_mu);
		while (p->n_written < s->batch - 1) pthread_cond_wait(&p->order_cv, &p->order_mu); /* batch b-2 is on disk: its text buffers are free */
		pthread_mutex_unlock(&p->order_mu);
		t0 = realtime();
#ifdef GD_HOST_LR
		rc = gd_multi_lr_map_sam(p->gm, s->n, s->names, s->off, s->len, s->buf, s->qual, &p->mo, &p->po, p->n_ref, p->ref_names, p->ref_off,
		                         p->ref_len, p->ref, &s->parts, &s->part_len, &s->n_parts);
#else
		rc = gd_multi_sr_map_sam(p->gm, s->n, s->names, s->off, s->len, s->buf, s->qual, &p->mo, &p->po, p->n_ref, p->ref_names, p->ref_off,
		                         p->ref_len, p->ref, &s->parts, &s->part_len, &s->n_parts);
#endif
		if (rc != GD_OK) gdh_die(gd_multi_strerror(p->gm));
		p->t_map += realtime() - t0;
		return s;
	} else if (step == 2) { /* write (map.c:1208-1256 + the frees of :1257-1278) */
		gdh_step_t *s = (gdh_step_t *)in;
		double t0 = realtime();
		int i;
		for (i = 0; i < s->n_parts; ++i) /* (the pieces belong to the gd_multi handle: valid until the batch after the next one) */
			if (s->part_len[i]) mm_err_fwrite(s->parts[i], 1, s->part_len[i], stdout);
		gd_free(s->parts), gd_free(s->part_len);
		for (i = 0; s->seq && i < s->n; ++i) {
			free(s->seq[i].seq), free(s->seq[i].name);
			if (s->seq[i].qual) free(s->seq[i].qual);
			if (s->seq[i].comment) free(s->seq[i].comment);
		}
		free(s->seq), free(s->names), free(s->off), free(s->len), free(s->name_blob);
		if (s->qual) gdh_qbuf_put(s->qual, s->qual_cap);
		gdh_buf_put(s->buf, s->buf_cap);
		p->t_write += realtime() - t0;
		pthread_mutex_lock(&p->order_mu);
		++p->n_written;
		pthread_cond_broadcast(&p->order_cv);
		pthread_mutex_unlock(&p->order_mu);
		if (mm_verbose >= 3)
			fprintf(stderr, "[M::%s::%.3f*%.2f] mapped %d sequences\n", __func__, realtime() - mm_realtime0,
			        cputime() / (realtime() - mm_realtime0), s->n);
		free(s);
	}
	return 0;
}

/* ---- devices and the index part that is alive (main.c: read a part, map every query file against it, destroy it) ---------- */
static struct {
	gd_multi *gm;
	int n_gpus;
	const mm_idx_t *mi;   /* the part the fields below belong to (0: none) */
	char *ref;            /* its contigs as ASCII, back to back: the device index is built from them, the long-read host stage reads them */
	int64_t *ref_off;
	int32_t *ref_len;
	const char **ref_names;
	int n_ref;
} gdh_dev;

static void gdh_devices(void)
{
	const char *env = getenv("GDIET_GPUS");
	const double t0 = realtime();
	if (gdh_dev.gm) return;
	gdh_dev.n_gpus = env && atoi(env) > 0 ? atoi(env) : 1;
	setenv("CUDA_MODULE_LOADING", "EAGER", 0); /* kernels are loaded here, beside the reading of the reference, not inside the first mini-batch (measured: -0.05..0.1 s) */
	if (gd_multi_init(gdh_dev.n_gpus, 0, &gdh_dev.gm) != GD_OK) gdh_die("no usable CUDA device (there is no CPU fallback in this build)");
	if (!getenv("GDIET_MAP_LANES")) { /* a one-shot program: two lanes per device.  Four (the library's default) map 17 % more reads
	                                   * per second once warm, but creating the extra contexts costs ~0.2 s per process, and this
	                                   * pipeline is bound by its reader and writer, not by the mapping step */
		int i;
		for (i = 0; i < gdh_dev.n_gpus; ++i) gd_set_option(gd_multi_ctx(gdh_dev.gm, i), "map_lanes", 2);
	}
#ifndef GD_HOST_LR
	/* lanes and page-locked text buffers for mini-batches of the preset's -K 50M (about four bytes of SAM text per base), made here,
	 * beside the reading of the reference, instead of inside the first two mapping calls (0.3 s) */
	gd_multi_prepare_sam(gdh_dev.gm, (size_t)200 << 20);
#endif
	if (mm_verbose >= 3)
		fprintf(stderr, "[M::%s::%.3f*%.2f] %d CUDA device(s) ready in %.3f s\n", __func__, realtime() - mm_realtime0, cputime() / (realtime() - mm_realtime0),
		        gdh_dev.n_gpus, realtime() - t0);
}
/* creating the CUDA contexts takes a second or two: it runs beside the reading of the reference */
static void *gdh_devices_thread(void *arg)
{
	(void)arg;
	gdh_devices();
	return 0;
}

static void gdh_drop_part(void)
{
	free(gdh_dev.ref), free(gdh_dev.ref_off), free(gdh_dev.ref_len), free((void *)gdh_dev.ref_names);
	gdh_dev.ref = 0, gdh_dev.ref_off = 0, gdh_dev.ref_len = 0, gdh_dev.ref_names = 0, gdh_dev.n_ref = 0, gdh_dev.mi = 0;
}

/* names / offsets from mi->seq, then the device index from gdh_dev.ref on the first GPU, broadcast to the others */
static void gdh_device_index(const mm_idx_t *mi, const char *pattern, int pattern_len)
{
	gd_index *gi = 0;
	int64_t tot = 0;
	uint32_t i;
	const double t0 = realtime();
	gdh_devices();
	gdh_dev.n_ref = (int)mi->n_seq;
	gdh_dev.ref_names = (const char **)malloc(sizeof(char *) * (mi->n_seq + 1));
	gdh_dev.ref_off = (int64_t *)malloc(sizeof(int64_t) * (mi->n_seq + 1)), gdh_dev.ref_len = (int32_t *)malloc(sizeof(int32_t) * (mi->n_seq + 1));
	for (i = 0; i < mi->n_seq; ++i)
		gdh_dev.ref_names[i] = mi->seq[i].name, gdh_dev.ref_off[i] = tot, gdh_dev.ref_len[i] = (int32_t)mi->seq[i].len, tot += mi->seq[i].len;
	if (gd_index_build(gd_multi_ctx(gdh_dev.gm, 0), gdh_dev.n_ref, gdh_dev.ref_off, gdh_dev.ref_len, gdh_dev.ref, mi->w, mi->k, pattern, pattern_len, &gi) != GD_OK)
		gdh_die(gd_strerror(gd_multi_ctx(gdh_dev.gm, 0)));
	if (gd_multi_index_bcast(gdh_dev.gm, gi, 1) != GD_OK) gdh_die(gd_multi_strerror(gdh_dev.gm));
	gdh_dev.mi = mi;
	if (mm_verbose >= 3)
		fprintf(stderr, "[M::%s::%.3f*%.2f] device index on %d GPU(s) in %.3f s (broadcast %.0f MB in %.3f s, %s)\n", __func__,
		        realtime() - mm_realtime0, cputime() / (realtime() - mm_realtime0), gdh_dev.n_gpus, realtime() - t0,
		        gd_multi_stat(gdh_dev.gm, "bcast_bytes") / 1e6, gd_multi_stat(gdh_dev.gm, "bcast_seconds"),
		        gd_multi_stat(gdh_dev.gm, "bcast_path") == 1 ? "ncclBroadcast" : gdh_dev.n_gpus > 1 ? "peer copies" : "single GPU");
}

#ifndef GDH_READER_ONLY /* (oracle/reader_check.c compiles this file for its reader alone) */
/* the reference's functions of these names, compiled from index.c under other names (oracle/Makefile build_prog_batched) */
mm_idx_t *gdref_cpu_mm_idx_reader_read(mm_idx_reader_t *r, int n_threads);
int32_t gdref_cpu_mm_idx_cal_max_occ(const mm_idx_t *mi, float f);
void gdref_cpu_mm_idx_stat(const mm_idx_t *mi);
void gdref_cpu_mm_idx_destroy(mm_idx_t *mi);
mm_idx_t *mm_idx_init(int w, int k, int b, int flag); /* index.c:46 (not in a header) */

/* The reference file itself: mm_idx_reader_open does not keep the name, so the open / eof / close trio is wrapped as well and a
 * regular uncompressed file is mapped and parsed by the reader of this file (same kseq rules as for the reads; the
 * reference's reader needs 7 s for 3.1 GB).  gzip / stdin references, .mmi files, -d and -H keep the reference's reader. */
mm_idx_reader_t *gdref_cpu_mm_idx_reader_open(const char *fn, const mm_idxopt_t *opt, const char *fn_out);
int gdref_cpu_mm_idx_reader_eof(const mm_idx_reader_t *r);
void gdref_cpu_mm_idx_reader_close(mm_idx_reader_t *r);
static struct {
	const mm_idx_reader_t *r;
	const char *map;
	size_t len, pos;
	int last;
} gdh_refmap;

mm_idx_reader_t *mm_idx_reader_open(const char *fn, const mm_idxopt_t *opt, const char *fn_out)
{
	mm_idx_reader_t *r = gdref_cpu_mm_idx_reader_open(fn, opt, fn_out);
	struct stat st;
	int fd;
	if (!r || r->is_idx || fn_out || (opt->flag & MM_I_HPC) || getenv("GDIET_REF_INDEX") || getenv("GDIET_REF_READER") || !strcmp(fn, "-")) return r;
	if ((fd = open(fn, O_RDONLY)) < 0) return r;
	if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 2) {
		void *mp = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
		if (mp != MAP_FAILED) {
			const unsigned char *u = (const unsigned char *)mp;
			if (u[0] == 0x1f && u[1] == 0x8b) munmap(mp, (size_t)st.st_size); /* gzip */
			else {
				madvise(mp, (size_t)st.st_size, MADV_SEQUENTIAL);
				gdh_refmap.r = r, gdh_refmap.map = (const char *)mp, gdh_refmap.len = (size_t)st.st_size, gdh_refmap.pos = 0, gdh_refmap.last = 0;
			}
		}
	}
	close(fd);
	return r;
}

int mm_idx_reader_eof(const mm_idx_reader_t *r)
{
	if (r && r == gdh_refmap.r) return gdh_refmap.pos >= gdh_refmap.len && gdh_refmap.last == 0;
	return gdref_cpu_mm_idx_reader_eof(r);
}

void mm_idx_reader_close(mm_idx_reader_t *r)
{
	if (r && r == gdh_refmap.r) munmap((void *)gdh_refmap.map, gdh_refmap.len), memset(&gdh_refmap, 0, sizeof(gdh_refmap));
	gdref_cpu_mm_idx_reader_close(r);
}

/* one mini-batch of reference sequences: from the mapped file, or through the reference's reader */
typedef struct {
	int n;
	gdh_step_t st;       /* mapped file */
	mm_bseq1_t *seq;     /* reference's reader */
} gdh_refbatch_t;
static int gdh_ref_next(mm_idx_reader_t *r, int64_t mini, gdh_refbatch_t *b)
{
	memset(b, 0, sizeof(*b));
	if (r == gdh_refmap.r) {
		gdh_pipeline_t q;
		memset(&q, 0, sizeof(q));
		q.map = gdh_refmap.map, q.map_len = gdh_refmap.len, q.map_pos = gdh_refmap.pos, q.map_last = gdh_refmap.last;
		q.mini_batch_size = mini, q.n_threads = 1, q.plain_buffers = 1;
		b->n = gdh_read_mapped(&q, &b->st, 0);
		gdh_refmap.pos = q.map_pos, gdh_refmap.last = q.map_last;
	} else b->seq = mm_bseq_read(r->fp.seq, mini, 0, &b->n);
	return b->n;
}

/* A FASTA reference that fits one index part (the usual case: -I 4G) is parsed on several threads: the lines that start with
 * '>' cut the file into records (kseq ends a sequence at such a line, kseq.h:209), every record goes through the reader above on
 * its own byte range, and the pieces are copied to their places.  A line that starts with '@' or '+' (FASTQ, where a quality
 * line may start with '>') or a record that does not come back as exactly one sequence sends the file to the one-thread path. */
typedef struct {
	const char *map;
	size_t len;
	int n_threads;
	size_t **starts;    /* per scan range: the record starts found in it */
	int *n_starts, *bad;
	size_t *rec_beg, *rec_end;
	gdh_step_t *rec;    /* per record: what the reader made of it */
	int *rec_bad;
	char *dst;
	uint64_t *dst_off;
} gdh_refpar_t;
static void gdh_refpar_scan(void *data, long i, int tid)
{
	gdh_refpar_t *P = (gdh_refpar_t *)data;
	const size_t lo = P->len / (size_t)P->n_threads * (size_t)i, hi = i + 1 == P->n_threads ? P->len : P->len / (size_t)P->n_threads * (size_t)(i + 1);
	const char *q = P->map + lo, *e = P->map + hi;
	int n = 0, m = 0;
	size_t *a = 0;
	(void)tid;
	if (lo > 0) { /* the first line start inside the range */
		q = (const char *)memchr(P->map + lo - 1, '\n', hi - (lo - 1));
		q = q ? q + 1 : e;
	}
	while (q < e) {
		const char *nl;
		if (*q == '>') {
			if (n == m) m = m ? m * 2 : 64, a = (size_t *)realloc(a, sizeof(size_t) * m);
			a[n++] = (size_t)(q - P->map);
		} else if (*q == '@' || *q == '+') P->bad[i] = 1;
		nl = (const char *)memchr(q, '\n', (size_t)(P->map + P->len - q));
		if (!nl) break;
		q = nl + 1;
	}
	P->starts[i] = a, P->n_starts[i] = n;
}
static void gdh_refpar_parse(void *data, long i, int tid)
{
	gdh_refpar_t *P = (gdh_refpar_t *)data;
	gdh_pipeline_t q;
	(void)tid;
	memset(&q, 0, sizeof(q));
	q.map = P->map, q.map_len = P->rec_end[i], q.map_pos = P->rec_beg[i], q.mini_batch_size = INT64_MAX, q.n_threads = 1, q.plain_buffers = 1;
	if (gdh_read_mapped(&q, &P->rec[i], 0) != 1 || q.map_pos < q.map_len || q.map_last != 0) P->rec_bad[i] = 1;
}
static void gdh_refpar_copy(void *data, long i, int tid)
{
	gdh_refpar_t *P = (gdh_refpar_t *)data;
	(void)tid;
	memcpy(P->dst + P->dst_off[i], P->rec[i].buf + P->rec[i].off[0], (size_t)P->rec[i].len[0]);
}
/* fills mi->seq / gdh_dev.ref from the whole mapped file; 0 = not this way (nothing has changed) */
static int gdh_ref_parallel(mm_idx_reader_t *r, mm_idx_t *mi, int n_threads, uint64_t *sum_len_out)
{
	gdh_refpar_t P;
	int T = n_threads < 16 ? n_threads : 16, i, n_rec = 0, k = 0, ok = 1;
	uint64_t sum = 0;
	if (r != gdh_refmap.r || gdh_refmap.pos != 0 || gdh_refmap.last != 0 || T < 4 || gdh_refmap.len < (getenv("GDIET_REF_PAR_MIN_BYTES") ? (size_t)atoll(getenv("GDIET_REF_PAR_MIN_BYTES")) : (size_t)4 << 20) ||
	    (uint64_t)gdh_refmap.len > r->opt.batch_size || gdh_refmap.map[0] != '>' || getenv("GDIET_REF_ONE_THREAD")) return 0;
	memset(&P, 0, sizeof(P));
	P.map = gdh_refmap.map, P.len = gdh_refmap.len, P.n_threads = T;
	P.starts = (size_t **)calloc(T, sizeof(size_t *)), P.n_starts = (int *)calloc(T, sizeof(int)), P.bad = (int *)calloc(T, sizeof(int));
	kt_for(T, gdh_refpar_scan, &P, T);
	for (i = 0; i < T; ++i) n_rec += P.n_starts[i], ok &= !P.bad[i];
	if (ok && n_rec > 0) {
		P.rec_beg = (size_t *)malloc(sizeof(size_t) * n_rec), P.rec_end = (size_t *)malloc(sizeof(size_t) * n_rec);
		P.rec = (gdh_step_t *)calloc(n_rec, sizeof(gdh_step_t)), P.rec_bad = (int *)calloc(n_rec, sizeof(int));
		P.dst_off = (uint64_t *)malloc(sizeof(uint64_t) * n_rec);
		for (i = 0; i < T; ++i) {
			int j;
			for (j = 0; j < P.n_starts[i]; ++j) P.rec_beg[k++] = P.starts[i][j];
		}
		for (i = 0; i < n_rec; ++i) P.rec_end[i] = i + 1 < n_rec ? P.rec_beg[i + 1] : P.len;
		kt_for(T, gdh_refpar_parse, &P, n_rec);
		for (i = 0; i < n_rec; ++i) ok &= !P.rec_bad[i];
		if (ok) {
			for (i = 0; i < n_rec; ++i) P.dst_off[i] = sum, sum += (uint64_t)P.rec[i].len[0];
			if (!(gdh_dev.ref = (char *)malloc((size_t)sum + 16))) gdh_die("out of memory for the reference sequences");
			P.dst = gdh_dev.ref;
			kt_for(T, gdh_refpar_copy, &P, n_rec);
			mi->seq = (mm_idx_seq_t *)krealloc(mi->km, mi->seq, ((size_t)n_rec + 1) * sizeof(mm_idx_seq_t));
			for (i = 0; i < n_rec; ++i) {
				mm_idx_seq_t *t = &mi->seq[mi->n_seq++];
				const char *name = P.rec[i].names[0];
				if (!(mi->flag & MM_I_NO_NAME)) {
					t->name = (char *)kmalloc(mi->km, strlen(name) + 1);
					strcpy(t->name, name);
				} else t->name = 0;
				t->len = (uint32_t)P.rec[i].len[0], t->offset = P.dst_off[i], t->is_alt = 0;
				if (t->len == 0 && mm_verbose >= 2) fprintf(stderr, "[WARNING] the length database sequence '%s' is 0\n", name);
			}
			gdh_refmap.pos = gdh_refmap.len, gdh_refmap.last = 0;
			*sum_len_out = sum;
		}
		for (i = 0; i < n_rec; ++i)
			free(P.rec[i].buf), free(P.rec[i].off), free(P.rec[i].len), free((void *)P.rec[i].names), free(P.rec[i].name_blob);
	} else ok = 0;
	for (i = 0; i < T; ++i) free(P.starts[i]);
	free(P.starts), free(P.n_starts), free(P.bad), free(P.rec_beg), free(P.rec_end), free(P.rec), free(P.rec_bad), free(P.dst_off);
	return ok;
}

/* mm_idx_reader_read (index.c:624-640) for a FASTA reference: what mm_idx_gen's step 0 does (index.c:309-364: names, lengths,
 * offsets; up to -I bases per part) with the bases kept as ASCII, then the device index instead of steps 1-2 + mm_idx_post. */
mm_idx_t *mm_idx_reader_read(mm_idx_reader_t *r, int n_threads)
{
	mm_idx_t *mi;
	uint64_t sum_len = 0, cap = 0;
	pthread_t dev_thread;
	int have_thread = 0, par = 0;
	const int64_t mini = (uint64_t)r->opt.mini_batch_size < r->opt.batch_size ? r->opt.mini_batch_size : (int64_t)r->opt.batch_size; /* index.c:394 */
	if (r->is_idx || r->fp_out || (r->opt.flag & MM_I_HPC) || getenv("GDIET_REF_INDEX")) return gdref_cpu_mm_idx_reader_read(r, n_threads);
	if (r == gdh_refmap.r ? mm_idx_reader_eof(r) : (r->fp.seq == 0 || mm_bseq_eof(r->fp.seq))) return 0;
	gdh_drop_part();
	if (!gdh_dev.gm) have_thread = pthread_create(&dev_thread, 0, gdh_devices_thread, 0) == 0;
	mi = mm_idx_init(r->opt.w, r->opt.k, r->opt.bucket_bits, r->opt.flag);
	if (gdh_ref_parallel(r, mi, n_threads, &sum_len)) cap = sum_len + 16, par = 1;
	else while (sum_len <= r->opt.batch_size) { /* (index.c:311-314: a part is closed by the first mini-batch that takes it past -I) */
		gdh_refbatch_t b;
		int n, i;
		uint32_t old_m, m;
		uint64_t add = 0;
		if ((n = gdh_ref_next(r, mini, &b)) <= 0) break;
		old_m = mi->n_seq, m = mi->n_seq + n;
		kroundup32(m);
		kroundup32(old_m);
		if (old_m != m) mi->seq = (mm_idx_seq_t *)krealloc(mi->km, mi->seq, m * sizeof(mm_idx_seq_t));
		for (i = 0; i < n; ++i) add += b.seq ? b.seq[i].l_seq : b.st.len[i];
		if (sum_len + add + 16 > cap) {
			cap = (sum_len + add) + (sum_len + add) / 2 + 4096;
			if (!(gdh_dev.ref = (char *)realloc(gdh_dev.ref, cap))) gdh_die("out of memory for the reference sequences");
		}
		for (i = 0; i < n; ++i) {
			mm_idx_seq_t *t = &mi->seq[mi->n_seq++];
			const char *name = b.seq ? b.seq[i].name : b.st.names[i];
			const uint32_t l = b.seq ? (uint32_t)b.seq[i].l_seq : (uint32_t)b.st.len[i];
			if (!(mi->flag & MM_I_NO_NAME)) {
				t->name = (char *)kmalloc(mi->km, strlen(name) + 1);
				strcpy(t->name, name);
			} else t->name = 0;
			t->len = l, t->offset = sum_len, t->is_alt = 0;
			memcpy(gdh_dev.ref + sum_len, b.seq ? b.seq[i].seq : b.st.buf + b.st.off[i], l);
			sum_len += l;
			if (l == 0 && mm_verbose >= 2) fprintf(stderr, "[WARNING] the length database sequence '%s' is 0\n", name);
			if (b.seq) free(b.seq[i].seq), free(b.seq[i].name);
		}
		if (b.seq) free(b.seq);
		else free(b.st.buf), free(b.st.off), free(b.st.len), free((void *)b.st.names), free(b.st.name_blob);
	}
	if (mm_verbose >= 3)
		fprintf(stderr, "[M::%s::%.3f*%.2f] read %u reference sequences, %ld bases%s\n", __func__, realtime() - mm_realtime0,
		        cputime() / (realtime() - mm_realtime0), mi->n_seq, (long)sum_len, par ? " (records parsed on several threads)" : "");
	/* (mi->S stays empty: nothing on this path reads it; the flag word is left alone because main.c tests MM_I_NO_SEQ) */
	if (have_thread) pthread_join(dev_thread, 0);
	gdh_device_index(mi, r->opt.pattern, r->opt.pattern_len);
	mi->index = r->n_parts++;
	return mi;
}

int32_t mm_idx_cal_max_occ(const mm_idx_t *mi, float f)
{
	int32_t v = 0;
	if (mi != gdh_dev.mi) return gdref_cpu_mm_idx_cal_max_occ(mi, f);
	if (f <= 0.) return INT32_MAX;
	if (gd_index_cal_max_occ(gd_multi_ctx(gdh_dev.gm, 0), gd_multi_index(gdh_dev.gm, 0), f, &v) != GD_OK) gdh_die(gd_strerror(gd_multi_ctx(gdh_dev.gm, 0)));
	return v;
}

void mm_idx_stat(const mm_idx_t *mi)
{
	const gd_index *gi;
	int64_t keys, recs, len;
	if (mi != gdh_dev.mi) {
		gdref_cpu_mm_idx_stat(mi);
		return;
	}
	gi = gd_multi_index(gdh_dev.gm, 0);
	keys = gd_index_stat(gi, "n_keys"), recs = gd_index_stat(gi, "n_minimizers"), len = gd_index_stat(gi, "total_len");
	fprintf(stderr, "[M::%s] kmer size: %d; skip: %d; is_hpc: %d; #seq: %d\n", __func__, mi->k, mi->w, mi->flag & MM_I_HPC, mi->n_seq);
	fprintf(stderr, "[M::%s::%.3f*%.2f] (device index) distinct minimizers: %ld; average occurrences: %.3lf; average spacing: %.3lf; total length: %ld\n",
	        __func__, realtime() - mm_realtime0, cputime() / (realtime() - mm_realtime0), (long)keys, keys ? (double)recs / keys : 0.0,
	        recs ? (double)len / recs : 0.0, (long)len);
}

void mm_idx_destroy(mm_idx_t *mi)
{
	if (mi && mi == gdh_dev.mi) gdh_drop_part(); /* (the device copies go when the next part is broadcast, or with the process) */
	gdref_cpu_mm_idx_destroy(mi);
}
#endif

static void gdh_options(gdh_pipeline_t *p)
{
	const mm_mapopt_t *opt = p->opt;
	memset(&p->mo, 0, sizeof(p->mo)), memset(&p->po, 0, sizeof(p->po));
	if (opt->pattern_len < 2 || opt->pattern_len > 63) gdh_die("pattern length (-W) must be 2..63 on the batched device path");
	p->mo.W = opt->pattern_len, memcpy(p->mo.Z, opt->pattern, opt->pattern_len);
	p->mo.max_seeds = opt->max_seeds, p->mo.frag_mode = !!(opt->flag & MM_F_FRAG_MODE), p->mo.max_frag_len = opt->max_frag_len;
	p->mo.mid_occ = opt->mid_occ, p->mo.max_max_occ = opt->max_max_occ, p->mo.occ_dist = opt->occ_dist, p->mo.q_occ_frac = opt->q_occ_frac;
	p->mo.for_only = !!(opt->flag & MM_F_FOR_ONLY), p->mo.rev_only = !!(opt->flag & MM_F_REV_ONLY);
	p->mo.a = opt->a, p->mo.b = opt->b, p->mo.q = opt->q, p->mo.e = opt->e, p->mo.q2 = opt->q2, p->mo.e2 = opt->e2;
	p->mo.zdrop = opt->zdrop, p->mo.end_bonus = opt->end_bonus;
#ifdef GD_HOST_LR
	p->mo.bw = opt->bw; /* LR/map.c:1374 */
	p->mo.vt_dis = opt->vt_dis, p->mo.vt_nb_loc = opt->vt_nb_loc, p->mo.vt_cov = opt->vt_cov, p->mo.vt_df1 = opt->vt_df1;
	p->mo.vt_df2 = opt->vt_df2, p->mo.vt_f = opt->vt_f, p->mo.max_max_gap = opt->max_max_gap, p->mo.max_min_gap = opt->max_min_gap;
	if (opt->flag & (MM_F_RADIX_SORT | MM_F_HEAP_SORT)) gdh_die("--sort=radix|heap is not covered by the long-read device path (the presets use merge)");
#else
	p->mo.bw_frac = opt->bw_frac, p->mo.bw_min = (uint32_t)opt->bw_min, p->mo.bw_max = (uint32_t)opt->bw_max; /* per read, map.c:624-631 */
	if (opt->bw_max <= 0) gdh_die("-r: bw_max must be positive");
	p->mo.min_cnt = opt->min_cnt, p->mo.rec_threshold_frac = opt->rec_threshold_frac, p->mo.af_max_loc = opt->AF_max_loc;
#endif
	p->po.a = opt->a, p->po.b = opt->b, p->po.q = opt->q, p->po.e = opt->e, p->po.q2 = opt->q2, p->po.e2 = opt->e2;
	p->po.min_dp_max = opt->min_dp_max, p->po.best_n = opt->best_n, p->po.no_print_2nd = !!(opt->flag & MM_F_NO_PRINT_2ND);
	p->po.is_sr = !!(opt->flag & MM_F_SR), p->po.sam_hit_only = !!(opt->flag & MM_F_SAM_HIT_ONLY), p->po.softclip = !!(opt->flag & MM_F_SOFTCLIP);
	p->po.n_threads = p->n_threads;
}

static void gdh_refuse_uncovered(const mm_mapopt_t *opt, int n_segs)
{
	if (n_segs != 1) gdh_die("paired / multi-file input is not covered by the batched device path (use the GDiet_cuda build)");
	if (!(opt->flag & MM_F_OUT_SAM) || !(opt->flag & MM_F_CIGAR)) gdh_die("the batched device path writes SAM with CIGAR (-a); PAF output is not covered");
	if (opt->split_prefix) gdh_die("--split-prefix is not covered by the batched device path");
	if (opt->flag & (MM_F_OUT_CS | MM_F_OUT_MD | MM_F_EQX | MM_F_COPY_COMMENT | MM_F_SPLICE | MM_F_OUT_CS_LONG | MM_F_NO_DIAG | MM_F_NO_DUAL))
		gdh_die("--cs / --MD / --eqx / -y / splice mode / -X are not covered by the batched device path");
	if (opt->sdust_thres > 0) gdh_die("-T (sdust) is not covered by the batched device path");
}

int mm_map_file_frag(const mm_idx_t *idx, int n_segs, const char **fn, const mm_mapopt_t *opt, int n_threads)
{
	gdh_pipeline_t pl;
	double t0;
	if (n_segs < 1) return -1;
	gdh_refuse_uncovered(opt, n_segs);
	memset(&pl, 0, sizeof(pl));
	pthread_mutex_init(&pl.order_mu, 0), pthread_cond_init(&pl.order_cv, 0);
	{ /* a regular, uncompressed file is mapped and parsed in place; gzip / stdin input goes through the reference's reader */
		struct stat st;
		int fd = strcmp(fn[0], "-") ? open(fn[0], O_RDONLY) : -1;
		if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 2 && !getenv("GDIET_REF_READER")) {
			void *mp = mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
			if (mp != MAP_FAILED) {
				const unsigned char *u = (const unsigned char *)mp;
				if (u[0] == 0x1f && u[1] == 0x8b) munmap(mp, (size_t)st.st_size); /* gzip */
				else {
					madvise(mp, (size_t)st.st_size, MADV_SEQUENTIAL);
					pl.map = (const char *)mp, pl.map_len = (size_t)st.st_size;
				}
			}
		}
		if (fd >= 0) close(fd);
	}
	if (!pl.map && (pl.fp = mm_bseq_open(fn[0])) == 0) {
		if (mm_verbose >= 1) fprintf(stderr, "ERROR: failed to open file '%s': %s\n", fn[0], strerror(errno));
		return -1;
	}
	pl.opt = opt, pl.mi = idx, pl.n_threads = n_threads > 1 ? n_threads : 1, pl.mini_batch_size = opt->mini_batch_size;
	/* two passes over the bytes: worth it from four threads on (measured: 2.8 M short reads/s on one thread, 2.4 M on two, 5.6 M on
	 * four, 11.7 M on eight) */
	pl.par_min_bytes = getenv("GDIET_READER_THREADS_MIN_BYTES") ? (size_t)atoll(getenv("GDIET_READER_THREADS_MIN_BYTES"))
	                   : n_threads >= 4 ? (size_t)1 << 20 : (size_t)-1;
	gdh_options(&pl);
	gdh_devices();
#ifdef GD_HOST_LR
	/* Optional (GDIET_LR_BATCH_BASES / GDIET_LR_BATCH_READS): close a mini-batch before -K once it holds that many bases AND
	 * reads.  Records are independent and written in input order, so the text does not depend on where batches end.  Off by
	 * default: measured on configs 3 and 4 it buys nothing (reading and writing are ~10 % of the mapping time) and an ONT batch
	 * cut in two leaves the second DP launch with a fifth of the pairs. */
	{
		const char *e = getenv("GDIET_LR_BATCH_BASES"), *e2 = getenv("GDIET_LR_BATCH_READS");
		pl.sub_bases = e ? atoll(e) : 0;
		pl.sub_reads = e2 ? atoi(e2) : 0;
	}
#endif
	if (gdh_dev.mi != idx) { /* the index came from an .mmi (or the reference's mm_idx_gen): the contigs it holds, as ASCII */
		uint32_t i;
		uint64_t tot = 0;
		if (idx->flag & MM_I_NO_SEQ) gdh_die("the index holds no sequences (built with --idx-no-seq)");
		gdh_drop_part();
		for (i = 0; i < idx->n_seq; ++i) tot += idx->seq[i].len;
		if (!(gdh_dev.ref = (char *)malloc((size_t)tot + 16))) gdh_die("out of memory for the reference sequences");
		for (i = 0, tot = 0; i < idx->n_seq; tot += idx->seq[i++].len) { /* S is 4-bit packed, 8 bases per word (index.c:157-166) */
			const uint64_t st = idx->seq[i].offset;
			char *dst = gdh_dev.ref + tot;
			uint32_t j;
			for (j = 0; j < idx->seq[i].len; ++j) dst[j] = "ACGTN"[(idx->S[(st + j) >> 3] >> (((st + j) & 7) << 2)) & 0xf];
		}
		gdh_device_index(idx, opt->pattern, opt->pattern_len);
	}
	pl.gm = gdh_dev.gm, pl.n_ref = gdh_dev.n_ref, pl.ref_names = gdh_dev.ref_names, pl.ref_off = gdh_dev.ref_off, pl.ref_len = gdh_dev.ref_len;
	pl.ref = gdh_dev.ref;

	t0 = realtime();
	kt_pipeline(n_threads == 1 ? 1 : 3, gdh_worker, &pl, 3); /* reader, mapper and writer overlap (map.c:1316-1317) */
	if (mm_verbose >= 3)
		fprintf(stderr, "[M::%s] %ld reads, %ld bases in %.3f s: %.0f reads/s (step seconds: read %.3f, map %.3f, write %.3f)\n", __func__,
		        (long)pl.n_reads, (long)pl.n_bases, realtime() - t0, pl.n_reads / (realtime() - t0 + 1e-9), pl.t_read, pl.t_map, pl.t_write);

	if (pl.fp) mm_bseq_close(pl.fp);
	if (pl.map) munmap((void *)pl.map, pl.map_len);
	return 0;
}

/* main.c calls this one-file form whenever the preset is not in fragment mode (every long-read preset; main.c:652-656 of the
 * long-read tree, map.c:2254): it has to land in the device pipeline too, not in the renamed CPU one. */
int mm_map_file(const mm_idx_t *idx, const char *fn, const mm_mapopt_t *opt, int n_threads)
{
	return mm_map_file_frag(idx, 1, &fn, opt, n_threads);
}
