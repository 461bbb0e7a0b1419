/* oracle/reader_check.c -- TEST INFRASTRUCTURE ONLY.
 * The mini-batch reader of genome-on-diet_b200/host/gd_batched_host.c (mapped file, memchr) against the reference's
 * own reader (mm_bseq_read3 over kseq.h, GDiet-ShortReads/bseq.c:73-113) on the same file and -K: every batch must hold
 * the same reads with the same names, bases and qualities.  No GPU needed (nothing of the library is called).
 *     reader_check <file> <mini_batch_bases> [threads]     prints "OK <batches> <reads> <batches read by the several-thread
 *                                                          reader>" or the first difference */
#include <stdlib.h>
#define gd_pinned_alloc(n) malloc(n) /* no GPU here: the batch buffers are plain memory in this check */
#define gd_pinned_free(p) free(p)
#define GDH_READER_ONLY
#include "../genome-on-diet_b200/host/gd_batched_host.c"

int main(int argc, char **argv)
{
	gdh_pipeline_t pl;
	mm_mapopt_t opt;
	mm_bseq_file_t *fp;
	struct stat st;
	int fd, batches = 0;
	long reads = 0;
	if (argc < 3) return 2;
	memset(&pl, 0, sizeof(pl)), memset(&opt, 0, sizeof(opt));
	pl.opt = &opt, pl.mini_batch_size = atol(argv[2]);
	pl.n_threads = argc > 3 ? atoi(argv[3]) : 1;
	fd = open(argv[1], O_RDONLY);
	if (fd < 0 || fstat(fd, &st) != 0) return 2;
	pl.map = st.st_size ? (const char *)mmap(0, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0) : "";
	pl.map_len = (size_t)st.st_size;
	if (argc > 4 && !strcmp(argv[4], "time")) { /* the product reader alone: reads per second (no comparison) */
		double t0 = realtime();
		for (;;) {
			gdh_step_t s;
			memset(&s, 0, sizeof(s));
			if (gdh_read_mapped(&pl, &s, 1) == 0) break;
			reads += s.n, ++batches;
			free(s.buf), free(s.qual), free(s.off), free(s.len), free(s.names), free(s.name_blob);
		}
		printf("TIME %d batches %ld reads %.3f s %.0f reads/s (%ld batches on several threads)\n", batches, reads, realtime() - t0,
		       reads / (realtime() - t0 + 1e-9), pl.n_par_batches);
		return 0;
	}
	fp = mm_bseq_open(argv[1]);
	for (;;) {
		gdh_step_t s;
		int n_ref = 0, n, i;
		mm_bseq1_t *r = mm_bseq_read3(fp, pl.mini_batch_size, 1, 0, 0, &n_ref);
		memset(&s, 0, sizeof(s));
		n = gdh_read_mapped(&pl, &s, 1);
		if (n != n_ref) {
			printf("DIFF batch %d: %d reads, reference %d\n", batches, n, n_ref);
			return 1;
		}
		if (n == 0) break;
		for (i = 0; i < n; ++i) {
			const int has_q = s.qual != 0;
			if (strcmp(s.names[i], r[i].name) || s.len[i] != r[i].l_seq || memcmp(s.buf + s.off[i], r[i].seq, r[i].l_seq) ||
			    (r[i].l_seq > 0 && has_q != (r[i].qual != 0)) || (has_q && r[i].qual && memcmp(s.qual + s.off[i], r[i].qual, r[i].l_seq))) {
				printf("DIFF batch %d read %d: '%s' len %d vs reference '%s' len %d (qual %d vs %d)\n", batches, i, s.names[i], s.len[i], r[i].name,
				       r[i].l_seq, has_q, r[i].qual != 0);
				return 1;
			}
			free(r[i].name), free(r[i].seq), free(r[i].qual), free(r[i].comment);
		}
		free(r);
		reads += n, ++batches;
	}
	printf("OK %d %ld %ld\n", batches, reads, pl.n_par_batches);
	return 0;
}
