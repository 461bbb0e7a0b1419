// gd_multi.cu -- several B200s of one box driven from ONE host process (the C host of INTEGRATION.md level 2).
//
// What the reference does with kt_for over host cores (GDiet-ShortReads/map.c:1045-1092,1206; order restored by
// kt_pipeline, kthread.c:101-115), this does over GPUs:
//   * one gd_ctx per device (gd_multi_init);
//   * the index is built (or loaded) ONCE on the first device and broadcast into replicas on the others
//     (gd_multi_index_bcast): NCCL's ncclBroadcast over NVLink when libnccl.so.2 can be loaded at run time (it is
//     dlopen'ed, so the library has no link-time dependency and shares the copy a host such as PyTorch already
//     loaded), else direct peer copies (cudaMemcpyPeerAsync, also NVLink);
//   * a mini-batch of reads is cut into contiguous shards, one per device, every shard is mapped by its own host
//     thread on its own device, and the results are handed back in INPUT ORDER -- either merged into the caller's
//     arrays (gd_multi_sr_map_batch / gd_multi_lr_map_batch) or, with the host stage included, as the SAM text
//     pieces of the shards in order (gd_multi_sr_map_sam / gd_multi_lr_map_sam).
// There is no data-path collective besides the one broadcast: reads are independent.
#include "gd_ctx.h"
#include <dlfcn.h>
#include <unistd.h>
#include <algorithm>
#include <chrono>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>
#include <string.h>

// long-read shards are cut into device slices of at least this many reads and bases (gd_multi_lr_map_sam; the
// environment variables GDIET_LR_SLICE_READS / GDIET_LR_SLICE_BASES override them, for tests)
#ifndef GD_LR_SLICE_READS
#define GD_LR_SLICE_READS 4096
#endif
#ifndef GD_LR_SLICE_BASES
#define GD_LR_SLICE_BASES (128ll << 20) // two 64 Mbase slices of gd_lr_map_batch: both of its lanes stay busy
#endif

// the few NCCL declarations used (nccl.h, NCCL 2.x ABI) -- resolved with dlsym
typedef struct ncclComm *gd_ncclComm_t;
typedef int (*nccl_comm_init_all_t)(gd_ncclComm_t *, int, const int *);
typedef int (*nccl_group_t)(void);
typedef int (*nccl_bcast_t)(const void *, void *, size_t, int /*ncclDataType_t*/, int, gd_ncclComm_t, cudaStream_t);
typedef int (*nccl_comm_destroy_t)(gd_ncclComm_t);
typedef const char *(*nccl_errstr_t)(int);

struct gd_multi {
	std::vector<gd_ctx *> ctx;
	std::vector<gd_index *> idx; // idx[0] is the caller's (not owned unless own_root)
	bool own_root = false;
	std::string err;
	// NCCL (optional)
	void *nccl_lib = nullptr;
	std::vector<gd_ncclComm_t> comms;
	nccl_comm_init_all_t p_init_all = nullptr;
	nccl_group_t p_group_start = nullptr, p_group_end = nullptr;
	nccl_bcast_t p_bcast = nullptr;
	nccl_comm_destroy_t p_destroy = nullptr;
	nccl_errstr_t p_errstr = nullptr;
	// stats of the last broadcast
	double bcast_s = 0;
	int64_t bcast_bytes = 0;
	int bcast_path = 0; // 1 = ncclBroadcast, 2 = peer copies
	// per-shard host staging of the merged calls
	struct Shard {
		std::vector<int64_t> cand_off;
		std::vector<gd_sr_cand_t> cand;
		std::vector<uint32_t> cigar;
		int64_t n_cand = 0, n_cig = 0;
		int rc = GD_OK;
		bool host_err = false; // rc came from the host SAM stage, not from the device
	};
	std::vector<Shard> sh;
	// text pieces of the *_map_sam calls belong to the handle: host-made (malloc'ed) pieces are freed two calls later, like the
	// device-made ones (pinned buffers of the contexts) are reused then
	std::vector<char *> owned[2];
	long sam_calls = 0;
};

extern "C" int gd_init(int device, gd_ctx **ctx);
extern "C" void gd_destroy(gd_ctx *ctx);

static bool multi_load_nccl(gd_multi *m)
{
	if (getenv("GDIET_NO_NCCL")) return false;
	const char *names[] = {"libnccl.so.2", "libnccl.so"};
	for (const char *nm : names)
		if ((m->nccl_lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL))) break;
	if (!m->nccl_lib) return false;
	m->p_init_all = (nccl_comm_init_all_t)dlsym(m->nccl_lib, "ncclCommInitAll");
	m->p_group_start = (nccl_group_t)dlsym(m->nccl_lib, "ncclGroupStart");
	m->p_group_end = (nccl_group_t)dlsym(m->nccl_lib, "ncclGroupEnd");
	m->p_bcast = (nccl_bcast_t)dlsym(m->nccl_lib, "ncclBroadcast");
	m->p_destroy = (nccl_comm_destroy_t)dlsym(m->nccl_lib, "ncclCommDestroy");
	m->p_errstr = (nccl_errstr_t)dlsym(m->nccl_lib, "ncclGetErrorString");
	return m->p_init_all && m->p_group_start && m->p_group_end && m->p_bcast && m->p_destroy;
}

extern "C" int gd_multi_init(int n_dev, const int *devices, gd_multi **out)
{
	if (!out || n_dev < 1) return GD_ERR_ARG;
	*out = nullptr;
	gd_multi *m = new gd_multi();
	for (int i = 0; i < n_dev; ++i) {
		gd_ctx *c = nullptr;
		const int rc = gd_init(devices ? devices[i] : i, &c);
		if (rc != GD_OK) {
			for (gd_ctx *x : m->ctx) gd_destroy(x);
			delete m;
			return rc;
		}
		m->ctx.push_back(c);
	}
	m->idx.assign(n_dev, nullptr);
	m->sh.resize(n_dev);
	*out = m;
	return GD_OK;
}

extern "C" void gd_multi_destroy(gd_multi *m)
{
	if (!m) return;
	for (size_t i = 0; i < m->idx.size(); ++i)
		if (m->idx[i] && (i > 0 || m->own_root)) gd_index_destroy(m->idx[i]);
	for (int g = 0; g < 2; ++g)
		for (char *p : m->owned[g]) gd_free(p);
	for (gd_ncclComm_t c : m->comms)
		if (c && m->p_destroy) m->p_destroy(c);
	for (gd_ctx *c : m->ctx) gd_destroy(c);
	// the NCCL library stays loaded: other users of the process (e.g. PyTorch) may share it
	delete m;
}

extern "C" int gd_multi_size(const gd_multi *m) { return m ? (int)m->ctx.size() : 0; }
extern "C" gd_ctx *gd_multi_ctx(gd_multi *m, int i) { return (m && i >= 0 && (size_t)i < m->ctx.size()) ? m->ctx[i] : nullptr; }
extern "C" const gd_index *gd_multi_index(const gd_multi *m, int i) { return (m && i >= 0 && (size_t)i < m->idx.size()) ? m->idx[i] : nullptr; }
extern "C" const char *gd_multi_strerror(const gd_multi *m) { return m ? m->err.c_str() : "gd_multi: null handle"; }

extern "C" double gd_multi_stat(const gd_multi *m, const char *key)
{
	if (!m || !key) return -1;
	if (!strcmp(key, "bcast_seconds")) return m->bcast_s;
	if (!strcmp(key, "bcast_bytes")) return (double)m->bcast_bytes;
	if (!strcmp(key, "bcast_path")) return m->bcast_path;
	return -1;
}

// The index of the first device (built with gd_index_build* or loaded with gd_index_load_mmi on gd_multi_ctx(m, 0))
// is replicated on every other device.  take_ownership: gd_multi_destroy frees root too.
extern "C" int gd_multi_index_bcast(gd_multi *m, gd_index *root, int take_ownership)
{
	if (!m || !root) return GD_ERR_ARG;
	const int n = (int)m->ctx.size();
	for (int i = 1; i < n; ++i)
		if (m->idx[i]) gd_index_destroy(m->idx[i]), m->idx[i] = nullptr;
	if (m->idx[0] && m->own_root && m->idx[0] != root) gd_index_destroy(m->idx[0]);
	m->idx[0] = root, m->own_root = take_ownership != 0;
	m->bcast_s = 0, m->bcast_bytes = 0, m->bcast_path = 0;
	if (n == 1) return GD_OK;
	gd_index_meta_t meta;
	int rc;
	if ((rc = gd_index_meta(root, &meta))) return rc;
	void *src[GD_INDEX_NBUF];
	size_t bytes[GD_INDEX_NBUF];
	gd_index_buffers(root, src, bytes);
	std::vector<std::vector<void *>> dst(n, std::vector<void *>(GD_INDEX_NBUF, nullptr));
	for (int i = 1; i < n; ++i) {
		if ((rc = gd_index_alloc(m->ctx[i], &meta, &m->idx[i]))) {
			m->err = std::string("gd_multi_index_bcast: ") + gd_strerror(m->ctx[i]);
			return rc;
		}
		size_t b2[GD_INDEX_NBUF];
		gd_index_buffers(m->idx[i], dst[i].data(), b2);
	}
	// the root's buffers are complete once its stream has drained
	cudaSetDevice(m->ctx[0]->device);
	cudaStreamSynchronize(m->ctx[0]->stream);
	// NCCL writes its version / debug lines to stdout, which is where the host program writes its SAM records (-o reopens
	// stdout, main.c): while NCCL initialises and broadcasts, file descriptor 1 points at stderr
	fflush(stdout);
	const int saved_stdout = dup(1);
	if (saved_stdout >= 0) dup2(2, 1);
	if (m->comms.empty() && !m->nccl_lib && multi_load_nccl(m)) {
		std::vector<int> devs(n);
		for (int i = 0; i < n; ++i) devs[i] = m->ctx[i]->device;
		m->comms.assign(n, nullptr);
		const int e = m->p_init_all(m->comms.data(), n, devs.data());
		if (e != 0) {
			fprintf(stderr, "[gdiet_cuda] ncclCommInitAll failed (%s): broadcasting the index with peer copies\n",
			        m->p_errstr ? m->p_errstr(e) : "?");
			m->comms.clear();
		}
	}
	cudaEvent_t e0, e1;
	cudaSetDevice(m->ctx[0]->device);
	cudaEventCreate(&e0), cudaEventCreate(&e1);
	cudaEventRecord(e0, m->ctx[0]->stream);
	bool ok = true;
	if (!m->comms.empty()) { // one grouped ncclBroadcast per buffer: every rank of the group is a device of this process
		m->bcast_path = 1;
		for (int b = 0; b < GD_INDEX_NBUF && ok; ++b) {
			if (!bytes[b]) continue;
			m->p_group_start();
			for (int i = 0; i < n; ++i) {
				void *buf = i == 0 ? src[b] : dst[i][b];
				const int e = m->p_bcast(buf, buf, bytes[b], /*ncclUint8*/ 1, 0, m->comms[i], m->ctx[i]->stream);
				if (e != 0) ok = false, m->err = std::string("ncclBroadcast: ") + (m->p_errstr ? m->p_errstr(e) : "?");
			}
			const int e = m->p_group_end();
			if (e != 0) ok = false, m->err = std::string("ncclGroupEnd: ") + (m->p_errstr ? m->p_errstr(e) : "?");
			m->bcast_bytes += (int64_t)bytes[b];
		}
	} else { // peer copies, each on the receiver's stream
		m->bcast_path = 2;
		for (int i = 1; i < n; ++i) {
			cudaSetDevice(m->ctx[i]->device);
			for (int b = 0; b < GD_INDEX_NBUF; ++b)
				if (bytes[b] && cudaMemcpyPeerAsync(dst[i][b], m->ctx[i]->device, src[b], m->ctx[0]->device, bytes[b], m->ctx[i]->stream) != cudaSuccess)
					ok = false, m->err = "gd_multi_index_bcast: cudaMemcpyPeerAsync failed";
		}
		for (int b = 0; b < GD_INDEX_NBUF; ++b) m->bcast_bytes += (int64_t)bytes[b];
	}
	for (int i = 0; i < n; ++i) {
		cudaSetDevice(m->ctx[i]->device);
		if (cudaStreamSynchronize(m->ctx[i]->stream) != cudaSuccess) ok = false, m->err = "gd_multi_index_bcast: stream failed";
	}
	fflush(stdout);
	if (saved_stdout >= 0) dup2(saved_stdout, 1), close(saved_stdout);
	cudaSetDevice(m->ctx[0]->device);
	cudaEventRecord(e1, m->ctx[0]->stream);
	cudaEventSynchronize(e1);
	float ms = 0;
	cudaEventElapsedTime(&ms, e0, e1);
	m->bcast_s = ms * 1e-3;
	cudaEventDestroy(e0), cudaEventDestroy(e1);
	if (!ok) return GD_ERR_CUDA;
	for (int i = 1; i < n; ++i)
		if ((rc = gd_index_commit(m->ctx[i], m->idx[i]))) {
			m->err = std::string("gd_multi_index_bcast: ") + gd_strerror(m->ctx[i]);
			return rc;
		}
	return GD_OK;
}

// contiguous shards with (nearly) equal numbers of bases: shard j = reads [cut[j], cut[j+1])
static std::vector<int> shard_cuts(int n, const int32_t *len, int parts)
{
	std::vector<int> cut(parts + 1, n);
	int64_t total = 0;
	for (int i = 0; i < n; ++i) total += len[i];
	cut[0] = 0;
	int64_t acc = 0;
	int j = 1;
	for (int i = 0; i < n && j < parts; ++i) {
		acc += len[i];
		while (j < parts && acc * parts >= total * j) cut[j++] = i + 1;
	}
	for (; j < parts; ++j) cut[j] = n;
	return cut;
}

template <class OPT, class FN>
static int multi_map(gd_multi *m, int n, const int64_t *off, const int32_t *len, const char *buf, const OPT *opt, FN map_fn,
                     int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap, int64_t *n_cigar)
{
	if (!m || n < 0 || !cand_off || !opt || (n > 0 && (!off || !len || !buf))) return GD_ERR_ARG;
	const int G = (int)m->ctx.size();
	for (int i = 0; i < G; ++i)
		if (!m->idx[i]) {
			m->err = "gd_multi: no index (gd_multi_index_bcast first)";
			return GD_ERR_ARG;
		}
	const std::vector<int> cut = shard_cuts(n, len, G);
	auto work = [&](int j) {
		gd_multi::Shard &S = m->sh[j];
		const int b = cut[j], cnt = cut[j + 1] - cut[j];
		S.n_cand = S.n_cig = 0, S.rc = GD_OK;
		S.cand_off.assign((size_t)cnt + 1, 0);
		if (cnt == 0) return;
		cudaSetDevice(m->ctx[j]->device);
		size_t want_cig = (size_t)cnt * 16;
		if (std::is_same<OPT, gd_lr_opt_t>::value) { // long reads: CIGARs grow with the bases; an undersized pool maps the shard twice
			int64_t bases = 0;
			for (int i = 0; i < cnt; ++i) bases += len[b + i];
			want_cig = std::max(want_cig, (size_t)std::min<int64_t>(std::max<int64_t>(cigar_cap, 0), bases / 3 + (int64_t)cnt * 64));
		}
		if (S.cand.size() < (size_t)cnt * 2) S.cand.resize((size_t)cnt * 2);
		if (S.cigar.size() < want_cig) S.cigar.resize(want_cig);
		for (int attempt = 0; attempt < 2; ++attempt) {
			int64_t ncig = 0;
			S.rc = map_fn(m->ctx[j], m->idx[j], cnt, off + b, len + b, buf, opt, S.cand_off.data(), S.cand.data(), (int64_t)S.cand.size(),
			              S.cigar.data(), (int64_t)S.cigar.size(), &ncig);
			S.n_cand = S.cand_off[cnt], S.n_cig = ncig;
			if (S.rc != GD_ERR_CAPACITY) break;
			S.cand.resize((size_t)S.n_cand + 16), S.cigar.resize((size_t)S.n_cig + 16);
		}
	};
	std::vector<std::thread> th;
	for (int j = 1; j < G; ++j) th.emplace_back(work, j);
	work(0);
	for (std::thread &t : th) t.join();
	for (int j = 0; j < G; ++j)
		if (m->sh[j].rc) {
			m->err = std::string("gd_multi: device ") + std::to_string(m->ctx[j]->device) + ": " + gd_strerror(m->ctx[j]);
			return m->sh[j].rc;
		}
	// input-order hand-back: shard j's records go behind those of shards 0..j-1 (offsets shifted), copied by G threads
	std::vector<int64_t> cb(G + 1, 0), gb(G + 1, 0);
	for (int j = 0; j < G; ++j) cb[j + 1] = cb[j] + m->sh[j].n_cand, gb[j + 1] = gb[j] + m->sh[j].n_cig;
	if (n_cigar) *n_cigar = gb[G];
	cand_off[n] = cb[G];
	if (cb[G] > cand_cap || gb[G] > cigar_cap || (cb[G] && !cand) || (gb[G] && !cigar)) {
		m->err = "gd_multi: output buffer too small";
		return GD_ERR_CAPACITY;
	}
	if (gb[G] > 0x7fffffff) {
		m->err = "gd_multi: CIGAR pool of one call exceeds 2^31 entries; map fewer reads per call";
		return GD_ERR_ARG;
	}
	auto merge = [&](int j) {
		const gd_multi::Shard &S = m->sh[j];
		const int b = cut[j], cnt = cut[j + 1] - cut[j];
		for (int i = 0; i < cnt; ++i) cand_off[b + i] = S.cand_off[i] + cb[j];
		for (int64_t c = 0; c < S.n_cand; ++c) {
			gd_sr_cand_t x = S.cand[(size_t)c];
			x.cigar_off += (int32_t)gb[j];
			cand[cb[j] + c] = x;
		}
		if (S.n_cig) memcpy(cigar + gb[j], S.cigar.data(), (size_t)S.n_cig * 4);
	};
	th.clear();
	for (int j = 1; j < G; ++j) th.emplace_back(merge, j);
	merge(0);
	for (std::thread &t : th) t.join();
	return GD_OK;
}

extern "C" int gd_multi_sr_map_batch(gd_multi *m, int n, const int64_t *off, const int32_t *len, const char *buf, const gd_sr_opt_t *opt,
                                     int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap,
                                     int64_t *n_cigar)
{
	return multi_map(m, n, off, len, buf, opt, gd_sr_map_batch, cand_off, cand, cand_cap, cigar, cigar_cap, n_cigar);
}

extern "C" int gd_multi_lr_map_batch(gd_multi *m, int n, const int64_t *off, const int32_t *len, const char *buf, const gd_lr_opt_t *opt,
                                     int64_t *cand_off, gd_sr_cand_t *cand, int64_t cand_cap, uint32_t *cigar, int64_t cigar_cap,
                                     int64_t *n_cigar)
{
	return multi_map(m, n, off, len, buf, opt, gd_lr_map_batch, cand_off, cand, cand_cap, cigar, cigar_cap, n_cigar);
}

// Mapping + the post-DP stage per shard: the SAM text of the mini-batch as pieces in input order.  Short reads: the whole
// stage runs on the device (gd_sr_map_sam_batch), only text crosses PCIe.  Long reads: gd_lr_map_batch, then the threaded
// host stage (gd_lr_sam_batch: concatenate_cigars and the logarithmic gap cost stay on the CPU) with post->n_threads /
// n_devices threads per shard.  The pieces BELONG TO THE HANDLE and stay valid until the call after the next one (a host
// writes batch i out while batch i+1 is mapped); the caller frees only the two arrays with gd_free.
template <class OPT, class FN>
static int multi_map_sam(gd_multi *m, int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                         const char *qual, const OPT *opt, FN map_fn, bool lr, const gd_sr_post_opt_t *post, int n_seq,
                         const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref, char ***parts,
                         size_t **part_len, int *n_parts)
{
	if (!m || n < 0 || !opt || !post || !parts || !part_len || !n_parts) return GD_ERR_ARG;
	const int G = (int)m->ctx.size();
	for (int i = 0; i < G; ++i)
		if (!m->idx[i]) {
			m->err = "gd_multi: no index (gd_multi_index_bcast first)";
			return GD_ERR_ARG;
		}
	*parts = nullptr, *part_len = nullptr, *n_parts = 0;
	const std::vector<int> cut = shard_cuts(n, len, G);
	std::vector<std::vector<std::pair<char *, size_t>>> out(G);
	std::vector<std::vector<char *>> host_made(G);
	const int gen = (int)(m->sam_calls++ & 1);
	for (char *p : m->owned[gen]) gd_free(p); // the pieces handed out two calls ago
	m->owned[gen].clear();
	gd_sr_post_opt_t po = *post;
	int total_threads = post->n_threads > 0 ? post->n_threads : (int)std::thread::hardware_concurrency();
	po.n_threads = std::max(1, total_threads / G);
	auto work = [&](int j) {
		gd_multi::Shard &S = m->sh[j];
		const int b = cut[j], cnt = cut[j + 1] - cut[j];
		S.rc = GD_OK, S.host_err = false;
		if (cnt == 0) return;
		cudaSetDevice(m->ctx[j]->device);
		if (!lr) { // short reads: reads in, text out
			char **pp = nullptr;
			size_t *pl = nullptr;
			int np = 0;
			S.rc = gd_sr_map_sam_batch(m->ctx[j], m->idx[j], cnt, names + b, off + b, len + b, seq, qual, (const gd_sr_opt_t *)opt, &po, n_seq,
			                           seq_names, &pp, &pl, &np);
			if (!S.rc) {
				for (int k = 0; k < np; ++k) out[j].push_back({pp[k], pl[k]});
				gd_free(pp), gd_free(pl);
			}
			return;
		}
		// Long reads: the shard goes through the device in slices of >= GD_LR_SLICE_READS reads and >= GD_LR_SLICE_BASES bases
		// (smaller launches leave the DP kernel short of pairs), and the host SAM stage of slice s runs on its own thread while
		// slice s+1 is on the device.  Every slice keeps its own records, so nothing is shared between the two sides.
		std::vector<int> sc(1, 0);
		{
			const char *e1 = getenv("GDIET_LR_SLICE_READS"), *e2 = getenv("GDIET_LR_SLICE_BASES");
			const int64_t min_reads = e1 ? std::max(1ll, atoll(e1)) : GD_LR_SLICE_READS, min_bases = e2 ? atoll(e2) : GD_LR_SLICE_BASES;
			int64_t bases = 0, total = 0;
			for (int i = 0; i < cnt; ++i) total += len[b + i];
			for (int i = 0, r = 0; i < cnt; ++i) {
				bases += len[b + i], total -= len[b + i], ++r;
				if (r >= min_reads && bases >= min_bases && cnt - 1 - i >= min_reads && total >= min_bases)
					sc.push_back(i + 1), bases = 0, r = 0;
			}
			sc.push_back(cnt);
		}
		const int ns = (int)sc.size() - 1;
		struct Slice {
			std::vector<int64_t> coff;
			gd_sr_cand_t *cand = nullptr; // malloc'ed, not value-initialised: only the pages the records reach are touched
			uint32_t *cig = nullptr;
			int64_t cand_cap = 0, cig_cap = 0;
			char *txt = nullptr;
			size_t tl = 0;
			int rc = GD_OK;
			std::thread sam;
			~Slice() { free(cand), free(cig); }
		};
		std::vector<Slice> sl(ns);
		const bool prof = getenv("GD_MAP_PROFILE") != nullptr;
		auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
		for (int s = 0; s < ns && !S.rc; ++s) {
			Slice &L = sl[s];
			const int sb = b + sc[s], scnt = sc[s + 1] - sc[s];
			const double t0 = now();
			L.coff.assign((size_t)scnt + 1, 0);
			// a CIGAR of a long read holds up to a few operations per ten bases (8 % error ONT reads: ~0.16 per base); too small a
			// pool means mapping the slice twice, so start above that
			int64_t sbases = 0;
			for (int i = 0; i < scnt; ++i) sbases += len[sb + i];
			L.cand_cap = (int64_t)scnt * 4 + 64, L.cig_cap = sbases / 3 + (int64_t)scnt * 64 + 1024;
			for (int attempt = 0; attempt < 2; ++attempt) {
				free(L.cand), free(L.cig);
				L.cand = (gd_sr_cand_t *)malloc((size_t)L.cand_cap * sizeof(gd_sr_cand_t)), L.cig = (uint32_t *)malloc((size_t)L.cig_cap * 4);
				if (!L.cand || !L.cig) {
					S.rc = GD_ERR_ARG, S.host_err = true;
					break;
				}
				int64_t ncig = 0;
				S.rc = map_fn(m->ctx[j], m->idx[j], scnt, off + sb, len + sb, seq, opt, L.coff.data(), L.cand, L.cand_cap, L.cig, L.cig_cap, &ncig);
				if (S.rc != GD_ERR_CAPACITY) break;
				L.cand_cap = L.coff[scnt] + 16, L.cig_cap = ncig + 16;
			}
			if (S.rc) break;
			if (prof) fprintf(stderr, "[gd_multi] device %d slice %d/%d: %d reads mapped in %.3f s\n", m->ctx[j]->device, s + 1, ns, scnt, now() - t0);
			L.sam = std::thread([&, sb, scnt, s]() {
				Slice &M = sl[s];
				const double t1 = now();
				M.rc = gd_lr_sam_batch(scnt, names + sb, off + sb, len + sb, seq, qual, M.coff.data(), M.cand, M.cig, n_seq,
				                       seq_names, ref_off, ref_len, ref, &po, &M.txt, &M.tl, nullptr, nullptr);
				if (prof) fprintf(stderr, "[gd_multi] device %d slice %d/%d: host SAM stage %.3f s (%d threads)\n", m->ctx[j]->device, s + 1, ns, now() - t1, po.n_threads);
			});
		}
		for (Slice &L : sl)
			if (L.sam.joinable()) L.sam.join();
		for (Slice &L : sl) {
			if (L.txt) host_made[j].push_back(L.txt); // owned by the handle also on failure, freed two calls on
			if (!S.rc && L.rc) S.rc = L.rc, S.host_err = true;
			if (!S.rc && L.txt) out[j].push_back({L.txt, L.tl});
		}
	};
	std::vector<std::thread> th;
	for (int j = 1; j < G; ++j) th.emplace_back(work, j);
	work(0);
	for (std::thread &t : th) t.join();
	int rc = GD_OK;
	for (int j = 0; j < G && !rc; ++j)
		if (m->sh[j].rc) {
			rc = m->sh[j].rc;
			m->err = m->sh[j].host_err ? std::string("gd_multi: host SAM stage of shard ") + std::to_string(j) + " failed"
			                           : std::string("gd_multi: device ") + std::to_string(m->ctx[j]->device) + ": " + gd_strerror(m->ctx[j]);
		}
	size_t np = 0;
	for (int j = 0; j < G; ++j) np += out[j].size();
	for (int j = 0; j < G; ++j)
		for (char *p : host_made[j]) m->owned[gen].push_back(p);
	if (rc) return rc;
	*parts = (char **)malloc((np + 1) * sizeof(char *)), *part_len = (size_t *)malloc((np + 1) * sizeof(size_t));
	size_t k = 0;
	for (int j = 0; j < G; ++j)
		for (auto &pr : out[j]) (*parts)[k] = pr.first, (*part_len)[k] = pr.second, ++k;
	*n_parts = (int)np;
	return GD_OK;
}

extern "C" int gd_multi_prepare_sam(gd_multi *m, size_t text_bytes)
{
	if (!m) return GD_ERR_ARG;
	const int G = (int)m->ctx.size();
	std::vector<std::thread> th;
	for (int j = 0; j < G; ++j) th.emplace_back([&, j]() { gd_sr_map_sam_prepare(m->ctx[j], text_bytes / (size_t)G + 1); });
	for (std::thread &t : th) t.join();
	return GD_OK;
}

extern "C" int gd_multi_sr_map_sam(gd_multi *m, int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                                   const char *qual, const gd_sr_opt_t *opt, const gd_sr_post_opt_t *post, int n_seq,
                                   const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                                   char ***parts, size_t **part_len, int *n_parts)
{
	return multi_map_sam(m, n, names, off, len, seq, qual, opt, gd_sr_map_batch, false, post, n_seq, seq_names, ref_off, ref_len, ref, parts,
	                     part_len, n_parts);
}

extern "C" int gd_multi_lr_map_sam(gd_multi *m, int n, const char *const *names, const int64_t *off, const int32_t *len, const char *seq,
                                   const char *qual, const gd_lr_opt_t *opt, const gd_sr_post_opt_t *post, int n_seq,
                                   const char *const *seq_names, const int64_t *ref_off, const int32_t *ref_len, const char *ref,
                                   char ***parts, size_t **part_len, int *n_parts)
{
	return multi_map_sam(m, n, names, off, len, seq, qual, opt, gd_lr_map_batch, true, post, n_seq, seq_names, ref_off, ref_len, ref, parts,
	                     part_len, n_parts);
}
