// index.cu — GPU construction of the high-count filter table (ha_ft_gen,
// htab.cpp:1136) and of the minimizer position index (ha_pt_gen, htab.cpp:1232).
//
// The reference counts into 4096 khashl tables from a 3-stage CPU pipeline; here
// counting is sort-based: hash every k-mer / sketch every read on the GPU, radix
// sort the hashes (CUB; the sort is stable, which reproduces the index's
// read-id / in-read position order, htab.cpp:646-672), run-length the sorted
// keys, build the peak statistics from the count histogram (ha_analyze_count,
// hist.cpp:74 — a 4096-bin histogram, evaluated on the host) and pour the kept
// runs into one open-addressing table with 16-byte slots.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include "hb_internal.h"
#include "hb_bloom.cuh"

#define YAK_MAX_COUNT 4095 /* htab.cpp:13-15 */

struct TmpBufs {
	hb_ctx *ctx; std::vector<void *> p;
	TmpBufs(hb_ctx *c) : ctx(c) {}
	template <typename T> T *get(uint64_t n)
	{ // an index build that does not fit beside the pass workspaces takes their memory (they are scratch: the next pass grows them again)
		void *q = 0;
		if (cudaMalloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); hb_ws_release(ctx); if (cudaMalloc(&q, (n ? n : 1) * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return 0; } }
		p.push_back(q); return (T *)q;
	}
	void drop(void *q) { for (auto &x : p) if (x == q) { cudaFree(x); x = 0; } }
	void *steal(void *q) { for (auto &x : p) if (x == q) x = 0; return q; }
	~TmpBufs() { for (void *q : p) if (q) cudaFree(q); }
};
#define NEED(ptr) do { if (!(ptr)) { hb_set_err(ctx, HB_E_NOMEM, "device allocation failed at %s:%d", __FILE__, __LINE__); return HB_E_NOMEM; } } while (0)
static inline unsigned nblk(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

// ---- coverage peaks of the k-mer count histogram (host side, 4096 bins) ----
// What ha_analyze_count (hist.cpp:74-157) decides when no homozygous peak is imposed (both call sites pass -1, htab.cpp:1155, 1252):
//   * the histogram falls from the error k-mers to a trough; the highest bin right of the trough is the main peak;
//   * a lower local maximum on one side is a second peak if it reaches 5 % of the main peak and the valley between the two drops below 95 % of it
//     (on the right also: not beyond 2.5 x the main peak's count);
//   * a right-hand second peak makes the main peak the heterozygous one; otherwise a left-hand one is the heterozygous peak.
// Ties follow the reference's scan directions: the main peak is the first bin holding the maximum, the left candidate the one nearest to the main peak,
// the right candidate the one nearest to it as well.
struct HistPeaks { int hom, het; };
static HistPeaks hist_peaks(const int64_t *h, int n_bins, int first_bin)
{
	HistPeaks r = { -1, -1 };
	int trough = std::max(h[1] > 0 ? 1 : 2, first_bin);
	while (trough + 1 < n_bins && h[trough + 1] <= h[trough]) trough++;
	if (trough == n_bins - 1) return r; // never rises again: no peak
	int top = trough + 1;
	for (int i = top + 1; i < n_bins; i++) if (h[i] > h[top]) top = i;
	const int64_t top_v = h[top];
	auto is_local_max = [&](int i) { return h[i] >= h[i - 1] && h[i] >= h[i + 1]; };
	auto separated = [&](int peak, int from, int to) { // valley strictly between the two bins, against the candidate's height
		int64_t valley = top_v; for (int i = from + 1; i < to; i++) valley = std::min(valley, h[i]);
		return !((double)h[peak] < (double)top_v * 0.05 || (double)valley > (double)h[peak] * 0.95);
	};
	int left = -1, right = -1;
	for (int i = top - 1; i > trough; i--) if (is_local_max(i) && (left < 0 || h[i] > h[left])) left = i;
	if (left >= 0 && !separated(left, left, top)) left = -1;
	for (int i = top + 1; i < n_bins - 1; i++) if (is_local_max(i) && (right < 0 || h[i] > h[right])) right = i;
	if (right >= 0 && (!separated(right, top, right) || (double)right > (double)top * 2.5)) right = -1;
	if (right > 0) { r.hom = right; r.het = top; }
	else { r.hom = top; if (left > 0) r.het = left; }
	return r;
}

// ---- kernels ---------------------------------------------------------------
// every (HPC) k-mer of a read, hashed like yak_hash_long (htab.h:162-167):
// mz1_count_seq_buf_HPC / mz1_count_seq_buf (htab.cpp:608-645).  One thread per
// read; slot i of the read's slice is written or left as the ~0 filler.
// pbits != 0: only the hashes whose top pbits bits equal pid (chunked counting: one hash range per pass); cnt_only != NULL: count them per read
// instead of writing them (the write pass then gets koff = the scan of the counts)
__global__ void k_all_kmers(DevReads R, int k, int is_hpc, const uint64_t *__restrict__ koff, uint64_t *__restrict__ out, unsigned long long *__restrict__ n_real,
                            int pbits, uint64_t pid, uint32_t *__restrict__ cnt_only)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= R.n) return;
	const uint8_t *seq = R.packed + R.off[r]; int32_t len = (int32_t)R.len[r], i, l = 0, last = -1;
	uint64_t x0 = 0, x1 = 0, x2 = 0, x3 = 0, mask = (1ULL << k) - 1, shift = k - 1, *dst = cnt_only ? 0 : out + koff[r]; uint32_t w = 0;
	uint64_t ni = R.noff[r], ne = R.noff[r + 1]; int32_t next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX;
	for (i = 0; i < len; ++i) {
		int c = hb_base(seq, (uint64_t)i);
		if (i == next_n) { c = 4; ++ni; next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX; }
		if (c < 4) {
			if (!is_hpc || c != last) {
				x0 = (x0 << 1 | (uint64_t)(c & 1)) & mask; x1 = (x1 << 1 | (uint64_t)(c >> 1)) & mask;
				x2 = x2 >> 1 | (uint64_t)(1 - (c & 1)) << shift; x3 = x3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
				if (++l >= k) {
					const uint64_t h = x1 < x3 ? hb_hash64(x0) + hb_hash64(x1) : hb_hash64(x2) + hb_hash64(x3);
					if (!pbits || (h >> (64 - pbits)) == pid) { if (!cnt_only) dst[w] = h; w++; }
				}
				last = c;
			}
		} else { l = 0; last = -1; x0 = x1 = x2 = x3 = 0; }
	}
	if (cnt_only) cnt_only[r] = w;
	else if (w && n_real) atomicAdd(n_real, (unsigned long long)w);
}

__global__ void k_iota(uint64_t n, uint64_t *__restrict__ out) { uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = i; }

struct IsHead { const uint64_t *k; __host__ __device__ bool operator()(uint64_t i) const { return i == 0 || k[i] != k[i - 1]; } };

// run lengths -> 4096-bin histogram of counts saturated at 4095 (ha_ct_hist, htab.cpp:240)
// fp != NULL: counting behind the Bloom filter (hb_bloom.cuh): fp[j] = the k-mer's first occurrence was a false positive; a k-mer whose
// only counted occurrences are none never entered the table and is not in the histogram
__global__ void k_run_hist(uint64_t n_runs, const uint64_t *__restrict__ head, uint64_t n_real, const uint8_t *__restrict__ fp, unsigned long long *__restrict__ hist)
{
	__shared__ unsigned int sh[4096];
	for (int i = threadIdx.x; i < 4096; i += blockDim.x) sh[i] = 0;
	__syncthreads();
	for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_runs; j += (uint64_t)gridDim.x * blockDim.x) {
		uint64_t len = (j + 1 < n_runs ? head[j + 1] : n_real) - head[j];
		if (fp) { len = hb_bf_count(len, fp[j]); if (!len) continue; }
		atomicAdd(&sh[len > YAK_MAX_COUNT ? YAK_MAX_COUNT : (uint32_t)len], 1u);
	}
	__syncthreads();
	for (int i = threadIdx.x; i < 4096; i += blockDim.x) if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}
// kept-run length (0 when dropped)
__global__ void k_run_keep(uint64_t n_runs, const uint64_t *__restrict__ head, uint64_t n_real, uint32_t lo, uint32_t hi, const uint8_t *__restrict__ fp, uint32_t *__restrict__ keep_len)
{
	uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_runs) return;
	uint64_t len = (j + 1 < n_runs ? head[j + 1] : n_real) - head[j];
	if (fp) len = hb_bf_count(len, fp[j]);
	if (len > YAK_MAX_COUNT) len = YAK_MAX_COUNT;
	keep_len[j] = (len >= lo && len <= hi) ? (uint32_t)len : 0;
}
// filter table insert: key -> count (INT32_MAX above max_cnt; gen_hh, htab.cpp:1038-1062)
__global__ void k_ft_insert(uint64_t n_runs, const uint64_t *__restrict__ head, const uint64_t *__restrict__ keys, const uint32_t *__restrict__ keep_len, uint32_t max_cnt,
                            uint64_t mask, uint64_t *__restrict__ t_key, int32_t *__restrict__ t_val)
{
	uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_runs || keep_len[j] == 0) return;
	uint64_t key = keys[head[j]], b = hb_bucket(key, mask); int32_t v = keep_len[j] > max_cnt ? INT32_MAX : (int32_t)keep_len[j];
	while (atomicCAS(&t_val[b], 0, v) != 0) b = (b + 1) & mask;
	t_key[b] = key;
}
// position index: warp per kept run copies its ha_idxpos_t words and inserts the key
__global__ void k_pt_fill(uint64_t n_runs, const uint64_t *__restrict__ head, const uint64_t *__restrict__ keys, const uint64_t *__restrict__ infos, const uint32_t *__restrict__ keep_len,
                          const uint64_t *__restrict__ pos_off, uint64_t mask, unsigned long long *t_slot /* ulonglong2 as 2 words */, uint64_t *__restrict__ pos)
{
	uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
	if (j >= n_runs) return;
	uint32_t len = keep_len[j];
	if (len == 0) return;
	uint64_t h = head[j], o = pos_off[j];
	for (uint32_t i = lane; i < len; i += 32) pos[o + i] = infos[h + i];
	if (lane == 0) {
		uint64_t key = keys[h], b = hb_bucket(key, mask), val = o << 12 | len;
		while (atomicCAS(&t_slot[2 * b + 1], 0ull, (unsigned long long)val) != 0ull) b = (b + 1) & mask;
		t_slot[2 * b] = key;
	}
}
__global__ void k_ft_query(DevFt ft, uint64_t n, const uint64_t *__restrict__ h, int32_t *__restrict__ out)
{ uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = hb_ft_lookup(ft, h[i]); }
__global__ void k_pt_query(DevPt pt, uint64_t n, const uint64_t *__restrict__ h, uint32_t *__restrict__ cnt, uint64_t *__restrict__ off)
{ uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) { uint64_t o; cnt[i] = hb_pt_lookup(pt, h[i], &o); off[i] = o; } }
__global__ void k_pt_gather(DevPt pt, uint64_t n, const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ off, const uint64_t *__restrict__ dst_off, uint64_t *__restrict__ dst)
{
	uint64_t j = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
	if (j >= n) return;
	for (uint32_t i = lane; i < cnt[j]; i += 32) dst[dst_off[j] + i] = pt.pos[off[j] + i];
}

// ---- the Bloom filter of ha_ft_gen (hb_bloom.cuh) --------------------------------------------------------------------------
// key of a distinct k-mer (run j of the sorted hashes): its filter block, then the ordinal of its first occurrence (the pair sort is
// stable, so the first value of a run is the smallest ordinal)
__global__ void k_bf_key(uint64_t n_runs, const uint64_t *__restrict__ head, const uint64_t *__restrict__ keys, const uint64_t *__restrict__ ord, int bf_local, int tbits,
                         uint64_t *__restrict__ skey, uint64_t *__restrict__ ridx)
{
	uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_runs) return;
	skey[j] = hb_bf_block(keys[head[j]], bf_local) << tbits | ord[head[j]]; ridx[j] = j;
}
// one thread walks a block's k-mers in first-occurrence order through the block's 512 bits (8 registers): fp = all 4 bits set before
__global__ void k_bf_walk(uint64_t n_runs, const uint64_t *__restrict__ skey, const uint64_t *__restrict__ ridx, int tbits, const uint64_t *__restrict__ head,
                          const uint64_t *__restrict__ keys, int bf_local, uint8_t *__restrict__ fp)
{
	uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_runs) return;
	const uint64_t b = skey[i] >> tbits;
	if (i && (skey[i - 1] >> tbits) == b) return; // not the first k-mer of its block
	uint64_t st[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
	for (uint64_t j = i; j < n_runs && (skey[j] >> tbits) == b; j++) { const uint64_t r = ridx[j]; fp[r] = hb_bf_insert(st, keys[head[r]], bf_local) == 4; }
}

// ---- sorted keys -> run heads + histogram ----------------------------------
static int hist_of(hb_ctx *ctx, TmpBufs &tb, const uint64_t *d_head, uint64_t n_runs, uint64_t n_real, const uint8_t *d_fp, int64_t *h_hist)
{
	unsigned long long *d_hist = tb.get<unsigned long long>(4096); NEED(d_hist);
	HB_CUDA(cudaMemsetAsync(d_hist, 0, 4096 * 8, ctx->stream));
	if (n_runs) k_run_hist<<<std::min<unsigned>(nblk(n_runs, 256), ctx->sm_count * 8), 256, 0, ctx->stream>>>(n_runs, d_head, n_real, d_fp, d_hist);
	HB_CUDA(cudaGetLastError());
	unsigned long long hh[4096];
	HB_CUDA(cudaMemcpyAsync(hh, d_hist, sizeof(hh), cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	for (int i = 0; i < 4096; i++) h_hist[i] = (int64_t)hh[i];
	tb.drop(d_hist);
	return HB_OK;
}
// h_hist == NULL: run heads only
static int runs_of(hb_ctx *ctx, TmpBufs &tb, const uint64_t *d_keys, uint64_t n_real, uint64_t **d_head, uint64_t *n_runs, int64_t *h_hist)
{
	size_t tmpb = 0; uint64_t *d_nsel = tb.get<uint64_t>(1); NEED(d_nsel);
	cub::CountingInputIterator<uint64_t> cnt_it(0); IsHead pred = { d_keys };
	// pass 1 of DeviceSelect needs the output sized for the worst case; count heads first with a transform-reduce-free trick:
	uint64_t *d_out = tb.get<uint64_t>(n_real + 1); NEED(d_out);
	HB_CUDA(cub::DeviceSelect::If(0, tmpb, cnt_it, d_out, d_nsel, (int64_t)n_real, pred, ctx->stream));
	void *tmp = tb.get<uint8_t>(tmpb); NEED(tmp);
	HB_CUDA(cub::DeviceSelect::If(tmp, tmpb, cnt_it, d_out, d_nsel, (int64_t)n_real, pred, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(n_runs, d_nsel, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	tb.drop(tmp);
	*d_head = d_out;
	return h_hist ? hist_of(ctx, tb, d_out, *n_runs, n_real, 0, h_hist) : HB_OK;
}

// ---- exact counting in P passes, one hash range each (row a3 at sizes whose k-mers do not fit HBM at once) --------------------------------------
// ha_ft_gen counts in 2^12 sub-tables by the low hash bits and, above a size, in several rounds over the input (htab.cpp:1136-1169); here a pass takes the
// k-mers whose hash starts with pid (top pbits bits): counts per read, scan, write, sort, run lengths.  The cut-off needs the histogram of ALL ranges, so the
// passes run twice: histogram first, then the kept (key, count) pairs into the table.  Same table contents as the one-pass build (slots may differ).
struct U32ToU64i { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };
static int ft_gen_chunked(hb_ctx *ctx, int pbits, int *hom_cov)
{
	const int k = ctx->opt.k_mer_length; const uint64_t n = ctx->n_reads, P = 1ull << pbits; int rc;
	TmpBufs keep(ctx); uint32_t *d_cnt = keep.get<uint32_t>(n + 1); uint64_t *d_off = keep.get<uint64_t>(n + 2); NEED(d_cnt); NEED(d_off);
	int64_t hist[4096]; memset(hist, 0, sizeof(hist));
	int cutoff = 0, max_cnt = 0; uint64_t cap = 0;
	for (int phase = 0; phase < 2; phase++) {
		for (uint64_t pid = 0; pid < P; pid++) {
			TmpBufs tb(ctx);
			HB_CUDA(cudaMemsetAsync(d_cnt, 0, (n + 1) * 4, ctx->stream));
			{ ProfScope ps(ctx, "k_all_kmers"); k_all_kmers<<<nblk(n, 64), 64, 0, ctx->stream>>>(hb_dev_reads(ctx), k, ctx->opt.is_hpc, 0, 0, 0, pbits, pid, d_cnt); }
			HB_CUDA(cudaGetLastError());
			{ size_t tbs = 0; cub::TransformInputIterator<uint64_t, U32ToU64i, const uint32_t *> in(d_cnt, U32ToU64i());
				HB_CUDA(cub::DeviceScan::ExclusiveSum(0, tbs, in, d_off, (int64_t)(n + 1), ctx->stream));
				void *tmp = tb.get<uint8_t>(tbs + 256); NEED(tmp);
				HB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tbs, in, d_off, (int64_t)(n + 1), ctx->stream)); tb.drop(tmp); }
			uint64_t np = 0; HB_CUDA(cudaMemcpyAsync(&np, d_off + n, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
			if (!np) continue;
			uint64_t *d_a = tb.get<uint64_t>(np + 1), *d_b = tb.get<uint64_t>(np + 1); NEED(d_a); NEED(d_b);
			{ ProfScope ps(ctx, "k_all_kmers"); k_all_kmers<<<nblk(n, 64), 64, 0, ctx->stream>>>(hb_dev_reads(ctx), k, ctx->opt.is_hpc, d_off, d_a, 0, pbits, pid, 0); }
			HB_CUDA(cudaGetLastError());
			size_t tmpb = 0; cub::DoubleBuffer<uint64_t> db(d_a, d_b);
			HB_CUDA(cub::DeviceRadixSort::SortKeys(0, tmpb, db, (int64_t)np, 0, 64 - pbits, ctx->stream)); // (the top bits are the same in the whole range)
			void *tmp = tb.get<uint8_t>(tmpb + 256); NEED(tmp);
			{ ProfScope ps(ctx, "sort_kmers"); HB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmpb, db, (int64_t)np, 0, 64 - pbits, ctx->stream)); }
			tb.drop(tmp); tb.drop(db.Alternate());
			const uint64_t *d_keys = db.Current(); uint64_t *d_head = 0, n_runs = 0; int64_t hp[4096];
			if ((rc = runs_of(ctx, tb, d_keys, np, &d_head, &n_runs, phase == 0 ? hp : 0))) return rc;
			if (phase == 0) { for (int i = 0; i < 4096; i++) hist[i] += hp[i]; continue; }
			uint32_t *d_keep = tb.get<uint32_t>(n_runs + 1); NEED(d_keep);
			k_run_keep<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, np, cutoff < 1 ? 1 : (uint32_t)cutoff, YAK_MAX_COUNT, 0, d_keep);
			k_ft_insert<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, d_keys, d_keep, (uint32_t)max_cnt, cap - 1, ctx->d_ft_key, ctx->d_ft_val);
			HB_CUDA(cudaGetLastError()); HB_CUDA(cudaStreamSynchronize(ctx->stream));
		}
		if (phase == 0) {
			const HistPeaks pk = hist_peaks(hist, 4096, ctx->opt.min_hist_kmer_cnt); // htab.cpp:1155-1161
			if (hom_cov) *hom_cov = pk.hom;
			cutoff = (int)(pk.hom * ctx->opt.high_factor); if (cutoff > YAK_MAX_COUNT - 1) cutoff = YAK_MAX_COUNT - 1;
			max_cnt = ctx->opt.max_kmer_cnt; if (max_cnt > YAK_MAX_COUNT - 1) max_cnt = YAK_MAX_COUNT - 1;
			uint64_t n_keep = 0; for (int i = cutoff < 0 ? 0 : cutoff; i <= YAK_MAX_COUNT; i++) n_keep += (uint64_t)hist[i];
			ctx->ft_n = n_keep;
			if (!n_keep) break;
			cap = 64; while (cap < 2 * n_keep) cap <<= 1;
			HB_CUDA(cudaMalloc((void **)&ctx->d_ft_key, cap * 8)); HB_CUDA(cudaMalloc((void **)&ctx->d_ft_val, cap * 4));
			HB_CUDA(cudaMemsetAsync(ctx->d_ft_key, 0, cap * 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(ctx->d_ft_val, 0, cap * 4, ctx->stream));
			ctx->ft_cap = cap;
		}
	}
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

// ---- ha_ft_gen ---------------------------------------------------------------
extern "C" void hb_ft_destroy(hb_ctx_t *ctx)
{
	if (!ctx) return;
	cudaFree(ctx->d_ft_key); cudaFree(ctx->d_ft_val); ctx->d_ft_key = 0; ctx->d_ft_val = 0; ctx->ft_n = ctx->ft_cap = 0;
	ctx->sk_reads = 0; // a kept sketch was made with this filter table
}
extern "C" int hb_ft_size(const hb_ctx_t *ctx, uint64_t *n) { *n = ctx->ft_n; return HB_OK; }

extern "C" int hb_ft_gen(hb_ctx_t *ctx, int *hom_cov)
{
	cudaSetDevice(ctx->device);
	if (!ctx->n_reads) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	hb_ft_destroy(ctx);
	TmpBufs tb(ctx); const int k = ctx->opt.k_mer_length; const uint64_t n = ctx->n_reads;
	std::vector<uint64_t> koff(n + 1); uint64_t tot = 0;
	for (uint64_t i = 0; i < n; i++) { koff[i] = tot; tot += ctx->h_rlen[i] >= (uint32_t)k ? ctx->h_rlen[i] - k + 1 : 0; }
	koff[n] = tot;
	if (!hb_bf_active(ctx->opt.bf_shift)) { // exact counting: one pass when the hashes fit (two arrays of them + the run heads), else 2^pbits hash ranges
		int pbits = ctx->ft_chunk_bits;
		if (pbits < 0) {
			size_t fr = 0, tt = 0; cudaMemGetInfo(&fr, &tt); size_t avail = fr + ctx->ws_cap;
			for (int i = 0; i < HB_MAX_LANES - 1; i++) if (ctx->lane[i]) avail += ctx->lane[i]->ws_cap; // (the workspaces are given back when an allocation fails)
			for (pbits = 0; pbits < 12 && (double)(tot >> pbits) * 17.0 > (double)avail * 0.9; pbits++) {}
		}
		if (pbits > 0) return ft_gen_chunked(ctx, pbits, hom_cov);
	}
	uint64_t *d_koff = tb.get<uint64_t>(n + 1), *d_a = tb.get<uint64_t>(tot + 1), *d_b = tb.get<uint64_t>(tot + 1); unsigned long long *d_nreal = tb.get<unsigned long long>(1);
	NEED(d_koff); NEED(d_a); NEED(d_b); NEED(d_nreal);
	HB_CUDA(cudaMemcpyAsync(d_koff, koff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemsetAsync(d_a, 0xff, (tot + 1) * 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_nreal, 0, 8, ctx->stream));
	{ ProfScope ps(ctx, "k_all_kmers"); k_all_kmers<<<nblk(n, 64), 64, 0, ctx->stream>>>(hb_dev_reads(ctx), k, ctx->opt.is_hpc, d_koff, d_a, d_nreal, 0, 0, 0); }
	HB_CUDA(cudaGetLastError());
	unsigned long long n_real = 0; HB_CUDA(cudaMemcpyAsync(&n_real, d_nreal, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	size_t tmpb = 0; cub::DoubleBuffer<uint64_t> db(d_a, d_b);
	const int bf_shift = ctx->opt.bf_shift, bf_local = bf_shift - 12; const bool use_bf = hb_bf_active(bf_shift);
	uint64_t *d_keys = 0, *d_head = 0, n_runs = 0; int64_t hist[4096]; uint8_t *d_fp = 0; int rc;
	if (!use_bf) {
		HB_CUDA(cub::DeviceRadixSort::SortKeys(0, tmpb, db, (int64_t)tot, 0, 64, ctx->stream));
		void *tmp = tb.get<uint8_t>(tmpb); NEED(tmp);
		{ ProfScope ps(ctx, "sort_kmers"); HB_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tmpb, db, (int64_t)tot, 0, 64, ctx->stream)); }
		tb.drop(tmp); tb.drop(db.Alternate());
		d_keys = db.Current();
		rc = runs_of(ctx, tb, d_keys, n_real, &d_head, &n_runs, hist); if (rc) return rc; // the ~0 fillers sort last and are not counted
	} else {
		// the reference's Bloom filter (hb_bloom.cuh): the counts depend on WHEN a k-mer first occurs, so the hashes travel with their ordinal
		// (slot in the (read, position)-ordered array) through a stable pair sort
		int tbits = 1; while (tbits < 64 && (tot >> tbits)) tbits++;
		if (12 + (bf_local - 9) + tbits > 64) { hb_set_err(ctx, HB_E_ARG, "Bloom-filtered counting (-f %d) of %llu k-mers does not fit one pass (chunked counting is not built)", bf_shift, (unsigned long long)tot); return HB_E_ARG; }
		uint64_t *d_va = tb.get<uint64_t>(tot + 1), *d_vb = tb.get<uint64_t>(tot + 1); NEED(d_va); NEED(d_vb);
		k_iota<<<nblk(tot, 256), 256, 0, ctx->stream>>>(tot, d_va);
		cub::DoubleBuffer<uint64_t> dv(d_va, d_vb);
		HB_CUDA(cub::DeviceRadixSort::SortPairs(0, tmpb, db, dv, (int64_t)tot, 0, 64, ctx->stream));
		void *tmp = tb.get<uint8_t>(tmpb); NEED(tmp);
		{ ProfScope ps(ctx, "sort_kmers"); HB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpb, db, dv, (int64_t)tot, 0, 64, ctx->stream)); }
		tb.drop(tmp); tb.drop(db.Alternate()); tb.drop(dv.Alternate());
		d_keys = db.Current(); const uint64_t *d_ord = dv.Current();
		rc = runs_of(ctx, tb, d_keys, n_real, &d_head, &n_runs, 0); if (rc) return rc;
		uint64_t *d_sk = tb.get<uint64_t>(n_runs + 1), *d_sk2 = tb.get<uint64_t>(n_runs + 1), *d_ri = tb.get<uint64_t>(n_runs + 1), *d_ri2 = tb.get<uint64_t>(n_runs + 1);
		d_fp = tb.get<uint8_t>(n_runs + 1); NEED(d_sk); NEED(d_sk2); NEED(d_ri); NEED(d_ri2); NEED(d_fp);
		if (n_runs) {
			k_bf_key<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, d_keys, d_ord, bf_local, tbits, d_sk, d_ri);
			cub::DoubleBuffer<uint64_t> bk(d_sk, d_sk2), bv(d_ri, d_ri2); size_t tb2 = 0;
			HB_CUDA(cub::DeviceRadixSort::SortPairs(0, tb2, bk, bv, (int64_t)n_runs, 0, 64, ctx->stream));
			void *tmp2 = tb.get<uint8_t>(tb2); NEED(tmp2);
			HB_CUDA(cub::DeviceRadixSort::SortPairs(tmp2, tb2, bk, bv, (int64_t)n_runs, 0, 64, ctx->stream));
			{ ProfScope ps(ctx, "k_bf_walk"); k_bf_walk<<<nblk(n_runs, 128), 128, 0, ctx->stream>>>(n_runs, bk.Current(), bv.Current(), tbits, d_head, d_keys, bf_local, d_fp); }
			HB_CUDA(cudaGetLastError()); HB_CUDA(cudaStreamSynchronize(ctx->stream));
			tb.drop(tmp2);
		}
		tb.drop(d_sk); tb.drop(d_sk2); tb.drop(d_ri); tb.drop(d_ri2); tb.drop((void *)d_ord);
		rc = hist_of(ctx, tb, d_head, n_runs, n_real, d_fp, hist); if (rc) return rc;
	}
	const HistPeaks pk = hist_peaks(hist, 4096, ctx->opt.min_hist_kmer_cnt); const int peak_het = pk.het, peak_hom = pk.hom; // htab.cpp:1155-1161
	if (hom_cov) *hom_cov = peak_hom;
	int cutoff = (int)(peak_hom * ctx->opt.high_factor); if (cutoff > YAK_MAX_COUNT - 1) cutoff = YAK_MAX_COUNT - 1;
	int max_cnt = ctx->opt.max_kmer_cnt; if (max_cnt > YAK_MAX_COUNT - 1) max_cnt = YAK_MAX_COUNT - 1;
	uint64_t n_keep = 0; for (int i = cutoff < 0 ? 0 : cutoff; i <= YAK_MAX_COUNT; i++) n_keep += (uint64_t)hist[i]; // ha_ct_shrink(h, cutoff, YAK_MAX_COUNT)
	ctx->ft_n = n_keep;
	if (n_keep) {
		uint64_t cap = 64; while (cap < 2 * n_keep) cap <<= 1;
		uint32_t *d_keep = tb.get<uint32_t>(n_runs + 1); NEED(d_keep);
		HB_CUDA(cudaMalloc((void **)&ctx->d_ft_key, cap * 8)); HB_CUDA(cudaMalloc((void **)&ctx->d_ft_val, cap * 4));
		HB_CUDA(cudaMemsetAsync(ctx->d_ft_key, 0, cap * 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(ctx->d_ft_val, 0, cap * 4, ctx->stream));
		k_run_keep<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, n_real, cutoff < 1 ? 1 : (uint32_t)cutoff, YAK_MAX_COUNT, d_fp, d_keep);
		k_ft_insert<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, d_keys, d_keep, (uint32_t)max_cnt, cap - 1, ctx->d_ft_key, ctx->d_ft_val);
		HB_CUDA(cudaGetLastError());
		ctx->ft_cap = cap;
	}
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

extern "C" int hb_ft_cnt(hb_ctx_t *ctx, const uint64_t *hash, uint64_t n, int32_t *cnt)
{
	cudaSetDevice(ctx->device); TmpBufs tb(ctx);
	if (n == 0) return HB_OK;
	uint64_t *d_h = tb.get<uint64_t>(n); int32_t *d_c = tb.get<int32_t>(n); NEED(d_h); NEED(d_c);
	HB_CUDA(cudaMemcpyAsync(d_h, hash, n * 8, cudaMemcpyHostToDevice, ctx->stream));
	k_ft_query<<<nblk(n, 256), 256, 0, ctx->stream>>>(hb_dev_ft(ctx), n, d_h, d_c);
	HB_CUDA(cudaMemcpyAsync(cnt, d_c, n * 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

// ---- ha_pt_gen ---------------------------------------------------------------
extern "C" void hb_pt_destroy(hb_ctx_t *ctx)
{
	if (!ctx) return;
	cudaFree(ctx->d_pt_slot); cudaFree(ctx->d_pt_pos); ctx->d_pt_slot = 0; ctx->d_pt_pos = 0; ctx->pt_keys = ctx->pt_npos = ctx->pt_cap = 0;
	ctx->sk_reads = 0; // the kept sketch belongs to the index
}
extern "C" int hb_pt_stat(const hb_ctx_t *ctx, uint64_t *n_keys, uint64_t *n_pos) { *n_keys = ctx->pt_keys; *n_pos = ctx->pt_npos; return HB_OK; }

extern "C" int hb_pt_gen(hb_ctx_t *ctx, int *hom_cov, int *het_cov)
{
	cudaSetDevice(ctx->device);
	if (!ctx->n_reads) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	hb_pt_destroy(ctx);
	TmpBufs tb(ctx); DevSketch sk; int rc = hb_run_sketch_retry(ctx, 0, ctx->n_reads, &sk); if (rc) return rc;
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	const uint64_t n = sk.total;
	if (ctx->sk_reuse) { // keep the minimizers for the pass over this index (engine.cu run_pass_impl); without room for the copy the pass simply sketches again
		const uint64_t nr = ctx->n_reads; bool ok = true;
		if (ctx->sk_mz_cap < n + 2) { cudaFree(ctx->d_sk_mz); ctx->d_sk_mz = 0; ctx->sk_mz_cap = 0; if (cudaMalloc((void **)&ctx->d_sk_mz, (n + 2 + (n >> 4)) * sizeof(hb_mz_t)) == cudaSuccess) ctx->sk_mz_cap = n + 2 + (n >> 4); else { cudaGetLastError(); ok = false; } }
		if (ok && ctx->sk_off_cap < nr + 2) { cudaFree(ctx->d_sk_off); ctx->d_sk_off = 0; ctx->sk_off_cap = 0; if (cudaMalloc((void **)&ctx->d_sk_off, (nr + 2) * 8) == cudaSuccess) ctx->sk_off_cap = nr + 2; else { cudaGetLastError(); ok = false; } }
		if (ok) {
			ctx->h_sk_off.resize(nr + 1);
			HB_CUDA(cudaMemcpyAsync(ctx->d_sk_mz, sk.mz, (n + 2) * sizeof(hb_mz_t), cudaMemcpyDeviceToDevice, ctx->stream)); // (+ the two pad entries behind the array)
			HB_CUDA(cudaMemcpyAsync(ctx->d_sk_off, sk.off, (nr + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
			HB_CUDA(cudaMemcpyAsync(ctx->h_sk_off.data(), sk.off, (nr + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			HB_CUDA(cudaStreamSynchronize(ctx->stream));
			ctx->sk_total = n; ctx->sk_reads = nr;
		}
	}
	// split {x,info} into key / value arrays and sort by key (stable LSD radix)
	uint64_t *d_k0 = tb.get<uint64_t>(n + 1), *d_k1 = tb.get<uint64_t>(n + 1), *d_v0 = tb.get<uint64_t>(n + 1), *d_v1 = tb.get<uint64_t>(n + 1);
	NEED(d_k0); NEED(d_k1); NEED(d_v0); NEED(d_v1);
	HB_CUDA(cudaMemcpy2DAsync(d_k0, 8, &sk.mz[0].x, 16, 8, n, cudaMemcpyDeviceToDevice, ctx->stream));
	HB_CUDA(cudaMemcpy2DAsync(d_v0, 8, &sk.mz[0].info, 16, 8, n, cudaMemcpyDeviceToDevice, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	hb_ws_reset(ctx);
	size_t tmpb = 0; cub::DoubleBuffer<uint64_t> dk(d_k0, d_k1), dv(d_v0, d_v1);
	HB_CUDA(cub::DeviceRadixSort::SortPairs(0, tmpb, dk, dv, (int64_t)n, 0, 64, ctx->stream));
	void *tmp = tb.get<uint8_t>(tmpb); NEED(tmp);
	{ ProfScope ps(ctx, "sort_minimizers"); HB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmpb, dk, dv, (int64_t)n, 0, 64, ctx->stream)); }
	tb.drop(tmp); tb.drop(dk.Alternate()); tb.drop(dv.Alternate());
	uint64_t *d_keys = dk.Current(), *d_infos = dv.Current(), *d_head = 0, n_runs = 0; int64_t hist[4096];
	rc = runs_of(ctx, tb, d_keys, n, &d_head, &n_runs, hist); if (rc) return rc;
	const HistPeaks pk = hist_peaks(hist, 4096, ctx->opt.min_hist_kmer_cnt); const int peak_het = pk.het, peak_hom = pk.hom; // htab.cpp:1252-1257
	if (hom_cov) *hom_cov = peak_hom;
	if (het_cov) *het_cov = peak_het;
	int lo = 2, hi = YAK_MAX_COUNT - 1; // htab.cpp:1259-1270
	if (ctx->ft_n == 0 && ctx->d_ft_key == 0 && ctx->no_kmer_flt) { /* HA_F_NO_KMER_FLT without a filter table (htab.cpp:1261) */ hi = (int)(peak_hom * ctx->opt.high_factor); if (hi > YAK_MAX_COUNT - 1) hi = YAK_MAX_COUNT - 1; }
	uint64_t n_keys = 0, n_pos = 0;
	for (int i = lo; i <= hi; i++) { n_keys += (uint64_t)hist[i]; n_pos += (uint64_t)hist[i] * i; }
	uint64_t cap = 64; while (cap < 2 * n_keys) cap <<= 1;
	uint32_t *d_keep = tb.get<uint32_t>(n_runs + 2); uint64_t *d_poff = tb.get<uint64_t>(n_runs + 2); NEED(d_keep); NEED(d_poff);
	HB_CUDA(cudaMemsetAsync(d_keep, 0, (n_runs + 2) * 4, ctx->stream));
	if (n_runs) k_run_keep<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_head, n, (uint32_t)lo, (uint32_t)hi, 0, d_keep);
	rc = hb_scan_u32_to_u64(ctx, d_keep, d_poff, n_runs); if (rc) return rc;
	HB_CUDA(cudaMalloc((void **)&ctx->d_pt_slot, cap * 16)); HB_CUDA(cudaMalloc((void **)&ctx->d_pt_pos, (n_pos + 4) * 8));
	HB_CUDA(cudaMemsetAsync(ctx->d_pt_slot, 0, cap * 16, ctx->stream)); HB_CUDA(cudaMemsetAsync(ctx->d_pt_pos, 0, (n_pos + 4) * 8, ctx->stream));
	if (n_runs) { ProfScope ps(ctx, "k_pt_fill"); k_pt_fill<<<nblk(n_runs * 32, 256), 256, 0, ctx->stream>>>(n_runs, d_head, d_keys, d_infos, d_keep, d_poff, cap - 1, (unsigned long long *)ctx->d_pt_slot, ctx->d_pt_pos); }
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->pt_keys = n_keys; ctx->pt_npos = n_pos; ctx->pt_cap = cap;
	return HB_OK;
}

extern "C" int hb_pt_get(hb_ctx_t *ctx, const uint64_t *hash, uint64_t n, uint32_t *cnt, uint64_t *pos, uint64_t pos_cap)
{
	cudaSetDevice(ctx->device); TmpBufs tb(ctx);
	if (!ctx->d_pt_slot) { hb_set_err(ctx, HB_E_STATE, "no position index"); return HB_E_STATE; }
	if (n == 0) return HB_OK;
	uint64_t *d_h = tb.get<uint64_t>(n), *d_o = tb.get<uint64_t>(n), *d_do = tb.get<uint64_t>(n + 2); uint32_t *d_c = tb.get<uint32_t>(n + 1);
	NEED(d_h); NEED(d_o); NEED(d_do); NEED(d_c);
	HB_CUDA(cudaMemcpyAsync(d_h, hash, n * 8, cudaMemcpyHostToDevice, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_c, 0, (n + 1) * 4, ctx->stream));
	k_pt_query<<<nblk(n, 256), 256, 0, ctx->stream>>>(hb_dev_pt(ctx), n, d_h, d_c, d_o);
	HB_CUDA(cudaMemcpyAsync(cnt, d_c, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	if (pos) {
		int rc = hb_scan_u32_to_u64(ctx, d_c, d_do, n); if (rc) return rc;
		uint64_t tot = 0; HB_CUDA(cudaMemcpyAsync(&tot, d_do + n, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
		if (tot > pos_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "position output capacity: need %llu", (unsigned long long)tot); return HB_E_OVERFLOW; }
		uint64_t *d_p = tb.get<uint64_t>(tot + 1); NEED(d_p);
		k_pt_gather<<<nblk(n * 32, 256), 256, 0, ctx->stream>>>(hb_dev_pt(ctx), n, d_c, d_o, d_do, d_p);
		HB_CUDA(cudaMemcpyAsync(pos, d_p, tot * 8, cudaMemcpyDeviceToHost, ctx->stream));
	}
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}
