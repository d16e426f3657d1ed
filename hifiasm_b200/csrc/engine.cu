// engine.cu — context, read store, and the host orchestration of the overlap
// pass (C-ABI of include/hifiasm_b200.h).  Host code is plumbing only: buffer
// management, launches, batch planning.  No compute runs on the host and there
// is no CPU fallback.
#include <stdarg.h>
#include <math.h>
#include <algorithm>
#include <thread>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>
#include "hb_internal.h"
#define HB_KERNELS_MAIN
#include "hb_kernels.cuh"
#include "hb_kernels_ea.cuh"
#include "hb_rechain_launch.h"

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
void hb_set_err(hb_ctx *ctx, int code, const char *fmt, ...)
{
	char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof(buf), fmt, ap); va_end(ap);
	if (ctx) ctx->err = buf;
	(void)code;
}

ProfScope::ProfScope(hb_ctx *c, const char *n) : ctx(c), name(n)
{
	cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream);
}
ProfScope::~ProfScope()
{
	float ms = 0; cudaEventRecord(b, ctx->stream); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
	cudaEventDestroy(a); cudaEventDestroy(b);
	if (ctx->in_stage) { bool f = false; for (auto &p : ctx->prof_stage) if (p.name == name) { p.launches++; p.ms += ms; f = true; break; } if (!f) { ProfEntry e = { name, 1, ms }; ctx->prof_stage.push_back(e); } }
	for (auto &p : ctx->prof) if (p.name == name) { p.launches++; p.ms += ms; return; }
	ProfEntry e = { name, 1, ms }; ctx->prof.push_back(e);
}
void hb_prof_reset(hb_ctx *ctx)
{ // start of a pass: the per-pass table and counters restart; inside hb_stage_run the finished pass's counters are added to the stage's
	if (ctx->in_stage) for (int i = 0; i < 12; i++) ctx->stage_counters[i] += ctx->counters[i];
	ctx->prof.clear(); memset(ctx->counters, 0, sizeof(ctx->counters));
}
void hb_stage_prof_begin(hb_ctx *ctx) { ctx->prof_stage.clear(); memset(ctx->stage_counters, 0, sizeof(ctx->stage_counters)); memset(ctx->counters, 0, sizeof(ctx->counters)); ctx->in_stage = 1; }
void hb_stage_prof_end(hb_ctx *ctx) { for (int i = 0; i < 12; i++) ctx->stage_counters[i] += ctx->counters[i]; ctx->in_stage = 0; }

DevReads hb_dev_reads(const hb_ctx *ctx)
{
	DevReads R; R.n = ctx->n_reads; R.packed = ctx->d_packed; R.off = ctx->d_roff; R.len = ctx->d_rlen; R.noff = ctx->d_noff; R.npos = ctx->d_npos;
	return R;
}
DevFt hb_dev_ft(const hb_ctx *ctx)
{
	DevFt f; f.mask = ctx->ft_n ? ctx->ft_cap - 1 : 0; f.key = ctx->d_ft_key; f.val = ctx->d_ft_val;
	return f;
}
DevPt hb_dev_pt(const hb_ctx *ctx)
{
	DevPt p; p.mask = ctx->pt_cap - 1; p.slot = ctx->d_pt_slot; p.pos = ctx->d_pt_pos;
	return p;
}

// ---------------------------------------------------------------------------
// workspace: no allocator calls inside a pass
// ---------------------------------------------------------------------------
void hb_ws_reset(hb_ctx *ctx) { ctx->ws_lo = 0; ctx->ws_hi = ctx->ws_cap; ctx->ws_virt = 0; }
static void *ws_short(hb_ctx *ctx, size_t bytes)
{ // a request that does not fit: keep counting what the caller goes on to ask for, so that ONE regrow is enough
	ctx->ws_virt += bytes + 256;
	size_t need = ctx->ws_lo + (ctx->ws_cap - ctx->ws_hi) + ctx->ws_virt; if (need > ctx->ws_need) ctx->ws_need = need;
	return 0;
}
static const bool g_trace_ws = getenv("HB_TRACE_WS") != 0; // (diagnostics: one process-wide read)
void *hb_ws_lo(hb_ctx *ctx, size_t bytes)
{
	bytes = (bytes + 255) & ~(size_t)255;
	if (g_trace_ws && bytes >= ((size_t)256 << 20)) fprintf(stderr, "[hb ws] lo %.2f GB at %.2f GB (cap %.2f)\n", bytes / 1e9, ctx->ws_lo / 1e9, ctx->ws_cap / 1e9);
	if (ctx->ws_virt || ctx->ws_lo + bytes > ctx->ws_hi) return ws_short(ctx, bytes);
	void *p = ctx->ws + ctx->ws_lo; ctx->ws_lo += bytes; return p;
}
void *hb_ws_hi(hb_ctx *ctx, size_t bytes)
{
	bytes = (bytes + 255) & ~(size_t)255;
	if (g_trace_ws && bytes >= ((size_t)256 << 20)) fprintf(stderr, "[hb ws] hi %.2f GB (hi at %.2f of %.2f GB)\n", bytes / 1e9, ctx->ws_hi / 1e9, ctx->ws_cap / 1e9);
	if (ctx->ws_virt || ctx->ws_lo + bytes + 256 > ctx->ws_hi) return ws_short(ctx, bytes);
	ctx->ws_hi = (ctx->ws_hi - bytes) & ~(size_t)255; return ctx->ws + ctx->ws_hi;
}
void *hb_ws_hi_packed(hb_ctx *ctx, size_t bytes)
{ // 16-byte granularity: consecutive calls return adjacent blocks (descending addresses)
	bytes = (bytes + 15) & ~(size_t)15;
	if (ctx->ws_virt || ctx->ws_lo + bytes > ctx->ws_hi) return ws_short(ctx, bytes);
	ctx->ws_hi -= bytes; return ctx->ws + ctx->ws_hi;
}
int hb_ws_grow(hb_ctx *ctx)
{ // called with the stack empty: replace the workspace by a larger one
	size_t want = ctx->ws_cap ? ctx->ws_cap + (ctx->ws_cap >> 1) : ((size_t)256 << 20), fr = 0, tot = 0;
	if (ctx->ws_need) want = ctx->ws_need + (ctx->ws_need >> 3);
	if (want < ((size_t)256 << 20)) want = (size_t)256 << 20;
	want = (want + 255) & ~(size_t)255;
	cudaStreamSynchronize(ctx->stream);
	if (ctx->ws) cudaFree(ctx->ws);
	ctx->ws = 0; ctx->ws_cap = 0;
	cudaMemGetInfo(&fr, &tot);
	if (want > fr - (fr >> 4)) want = (fr - (fr >> 4)) & ~(size_t)255;
	if (want < ctx->ws_need) { hb_set_err(ctx, HB_E_NOMEM, "workspace of %zu bytes does not fit in free HBM (%zu); lower HB_ANCHOR_BUDGET", ctx->ws_need, fr); return HB_E_NOMEM; }
	if (cudaMalloc((void **)&ctx->ws, want) != cudaSuccess) { cudaGetLastError(); hb_set_err(ctx, HB_E_NOMEM, "cudaMalloc(%zu) for the workspace failed", want); return HB_E_NOMEM; }
	ctx->ws_cap = want; ctx->ws_need = 0; hb_ws_reset(ctx);
	return HB_OK;
}

// scoped scratch from the low end of the workspace (rewound on scope exit); hi<T>() takes
// from the high end and lives until the pass ends
struct Arena {
	hb_ctx *ctx; size_t mark; bool failed; std::vector<void *> ptrs; // ptrs: kept for source compatibility, unused
	Arena(hb_ctx *c) : ctx(c), mark(c->ws_lo), failed(false) {}
	template <typename T> T *get(uint64_t n) { void *p = hb_ws_lo(ctx, (n ? n : 1) * sizeof(T)); if (!p) failed = true; return (T *)p; }
	template <typename T> T *zero(uint64_t n) { T *p = get<T>(n); if (p) cudaMemsetAsync(p, 0, (n ? n : 1) * sizeof(T), ctx->stream); return p; }
	template <typename T> T *hi(uint64_t n) { void *p = hb_ws_hi(ctx, (n ? n : 1) * sizeof(T)); if (!p) failed = true; return (T *)p; }
	void release(void *) {}
	~Arena() { ctx->ws_lo = mark; }
};
#define HB_ALLOC_CHECK(ar) do { if ((ar).failed) return HB_E_WS; } while (0)

static inline unsigned nblk(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

struct U32ToU64 { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };
int hb_scan_u32_to_u64(hb_ctx *ctx, const uint32_t *d_in, uint64_t *d_out, uint64_t n)
{ // exclusive sum over n+1 items (the last input is ignored by callers that pad with 0)
	size_t tb = 0; Arena ar(ctx);
	cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t *> in(d_in, U32ToU64());
	HB_CUDA(cub::DeviceScan::ExclusiveSum(0, tb, in, d_out, n + 1, ctx->stream));
	void *tmp = ar.get<uint8_t>(tb); HB_ALLOC_CHECK(ar);
	HB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, in, d_out, n + 1, ctx->stream));
	return HB_OK;
}

// ---------------------------------------------------------------------------
// options / context
// ---------------------------------------------------------------------------
extern "C" void hb_opt_init(hb_opt_t *o)
{ // init_opt, CommandLines.cpp:243-380
	memset(o, 0, sizeof(*o));
	o->k_mer_length = 51; o->mz_win = 51; o->is_hpc = 1; o->mz_sample_dist = 500; o->mz_rewin = 1000;
	o->min_hist_kmer_cnt = 5; o->max_kmer_cnt = 2000; o->max_n_chain = 100; o->high_factor = 5.0; o->hom_cov = 20; o->het_cov = -1024;
}
extern "C" void hb_opt_update_cov(hb_opt_t *o, int hom_cov)
{ // ha_opt_update_cov, CommandLines.cpp:411-418
	int m = (int)(hom_cov * o->high_factor + .499);
	o->hom_cov = hom_cov;
	if (o->max_n_chain < m) o->max_n_chain = m;
}
extern "C" int hb_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; } return n; }

static int check_opt(hb_ctx *ctx, const hb_opt_t *o)
{
	if (o->k_mer_length < 2 || o->k_mer_length > 63 || o->mz_win < 1 || o->mz_win > 255) { hb_set_err(ctx, HB_E_ARG, "k must be in [2,63], w in [1,255]"); return HB_E_ARG; }
	if (o->is_ont) { hb_set_err(ctx, HB_E_ARG, "--ont mode is not implemented on the GPU path yet"); return HB_E_ARG; }
	return HB_OK;
}

extern "C" int hb_create(hb_ctx_t **out, int device, const hb_opt_t *opt)
{
	int n = hb_device_count();
	*out = 0;
	if (n <= 0 || device < 0 || device >= n) return HB_E_NO_DEVICE; // no CUDA device: fail loudly, there is no CPU path
	hb_ctx *ctx = new hb_ctx();
	ctx->device = device;
	if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreate(&ctx->stream) != cudaSuccess) { delete ctx; return HB_E_CUDA; }
	cudaDeviceProp pr; cudaGetDeviceProperties(&pr, device); ctx->sm_count = pr.multiProcessorCount;
	cudaMemPool_t pool; cudaDeviceGetDefaultMemPool(&pool, device);
	uint64_t thr = UINT64_MAX; cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
	if (opt) ctx->opt = *opt; else hb_opt_init(&ctx->opt);
	if (check_opt(ctx, &ctx->opt)) { cudaStreamDestroy(ctx->stream); delete ctx; return HB_E_ARG; }
	ctx->n_reads = 0; ctx->d_packed = 0; ctx->d_roff = 0; ctx->d_rlen = 0; ctx->d_noff = 0; ctx->d_npos = 0;
	ctx->ft_n = ctx->ft_cap = 0; ctx->d_ft_key = 0; ctx->d_ft_val = 0;
	ctx->pt_keys = ctx->pt_npos = ctx->pt_cap = 0; ctx->d_pt_slot = 0; ctx->d_pt_pos = 0;
	ctx->d_prev0 = ctx->d_prev1 = 0; ctx->d_prev0_off = ctx->d_prev1_off = 0; ctx->n_prev0 = ctx->n_prev1 = 0;
	ctx->d_out0 = ctx->d_out1 = 0; ctx->d_out0_off = ctx->d_out1_off = 0; ctx->n_out0 = ctx->n_out1 = ctx->out_reads = 0;
	ctx->anchor_budget = 96ull << 20; ctx->last_pass_ms = 0; ctx->stage_buf = 0; ctx->in_stage = 0; memset(ctx->stage_counters, 0, sizeof(ctx->stage_counters));
	ctx->ws = 0; ctx->ws_cap = ctx->ws_lo = ctx->ws_hi = ctx->ws_need = ctx->ws_virt = 0; ctx->h_stage = 0; ctx->h_stage_cap = 0; ctx->packed_cap = ctx->reads_cap = ctx->npos_cap = 0; ctx->out0_cap = ctx->out1_cap = ctx->outoff_cap = 0;
	const char *e = getenv("HB_ANCHOR_BUDGET"); if (e) ctx->anchor_budget = strtoull(e, 0, 10);
	ctx->ecb_path_words = getenv("HB_ECB_PATH_WORDS") ? strtoull(getenv("HB_ECB_PATH_WORDS"), 0, 10) : 4096; ctx->ecb_cig_words = getenv("HB_ECB_CIG_WORDS") ? atoi(getenv("HB_ECB_CIG_WORDS")) : 4096;
	ctx->trace = getenv("HB_TRACE") != 0; ctx->trace_ec = getenv("HB_TRACE_EC") != 0; ctx->no_kmer_flt = getenv("HB_NO_KMER_FLT") != 0; ctx->ft_chunk_bits = getenv("HB_FT_CHUNK_BITS") ? atoi(getenv("HB_FT_CHUNK_BITS")) : -1; if (ctx->ft_chunk_bits > 12) ctx->ft_chunk_bits = 12;
	ctx->d_sk_mz = 0; ctx->d_sk_off = 0; ctx->sk_reads = ctx->sk_total = ctx->sk_mz_cap = ctx->sk_off_cap = 0; ctx->sk_reuse = getenv("HB_NO_SKETCH_REUSE") == 0;
	ctx->n_lanes = getenv("HB_LANES") ? atoi(getenv("HB_LANES")) : 3; if (ctx->n_lanes < 1 || ctx->n_lanes > HB_MAX_LANES) ctx->n_lanes = 3; for (int i = 0; i < HB_MAX_LANES - 1; i++) ctx->lane[i] = 0;
	ctx->cns_g_nodes = getenv("HB_CNS_G_NODES") ? (uint32_t)atoi(getenv("HB_CNS_G_NODES")) : 4096; ctx->cns_g_arcs = getenv("HB_CNS_G_ARCS") ? (uint32_t)atoi(getenv("HB_CNS_G_ARCS")) : 32768;
	hb_prof_reset(ctx);
	cudaFuncSetAttribute(k_post_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, POST_WARPS * POST_SMEM_PER_WARP);
	cudaFuncSetAttribute(k_merge_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, MERGE_WARPS * MERGE_SMEM_PER_WARP);
	*out = ctx;
	return HB_OK;
}

static void free_reads(hb_ctx *ctx)
{
	cudaFree(ctx->d_packed); cudaFree(ctx->d_roff); cudaFree(ctx->d_rlen); cudaFree(ctx->d_noff); cudaFree(ctx->d_npos);
	ctx->d_packed = 0; ctx->d_roff = 0; ctx->d_rlen = 0; ctx->d_noff = 0; ctx->d_npos = 0; ctx->n_reads = 0; ctx->packed_cap = ctx->reads_cap = ctx->npos_cap = 0;
}
static void free_prev(hb_ctx *ctx)
{
	cudaFree(ctx->d_prev0); cudaFree(ctx->d_prev1); cudaFree(ctx->d_prev0_off); cudaFree(ctx->d_prev1_off);
	ctx->d_prev0 = ctx->d_prev1 = 0; ctx->d_prev0_off = ctx->d_prev1_off = 0;
}
static void free_out(hb_ctx *ctx)
{
	cudaFree(ctx->d_out0); cudaFree(ctx->d_out1); cudaFree(ctx->d_out0_off); cudaFree(ctx->d_out1_off);
	ctx->d_out0 = ctx->d_out1 = 0; ctx->d_out0_off = ctx->d_out1_off = 0; ctx->n_out0 = ctx->n_out1 = ctx->out_reads = 0; ctx->out0_cap = ctx->out1_cap = ctx->outoff_cap = 0;
}
extern "C" void hb_destroy(hb_ctx_t *ctx)
{
	if (!ctx) return;
	cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream);
	free_reads(ctx); hb_ft_destroy(ctx); hb_pt_destroy(ctx); free_prev(ctx); free_out(ctx); cudaFree(ctx->d_scc); cudaFree(ctx->d_scc_off); cudaFree(ctx->ws); if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
	hb_stage_buf_free(ctx); hb_lane_free(ctx); cudaFree(ctx->d_sk_mz); cudaFree(ctx->d_sk_off);
	cudaStreamDestroy(ctx->stream);
	delete ctx;
}
extern "C" const char *hb_last_error(const hb_ctx_t *ctx) { return ctx ? ctx->err.c_str() : "no context (no CUDA device?)"; }
extern "C" int hb_set_opt(hb_ctx_t *ctx, const hb_opt_t *opt) { int rc = check_opt(ctx, opt); if (rc) return rc; ctx->opt = *opt; return HB_OK; }
extern "C" int hb_get_opt(const hb_ctx_t *ctx, hb_opt_t *opt) { *opt = ctx->opt; return HB_OK; }

// ---------------------------------------------------------------------------
// read store
// ---------------------------------------------------------------------------
// Mirrors the reads into HBM.  The host side only re-lays the packed bytes out on
// 8-byte boundaries into a cached pinned staging buffer (several threads: this is a
// memory copy, not compute) and issues one H2D copy per array.
static int upload_common(hb_ctx *ctx, uint64_t n, const uint64_t *len, const uint8_t *flat, const uint64_t *boff, uint8_t *const *ptrs,
                         const uint64_t *n_pos, const uint64_t *n_off, uint64_t *const *N_site)
{
	cudaSetDevice(ctx->device);
	hb_pt_destroy(ctx); // a new read store: the resident index (if any) describes the old one
	if (n >= (1ull << 28)) { hb_set_err(ctx, HB_E_ARG, "no more than 2^28 reads (htab.cpp:765)"); return HB_E_ARG; }
	std::vector<uint64_t> off(n + 1), noff(n + 1); std::vector<uint32_t> npos;
	ctx->h_rlen.resize(n); ctx->total_bases = 0;
	uint64_t o = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (len[i] >= (1ull << 27)) { hb_set_err(ctx, HB_E_ARG, "read %llu longer than 2^27", (unsigned long long)i); return HB_E_ARG; }
		off[i] = o; ctx->h_rlen[i] = (uint32_t)len[i]; ctx->total_bases += len[i];
		o += ((len[i] / 4 + 1) + 31) & ~31ull; // 32-byte aligned reads: one sector = 128 bases
	}
	off[n] = o;
	if (ctx->h_stage_cap < o + 64) {
		if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
		ctx->h_stage = 0; ctx->h_stage_cap = 0;
		if (cudaMallocHost((void **)&ctx->h_stage, o + 64 + (o >> 4)) != cudaSuccess) { cudaGetLastError(); hb_set_err(ctx, HB_E_NOMEM, "pinned staging buffer"); return HB_E_NOMEM; }
		ctx->h_stage_cap = o + 64 + (o >> 4);
	}
	uint8_t *h_packed = ctx->h_stage;
	{
		unsigned nt = std::thread::hardware_concurrency(); if (nt > 16) nt = 16; if (nt < 1) nt = 1; if (n < 4096) nt = 1;
		auto work = [&](uint64_t a, uint64_t b) {
			for (uint64_t i = a; i < b; i++) {
				const uint8_t *src = ptrs ? ptrs[i] : flat + boff[i];
				uint64_t nb = len[i] / 4 + 1, slot = off[i + 1] - off[i];
				memcpy(h_packed + off[i], src, nb);
				// mask the undefined pad bits so equal sequences have equal bytes
				if (len[i] % 4 == 0) h_packed[off[i] + nb - 1] = 0;
				else h_packed[off[i] + nb - 1] &= (uint8_t)(0xFF << (2 * (4 - len[i] % 4)));
				memset(h_packed + off[i] + nb, 0, slot - nb);
			}
		};
		std::vector<std::thread> th; uint64_t per = (n + nt - 1) / nt;
		for (unsigned t = 1; t < nt; t++) th.emplace_back(work, std::min(n, t * per), std::min(n, (t + 1) * per));
		work(0, std::min(n, per));
		for (auto &t : th) t.join();
		memset(h_packed + o, 0, 16);
	}
	noff[0] = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (N_site) {
			uint64_t c = N_site[i] ? N_site[i][0] : 0;
			for (uint64_t j = 0; j < c; j++) npos.push_back((uint32_t)N_site[i][1 + j]);
			noff[i + 1] = noff[i] + c;
		} else if (n_off) {
			for (uint64_t j = n_off[i]; j < n_off[i + 1]; j++) npos.push_back((uint32_t)n_pos[j]);
			noff[i + 1] = n_off[i + 1];
		} else noff[i + 1] = 0;
	}
	int rc = HB_OK;
	if (ctx->packed_cap < o + 64 || ctx->reads_cap < n + 1 || ctx->npos_cap < npos.size() + 1) {
		free_reads(ctx);
		uint64_t pc = o + 64 + (o >> 4), rc_ = n + 1 + (n >> 4), nc = npos.size() + 1 + (npos.size() >> 2);
		if (cudaMalloc((void **)&ctx->d_packed, pc) != cudaSuccess || cudaMalloc((void **)&ctx->d_roff, (rc_ + 1) * 8) != cudaSuccess ||
		    cudaMalloc((void **)&ctx->d_rlen, (rc_ + 1) * 4) != cudaSuccess || cudaMalloc((void **)&ctx->d_noff, (rc_ + 1) * 8) != cudaSuccess ||
		    cudaMalloc((void **)&ctx->d_npos, nc * 4) != cudaSuccess) { cudaGetLastError(); hb_set_err(ctx, HB_E_NOMEM, "read store does not fit in HBM"); free_reads(ctx); return HB_E_NOMEM; }
		ctx->packed_cap = pc; ctx->reads_cap = rc_; ctx->npos_cap = nc;
	}
	ctx->n_reads = n; ctx->packed_bytes = o; ctx->n_npos = npos.size(); ctx->scc_reads = 0; // staged edit scripts belong to the previous read set
	cudaMemcpyAsync(ctx->d_packed, h_packed, o + 16, cudaMemcpyHostToDevice, ctx->stream); // (+ the zeroed guard bytes)
	cudaMemcpyAsync(ctx->d_roff, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
	cudaMemcpyAsync(ctx->d_rlen, ctx->h_rlen.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream);
	cudaMemcpyAsync(ctx->d_noff, noff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream);
	if (npos.size()) cudaMemcpyAsync(ctx->d_npos, npos.data(), npos.size() * 4, cudaMemcpyHostToDevice, ctx->stream);
	if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) { hb_set_err(ctx, HB_E_CUDA, "read upload failed: %s", cudaGetErrorString(cudaGetLastError())); rc = HB_E_CUDA; }
	if (rc) free_reads(ctx);
	return rc;
}
extern "C" int hb_reads_upload(hb_ctx_t *ctx, uint64_t n, const uint64_t *len, const uint8_t *packed, const uint64_t *byte_off, const uint64_t *n_pos, const uint64_t *n_off)
{ return upload_common(ctx, n, len, packed, byte_off, 0, n_pos, n_off, 0); }
extern "C" int hb_reads_upload_ptrs(hb_ctx_t *ctx, uint64_t n, const uint64_t *len, uint8_t *const *read_sperate, uint64_t *const *N_site)
{ return upload_common(ctx, n, len, 0, 0, read_sperate, 0, 0, N_site); }

// ---------------------------------------------------------------------------
// sketch of a read range -> dense minimizer arrays
// ---------------------------------------------------------------------------
int hb_run_sketch(hb_ctx *ctx, uint64_t r0, uint64_t r1, int rid_mode, DevSketch *out)
{ // two-stage sketch, one chunk of reads at a time so that the scratch (16 B event per base + per-read slices) stays
  // bounded whatever the range.  Chunks run last to first and each one's dense minimizers are stacked downwards on the
  // workspace's hi side, which leaves ONE contiguous array in read order when the first chunk is done.
	(void)rid_mode;
	out->mz = 0; out->off = 0; out->total = 0;
	const uint64_t nR = r1 - r0; Arena ar(ctx);
	SketchPar P = { ctx->opt.mz_win, ctx->opt.k_mer_length, ctx->opt.is_hpc, ctx->opt.mz_sample_dist, ctx->opt.mz_rewin };
	const uint64_t chunk_bases = 1600000000ull; // reads per chunk: 16 bytes of event scratch per base (26 GB at most)
	std::vector<uint64_t> cb(1, 0);
	for (uint64_t c0 = 0; c0 < nR;) {
		uint64_t c1 = c0, b = 0;
		while (c1 < nR && (c1 == c0 || b + ctx->h_rlen[r0 + c1] + 1 <= chunk_bases)) { b += ctx->h_rlen[r0 + c1] + 1; c1++; }
		cb.push_back(c1); c0 = c1;
	}
	uint32_t *d_n = ar.get<uint32_t>(nR + 1); int *d_err = ar.get<int>(1); uint64_t *d_off = ar.get<uint64_t>(nR + 2);
	HB_ALLOC_CHECK(ar);
	const size_t hi_mark = ctx->ws_hi & ~(size_t)255;
	for (int attempt = 0, div = 12; attempt < 3; attempt++, div = div > 4 ? div / 3 : 1) {
		ctx->ws_hi = hi_mark;
		HB_CUDA(cudaMemsetAsync(d_n, 0, (nR + 1) * 4, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_err, 0, 4, ctx->stream));
		hb_mz_t *d_dense = (hb_mz_t *)hb_ws_hi_packed(ctx, 2 * sizeof(hb_mz_t)); // two mapped pad entries past the end
		if (!d_dense) return HB_E_WS;
		uint64_t total = 0; bool overflow = false;
		for (size_t ci = cb.size() - 1; ci-- > 0 && !overflow;) {
			const uint64_t c0 = cb[ci], nc = cb[ci + 1] - c0; Arena ca(ctx);
			std::vector<uint64_t> cap_off(nc + 1), ev_off(nc + 1); uint64_t tot = 0, et = 0;
			for (uint64_t i = 0; i < nc; i++) { const uint64_t l = ctx->h_rlen[r0 + c0 + i]; cap_off[i] = tot; tot += l / div + 32; ev_off[i] = et; et += l + 1; }
			cap_off[nc] = tot; ev_off[nc] = et;
			uint64_t *d_cap = ca.get<uint64_t>(nc + 1), *d_evoff = ca.get<uint64_t>(nc + 1), *d_loff = ca.get<uint64_t>(nc + 2);
			hb_mz_t *d_mz = ca.get<hb_mz_t>(tot); uint32_t *d_l = ca.get<uint32_t>(tot);
			ulonglong2 *d_evs = ca.get<ulonglong2>(et + 1); uint32_t *d_nev = ca.get<uint32_t>(nc + 1), *d_tl = ca.get<uint32_t>(nc + 1);
			if (ca.failed) return HB_E_WS;
			HB_CUDA(cudaMemcpyAsync(d_cap, cap_off.data(), (nc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
			HB_CUDA(cudaMemcpyAsync(d_evoff, ev_off.data(), (nc + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
			{
				ProfScope ps(ctx, "k_sketch_events");
				k_sketch_events<<<nblk(nc, 128), 128, 0, ctx->stream>>>(hb_dev_reads(ctx), hb_dev_ft(ctx), P, r0 + c0, nc, d_evoff, d_evs, d_nev, d_tl);
			}
			{
				ProfScope ps(ctx, "k_sketch_select");
				const int tile = (SK2_TS / P.w) * P.w; size_t smem = (size_t)(tile + P.w) * 40;
				cudaFuncSetAttribute(k_sketch_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
				k_sketch_select<<<(unsigned)nc, SK2_THREADS, smem, ctx->stream>>>(hb_dev_reads(ctx), hb_dev_ft(ctx), P, r0 + c0, nc, d_evoff, d_evs, d_nev, d_tl, d_cap, d_mz, d_l, d_n + c0, d_err);
			}
			HB_CUDA(cudaGetLastError());
			int rc = hb_scan_u32_to_u64(ctx, d_n + c0, d_loff, nc); if (rc) return rc;
			struct { uint64_t tot; int err; } h = { 0, 0 };
			HB_CUDA(cudaMemcpyAsync(&h.tot, d_loff + nc, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&h.err, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
			HB_CUDA(cudaStreamSynchronize(ctx->stream));
			if (h.err) { overflow = true; break; } // a slice overflowed: start over with larger slices
			d_dense = (hb_mz_t *)hb_ws_hi_packed(ctx, h.tot * sizeof(hb_mz_t));
			if (!d_dense) return HB_E_WS;
			{
				ProfScope ps(ctx, "k_compact_mz");
				if (nc) k_compact_mz<<<nblk(nc * 32, 256), 256, 0, ctx->stream>>>(nc, d_cap, d_loff, d_mz, d_dense);
			}
			HB_CUDA(cudaGetLastError());
			total += h.tot;
			if (ci) HB_CUDA(cudaStreamSynchronize(ctx->stream)); // the chunk's scratch is reused by the next one
		}
		if (overflow) continue;
		int rc = hb_scan_u32_to_u64(ctx, d_n, d_off, nR); if (rc) return rc;
		uint64_t *d_off_keep = ar.hi<uint64_t>(nR + 2); HB_ALLOC_CHECK(ar);
		HB_CUDA(cudaMemcpyAsync(d_off_keep, d_off, (nR + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
		out->mz = d_dense; out->off = d_off_keep; out->total = total;
		return HB_OK;
	}
	hb_set_err(ctx, HB_E_OVERFLOW, "minimizer slices overflowed even at one slot per base");
	return HB_E_OVERFLOW;
}

int hb_run_sketch_retry(hb_ctx *ctx, uint64_t r0, uint64_t r1, DevSketch *out)
{ // for callers outside a pass: leaves the outputs on the hi side of a freshly reset workspace
	int rc = HB_OK;
	for (int attempt = 0; attempt < 12; attempt++) {
		if (!ctx->ws && (rc = hb_ws_grow(ctx))) return rc;
		hb_ws_reset(ctx);
		rc = hb_run_sketch(ctx, r0, r1, 0, out);
		if (rc != HB_E_WS) return rc;
		if ((rc = hb_ws_grow(ctx))) return rc;
	}
	hb_set_err(ctx, HB_E_NOMEM, "workspace kept overflowing"); return HB_E_NOMEM;
}

// ---------------------------------------------------------------------------
// the overlap pass over reads [r0,r1)
// ---------------------------------------------------------------------------
struct PassOut { // what the stage APIs want back (device pointers, stolen from the pass arena)
	int want_stage;            // 0 = final pass, 1 = sketch only, 2 = anchors, 3 = chains
	// stage 2: grouped+ordered anchors for the whole range
	hb_hit_t *hits; uint64_t *a_off; uint64_t n_hits;
};

static void weight_table(uint32_t *w_tab, uint32_t high_occ, uint32_t low_occ)
{ // weight of an anchor from its minimizer's occurrence count (anchor.cpp:991-999,
  // 1066-1075); tabulated once on the host (libm pow), 4096 entries (n <= 4094)
	uint64_t max_cnt = high_occ < 2 ? 2 : high_occ, min_cnt = low_occ < 2 ? 2 : low_occ;
	for (uint32_t n = 0; n < 4096; n++) {
		uint32_t w;
		if (n < max_cnt && n > min_cnt) w = 1;
		else if (n <= min_cnt) w = 2;
		else { w = (uint32_t)(1 + ((n + (max_cnt << 1) - 1) / (max_cnt << 1))); w = (uint32_t)pow((double)w, 1.1); }
		if (w > 0xffffffu) w = 0xffffffu;
		w_tab[n] = w;
	}
}

static ChainPar chain_par(const hb_ctx *ctx, double bw)
{ // set_lchain_dp_op(is_accurate=1), anchor.cpp:2272-2285; arguments of ecovlp.cpp:3957
	ChainPar P; double tmp = expf((float)(-0.01 * (double)ctx->opt.k_mer_length));
	P.pen_gap = 0.5f; P.pen_skip = 0.0005f; P.pen_gap *= tmp; P.pen_skip *= tmp; P.bw_rate = bw;
	P.max_skip = 25; P.max_iter = 5000; P.max_dis = 5000; P.mcopy_num = 3; P.mcopy_khit_cutoff = 32; P.mcopy_rate = 0.7;
	P.max_n_chain = ctx->opt.max_n_chain; P.chain_cutoff = 2; P.ocv_w = 3072;
	return P;
}

__global__ void k_cap_per_read(uint64_t nR, const uint32_t *__restrict__ sc, const uint64_t *__restrict__ in0_off, uint32_t *__restrict__ cap)
{ uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r < nR) cap[r] = sc[r] + (uint32_t)(in0_off[r + 1] - in0_off[r]); }
__global__ void k_cc_per_read(uint64_t nR, uint64_t r0, const uint32_t *__restrict__ rlen, uint32_t ocv_w, uint32_t *__restrict__ cap)
{ uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r < nR) cap[r] = rlen[r0 + r] / ocv_w + 2; }
__global__ void k_rebase(uint64_t n, const uint64_t *__restrict__ in, uint64_t base, uint64_t *__restrict__ out)
{ uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = in[i] - base; }
__global__ void k_add_off(uint64_t n, uint64_t *__restrict__ a, uint64_t add)
{ uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] += add; }
// stage-3 output assembly: per read, chains in final order with compacted anchor index, chain anchors, fake cigars
__global__ void k_stage3_counts(uint64_t nR, const uint64_t *__restrict__ c_off, const hb_chain_t *__restrict__ ch, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ n_ol,
                                uint32_t *__restrict__ n_hit, uint32_t *__restrict__ n_fc)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint64_t cb = c_off[r]; uint32_t ns = (uint32_t)(c_off[r + 1] - cb), h = 0, f = 0;
	for (uint32_t s = 0; s < ns; s++) h += ch[cb + s].n_hits;
	for (uint32_t i = 0; i < n_ol[r]; i++) f += ch[cb + idx[cb + i]].fc_n;
	n_hit[r] = h; n_fc[r] = f;
}
__global__ void k_stage3_fill(uint64_t nR, const uint64_t *__restrict__ c_off, const uint64_t *__restrict__ a_off, uint64_t a_base, const hb_chain_t *__restrict__ ch, const GroupDir *dirless,
                              const uint32_t *__restrict__ idx, const uint32_t *__restrict__ n_ol, const hb_hit_t *__restrict__ chits, const hb_hit_t *__restrict__ ghits, const uint64_t *__restrict__ fc,
                              const uint64_t *__restrict__ o_ch, const uint64_t *__restrict__ o_hit, const uint64_t *__restrict__ o_fc,
                              hb_chain_t *__restrict__ out_ch, hb_hit_t *__restrict__ out_hit, uint64_t *__restrict__ out_fc, const uint64_t *__restrict__ fc_grp_base)
{
	(void)dirless; (void)a_off; (void)a_base;
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (r >= nR) return;
	uint64_t cb = c_off[r]; uint32_t ns = (uint32_t)(c_off[r + 1] - cb), ord = 0; uint64_t m = o_hit[r], fo = o_fc[r];
	for (uint32_t s = 0; s < ns; s++) { // compacted chain anchors, tagged with the chain ordinal (Hash_Table.cpp:2240,2279)
		const hb_chain_t &c = ch[cb + s];
		if (!c.n_hits) continue;
		const hb_hit_t *hb_ = (c.pad & HB_CHAIN_INPLACE) ? ghits : chits;
		for (uint32_t h = 0; h < c.n_hits; h++) { hb_hit_t v = hb_[c.first_hit + h]; v.id_strand = (v.id_strand & 0x80000000u) | ord; out_hit[m++] = v; }
		ord++;
	}
	for (uint32_t i = 0; i < n_ol[r]; i++) {
		hb_chain_t c = ch[cb + idx[cb + i]];
		const uint64_t *src = fc + fc_grp_base[cb + idx[cb + i]] + c.fc_off;
		for (uint32_t j = 0; j < c.fc_n; j++) out_fc[fo + j] = src[j];
		c.fc_off = (uint32_t)(fo - o_fc[r]); fo += c.fc_n;
		c.first_hit = c.pad & ~HB_CHAIN_INPLACE; c.pad = 0;
		out_ch[o_ch[r] + i] = c;
	}
}
__global__ void k_fc_grp_base(const GroupDir *__restrict__ dir, const uint32_t *__restrict__ dir_n, const uint64_t *__restrict__ a_off, uint64_t a_base, const uint64_t *__restrict__ c_off,
                              int32_t mcopy_num, int32_t cutoff, uint64_t *__restrict__ base)
{
	uint32_t n = *dir_n;
	for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
		GroupDir d = dir[g]; if (d.slot == GRP_EMPTY) continue;
		uint64_t ab = a_off[d.read] - a_base + d.start, cb = c_off[d.read] + d.slot; int32_t ns = d.count >= (uint32_t)cutoff ? mcopy_num : 1;
		for (int32_t s = 0; s < ns; s++) base[cb + s] = ab + 2 * cb;
	}
}

struct StageOut { // host destinations of the stage APIs (all optional)
	uint64_t *off; void *rec; uint64_t rec_cap;      // stage 1: minimizers / stage 2: anchors / stage 3: chains
	uint64_t *hit_off; hb_hit_t *hits; uint64_t hit_cap; uint64_t *fc_off; uint64_t *fc; uint64_t fc_cap;
	double e_rate; int32_t w_l; int32_t gaps, use_prev; // window pass; step C on/off; row a12 (needs the previous round's overlaps staged)
	hb_wl_t *wl; uint64_t wl_cap; uint16_t *cig; uint64_t cig_cap; uint64_t n_wl, n_cig; // step A of the EC alignment stage (mode 5)
	hb_ma_hit_t *rec2; uint64_t rec2_cap; uint64_t *off2; uint8_t *flags; // mode 9: rec / off = the round's paf[], rec2 / off2 = reverse_paf[], flags = 2 bytes per read
	int cns; std::vector<uint16_t> *h_scc; std::vector<uint64_t> *h_scc_off; uint8_t *status; uint64_t n_corrected; // mode 9 with the consensus on the device (row a14): scripts of the range, per-read status
};

// mode: 0 final pass (results kept in ctx->d_out*), 2 anchors, 3 chains
static int run_pass_impl(hb_ctx *ctx, uint64_t r0, uint64_t r1, int mode, double bw, StageOut *so, uint64_t *stat_out);
static int run_pass(hb_ctx *ctx, uint64_t r0, uint64_t r1, int mode, double bw, StageOut *so, uint64_t *stat_out)
{ // brackets the pass with events on the context stream
	cudaSetDevice(ctx->device);
	cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, ctx->stream);
	int rc = HB_OK;
	for (int attempt = 0; attempt < 12; attempt++) { // a workspace that turns out too small is grown and the pass rerun (first pass only in steady state)
		if (!ctx->ws && (rc = hb_ws_grow(ctx))) break;
		hb_ws_reset(ctx);
		rc = run_pass_impl(ctx, r0, r1, mode, bw, so, stat_out);
		hb_ws_reset(ctx);
		if (rc != HB_E_WS) break;
		{ bool grew = false, fail = false;
			for (int i = 0; i < HB_MAX_LANES - 1 && !fail; i++) {
				hb_ctx *L = ctx->lane[i]; if (!L || !L->ws_need) continue;
				hb_ws_reset(L); grew = true;
				if ((rc = hb_ws_grow(L)) == HB_E_NOMEM) { ctx->n_lanes = i + 1; L->ws_need = 0; } // no room for this lane's workspace: the pass is rerun on fewer lanes
				else if (rc) { hb_set_err(ctx, rc, "%s", L->err.c_str()); fail = true; }
			}
			if (fail) break;
			if (grew && !ctx->ws_need) { rc = HB_E_WS; continue; } } // (only other lanes ran short)
		if ((rc = hb_ws_grow(ctx))) break;
		cudaEventRecord(a, ctx->stream); // time the attempt that succeeds
		rc = HB_E_WS;
	}
	if (rc == HB_E_WS) { hb_set_err(ctx, HB_E_NOMEM, "workspace kept overflowing"); rc = HB_E_NOMEM; }
	float ms = 0; cudaEventRecord(b, ctx->stream); cudaEventSynchronize(b); cudaEventElapsedTime(&ms, a, b);
	cudaEventDestroy(a); cudaEventDestroy(b);
	ctx->last_pass_ms = ms;
	return rc;
}
#include <chrono>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>
#include <functional>

// ---- lanes: the batches of a pass on two streams (see run_pass_impl) ----
// What a batch adds to the host-side results of the pass is added in batch order: commit(k, fn) runs fn when the batches before k have committed.
#include "hb_order.h"
// the second lane's context: a copy of the pass's context (read store, index, staged lists, options: shared, read-only inside a pass) with a stream, workspace,
// profile and counters of its own.  Owns only its stream and workspace (hb_lane_free).
static int lane_refresh(hb_ctx *ctx, int li)
{
	hb_ctx *L = ctx->lane[li]; cudaStream_t st = 0; uint8_t *ws = 0; size_t ws_cap = 0, ws_need = 0;
	if (L) { st = L->stream; ws = L->ws; ws_cap = L->ws_cap; ws_need = L->ws_need; }
	else { L = new hb_ctx(); if (cudaStreamCreate(&st) != cudaSuccess) { delete L; hb_set_err(ctx, HB_E_CUDA, "stream of lane %d", li + 1); return HB_E_CUDA; } }
	*L = *ctx;
	L->stream = st; L->ws = ws; L->ws_cap = ws_cap; L->ws_need = ws_need; for (int i = 0; i < HB_MAX_LANES - 1; i++) L->lane[i] = 0; L->n_lanes = 1; L->stage_buf = 0; L->h_stage = 0; L->h_stage_cap = 0;
	L->prof.clear(); L->prof_stage.clear(); L->in_stage = 0; memset(L->counters, 0, sizeof(L->counters)); memset(L->stage_counters, 0, sizeof(L->stage_counters)); L->err.clear();
	ctx->lane[li] = L;
	if (!L->ws) { L->ws_need = std::max<size_t>(L->ws_need, std::max<size_t>(ctx->ws_cap / 2, (size_t)256 << 20)); /* a first guess: half of what the pass's own lane has grown to */ int rc = hb_ws_grow(L); if (rc) { hb_set_err(ctx, rc, "%s", L->err.c_str()); return rc; } }
	hb_ws_reset(L);
	return HB_OK;
}
void hb_lane_free(hb_ctx *ctx)
{
	for (int i = 0; i < HB_MAX_LANES - 1; i++) {
		hb_ctx *L = ctx->lane[i]; if (!L) continue;
		cudaStreamSynchronize(L->stream); cudaFree(L->ws); cudaStreamDestroy(L->stream);
		delete L; ctx->lane[i] = 0;
	}
}
void hb_ws_release(hb_ctx *ctx)
{ // between passes: give the workspaces back (an index build needs the room); the lanes keep their streams
	cudaStreamSynchronize(ctx->stream);
	size_t keep = ctx->ws_cap; if (ctx->ws) cudaFree(ctx->ws); ctx->ws = 0; ctx->ws_cap = 0; ctx->ws_need = keep - (keep >> 3); hb_ws_reset(ctx); // (regrown to about the old size in one step)
	for (int i = 0; i < HB_MAX_LANES - 1; i++) { hb_ctx *L = ctx->lane[i]; if (!L || !L->ws) continue; cudaStreamSynchronize(L->stream); keep = L->ws_cap; cudaFree(L->ws); L->ws = 0; L->ws_cap = 0; L->ws_need = keep - (keep >> 3); }
}
static void lane_merge(hb_ctx *ctx, hb_ctx *L)
{ // the lane's kernel times and counters into the pass's
	for (auto &q : L->prof) {
		bool f = false; for (auto &p : ctx->prof) if (p.name == q.name) { p.launches += q.launches; p.ms += q.ms; f = true; break; } if (!f) ctx->prof.push_back(q);
		if (ctx->in_stage) { f = false; for (auto &p : ctx->prof_stage) if (p.name == q.name) { p.launches += q.launches; p.ms += q.ms; f = true; break; } if (!f) ctx->prof_stage.push_back(q); }
	}
	for (int i = 0; i < 12; i++) ctx->counters[i] += L->counters[i];
}
typedef std::function<int(hb_ctx *, size_t)> BatchFn;
static int run_batches(hb_ctx *ctx, int mode, size_t n, PassOrder &order, const BatchFn &body)
{
	// the stage views (modes 2..8) place their output at running offsets all through the batch: one lane
	const int lanes = (mode == 0 || mode == 9) && !ctx->trace && n > 1 ? ctx->n_lanes : 1;
	auto one = [&](hb_ctx *c, size_t k) -> int {
		int rc = body(c, k);
		if (!rc) { bool d; { std::lock_guard<std::mutex> lk(order.mu); d = order.done[k] != 0; } if (!d) rc = order.commit(k, []() { return HB_OK; }); }
		if (rc) order.abort();
		return rc;
	};
	if (lanes < 2) { for (size_t k = 0; k < n; k++) { const int rc = one(ctx, k); if (rc) return rc; } return HB_OK; }
	int extra = (int)std::min<size_t>((size_t)lanes, n) - 1; int rc;
	for (int i = 0; i < extra; i++) if ((rc = lane_refresh(ctx, i))) { if (rc != HB_E_NOMEM) return rc; extra = i; break; } // no room for another workspace: fewer lanes
	std::atomic<size_t> nextk(0); int rcs[HB_MAX_LANES]; for (int i = 0; i < HB_MAX_LANES; i++) rcs[i] = HB_OK;
	auto worker = [&](hb_ctx *c, int *out) {
		cudaSetDevice(c->device);
		for (;;) { const size_t k = nextk.fetch_add(1); if (k >= n) break; { std::lock_guard<std::mutex> lk(order.mu); if (order.aborted) break; } const int r = one(c, k); if (r) { *out = r; break; } }
		cudaStreamSynchronize(c->stream);
	};
	std::vector<std::thread> th;
	for (int i = 0; i < extra; i++) th.emplace_back(worker, ctx->lane[i], &rcs[i + 1]);
	worker(ctx, &rcs[0]);
	for (auto &t : th) t.join();
	for (int i = 0; i < extra; i++) lane_merge(ctx, ctx->lane[i]);
	// which error to report: a workspace that was too small on any lane first (the pass is rerun after growing it), then the lane that failed by itself
	for (int i = 0; i <= extra; i++) if (rcs[i] == HB_E_WS) return HB_E_WS;
	if (rcs[0] && rcs[0] != HB_E_ABORTED) return rcs[0];
	for (int i = 1; i <= extra; i++) if (rcs[i] && rcs[i] != HB_E_ABORTED) { hb_set_err(ctx, rcs[i], "%s", ctx->lane[i - 1]->err.c_str()); return rcs[i]; }
	for (int i = 0; i <= extra; i++) if (rcs[i]) { hb_set_err(ctx, HB_E_STATE, "a batch was abandoned without an error"); return HB_E_STATE; }
	return HB_OK;
}

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
#define TRACE(label) do { if (trace) { cudaStreamSynchronize(ctx->stream); double t_ = now_ms(); fprintf(stderr, "[hb trace] %-18s +%8.2f ms (total %8.2f)\n", label, t_ - t_last, t_ - t_begin); t_last = now_ms(); } } while (0)
static int run_pass_impl(hb_ctx *ctx, uint64_t r0, uint64_t r1, int mode, double bw, StageOut *so, uint64_t *stat_out)
{
	const bool trace = ctx->trace != 0; double t_begin = now_ms(), t_last = t_begin;
	cudaSetDevice(ctx->device);
	if (r1 > ctx->n_reads || r0 > r1) { hb_set_err(ctx, HB_E_ARG, "read range out of bounds"); return HB_E_ARG; }
	if (!ctx->d_pt_slot) { hb_set_err(ctx, HB_E_STATE, "no position index: call hb_pt_gen first"); return HB_E_STATE; }
	if (mode == 0 && !ctx->d_prev0_off) { hb_set_err(ctx, HB_E_STATE, "previous overlaps are not staged"); return HB_E_STATE; }
	const uint64_t nR = r1 - r0; Arena ar(ctx); int rc;
	DevReads R = hb_dev_reads(ctx); DevPt PT = hb_dev_pt(ctx);
	const uint32_t high_occ = (uint32_t)(ctx->opt.hom_cov * (2.0 - 0.333)), low_occ = (uint32_t)(ctx->opt.hom_cov * 0.333); // ecovlp.cpp:3952-3953
	uint32_t h_wtab[4096]; weight_table(h_wtab, high_occ, low_occ);
	ChainPar CP = chain_par(ctx, bw);
	hb_prof_reset(ctx);
	if (mode == 9 && so->cns) { so->h_scc->clear(); so->h_scc_off->assign(nR + 1, 0); so->n_corrected = 0; } // (a rerun after a workspace growth starts over)
	if (nR == 0) { if (so && so->off) so->off[0] = 0; return HB_OK; }

	// ---- sketch + probe for the whole range
	DevSketch sk;
	if (ctx->sk_reads && ctx->sk_reads == ctx->n_reads) { // hb_pt_gen sketched these very reads with this filter table when it built the index: the pass reads that sketch
		const uint64_t base = ctx->h_sk_off[r0]; uint64_t *d_o = ar.get<uint64_t>(nR + 2); HB_ALLOC_CHECK(ar);
		k_rebase<<<nblk(nR + 1, 256), 256, 0, ctx->stream>>>(nR + 1, ctx->d_sk_off + r0, base, d_o);
		sk.mz = ctx->d_sk_mz + base; sk.off = d_o; sk.total = ctx->h_sk_off[r1] - base;
	} else { rc = hb_run_sketch(ctx, r0, r1, 0, &sk); if (rc) return rc; }
	TRACE("sketch");
	uint64_t *d_seeds = ar.get<uint64_t>(sk.total + 1); uint32_t *d_spre = ar.get<uint32_t>(sk.total + 1), *d_acnt = ar.zero<uint32_t>(nR + 1), *d_wtab = ar.get<uint32_t>(4096);
	uint64_t *d_aoff = ar.get<uint64_t>(nR + 2);
	HB_ALLOC_CHECK(ar);
	HB_CUDA(cudaMemcpyAsync(d_wtab, h_wtab, sizeof(h_wtab), cudaMemcpyHostToDevice, ctx->stream));
	{
		ProfScope ps(ctx, "k_probe_count");
		k_probe_count<<<nblk(nR * 32, 256), 256, 0, ctx->stream>>>(PT, nR, sk.off, sk.mz, d_seeds, d_spre, d_acnt);
	}
	HB_CUDA(cudaGetLastError());
	rc = hb_scan_u32_to_u64(ctx, d_acnt, d_aoff, nR); if (rc) return rc;
	std::vector<uint64_t> h_aoff(nR + 1), h_mzoff(nR + 1);
	HB_CUDA(cudaMemcpyAsync(h_aoff.data(), d_aoff, (nR + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(h_mzoff.data(), sk.off, (nR + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	TRACE("probe+scan");
	ctx->counters[0] = nR; ctx->counters[2] = sk.total; ctx->counters[3] = h_aoff[nR];
	for (uint64_t i = r0; i < r1; i++) ctx->counters[1] += ctx->h_rlen[i];

	// ---- whole-range outputs
	hb_hit_t *d_all_hits = 0; // mode 2
	if (mode == 2) { d_all_hits = ar.get<hb_hit_t>(h_aoff[nR] + 1); HB_ALLOC_CHECK(ar); }
	std::vector<uint64_t> st3_off, st3_hit_off, st3_fc_off; // mode 3 host offsets (accumulated per batch)
	uint64_t st3_n = 0, st3_nh = 0, st3_nf = 0;
	if (mode >= 3) { st3_off.assign(nR + 1, 0); st3_hit_off.assign(nR + 1, 0); st3_fc_off.assign(nR + 1, 0); }
	unsigned long long *d_stat = ar.zero<unsigned long long>(16);
	uint32_t *d_m0 = 0, *d_m1 = 0; // per-read result counts (final pass)
	struct BatchRes { hb_ma_hit_t *o0, *o1; uint64_t *ooff; uint64_t b0, b1; };
	std::vector<BatchRes> bres;
	if (mode == 0) { d_m0 = ar.zero<uint32_t>(nR + 1); d_m1 = ar.zero<uint32_t>(nR + 1); }
	HB_ALLOC_CHECK(ar);

	// ---- batches bounded by the anchor budget.  The batches are independent: with more than one lane (HB_LANES, default 3, at most 4) as many host threads take them in turn, each with
	// its own stream and workspace (a clone of the context, lane_refresh), so that the launches of one batch whose time is set by their heaviest read (phasing, chain
	// post-processing, the merge of step B: one wave of warps, most SMs idle behind the last one) share the device with the throughput-bound launches of the other.
	// What a batch adds to the pass's host-side results is added in batch order (PassOrder), so the results do not depend on which lane ran what.
	std::vector<std::pair<uint64_t, uint64_t>> batches;
	for (uint64_t b0 = 0; b0 < nR;) {
		uint64_t b1 = b0 + 1;
		while (b1 < nR && h_aoff[b1 + 1] - h_aoff[b0] <= ctx->anchor_budget) b1++;
		batches.push_back(std::make_pair(b0, b1)); b0 = b1;
	}
	PassOrder order(batches.size());
	hb_ctx *const pass_ctx = ctx;
	auto batch_body = [&](hb_ctx *ctx, const size_t bk) -> int { // `ctx` = the lane's context: stream, workspace, profile and counters of its own; everything else is shared and read-only
		const uint64_t b0 = batches[bk].first, b1 = batches[bk].second; int rc;
		const uint64_t nb = b1 - b0, B = h_aoff[b1] - h_aoff[b0], a_base = h_aoff[b0], mz_b = h_mzoff[b0], n_mz = h_mzoff[b1] - h_mzoff[b0];
		if (B >= (1ull << 32)) { hb_set_err(ctx, HB_E_OVERFLOW, "a single read produces >= 2^32 anchors"); return HB_E_OVERFLOW; }
		Arena ba(ctx);
		hb_hit_t *d_raw = ba.get<hb_hit_t>(B + 1), *d_hits = mode == 2 ? d_all_hits + a_base : ba.get<hb_hit_t>(B + 1);
		uint32_t *d_sc = ba.zero<uint32_t>(nb + 1), *d_dirn = ba.zero<uint32_t>(1); int *d_err = ba.zero<int>(1);
		uint64_t *d_coff = ba.get<uint64_t>(nb + 2); unsigned long long *d_arena_used = ba.zero<unsigned long long>(1);
		HB_ALLOC_CHECK(ba);
		if (B) {
			ProfScope ps(ctx, "k_expand");
			const int U = 2; // query minimizers in flight per half-warp: full occupancy at 40 registers (0.62 of the measured HBM peak; 1 / 4 / 8 in flight were slower)
			unsigned grid = (unsigned)std::min<uint64_t>((n_mz * 16 / U + 255) / 256, (uint64_t)ctx->sm_count * 32);
			if (!grid) grid = 1;
			k_expand<2, 6><<<grid, 256, 0, ctx->stream>>>(R, PT, r0, n_mz, sk.mz + mz_b, d_seeds + mz_b, d_spre + mz_b, d_aoff, a_base, d_wtab, d_raw);
		}
		HB_CUDA(cudaGetLastError());
	TRACE("expand");
		// group: one stable device-wide sort of the batch's anchors on (read, target, strand) + run-length encoding (group.cu)
		GroupDir *d_dir = 0; uint32_t h_dirn = 0;
		if ((rc = hb_group_sort(ctx, nb, r0 + b0, d_aoff + b0, a_base, B, d_raw, d_hits, &d_dir, d_dirn, &h_dirn, d_sc, CP.mcopy_num, CP.mcopy_khit_cutoff))) return rc;
	TRACE("group");
		ba.release(d_raw);
		ctx->counters[4] += h_dirn;
		rc = hb_scan_u32_to_u64(ctx, d_sc, d_coff, nb); if (rc) return rc;
		uint64_t n_slots = 0; HB_CUDA(cudaMemcpyAsync(&n_slots, d_coff + nb, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));

		// chain
		hb_hit_t *d_chits = ba.get<hb_hit_t>(B + 1); int32_t *d_f = ba.get<int32_t>(B + 1), *d_p = ba.get<int32_t>(B + 1), *d_ii = ba.get<int32_t>(B + 1); int64_t *d_t = ba.get<int64_t>(B + 1);
		hb_chain_t *d_ch = ba.zero<hb_chain_t>(n_slots + 1); uint32_t *d_slot_read = ba.get<uint32_t>(n_slots + 1);
		uint64_t *d_fc = mode >= 3 ? ba.get<uint64_t>(B + 2 * n_slots + 4) : 0;
		HB_ALLOC_CHECK(ba);
		{
			ChainArgs C; C.R = R; C.r0 = r0 + b0; C.dir = d_dir; C.dir_n = d_dirn; C.a_off = d_aoff + b0; C.a_base = a_base; C.c_off = d_coff;
			C.hits = d_hits; C.chits = d_chits; C.f = d_f; C.p = d_p; C.ii = d_ii; C.t = d_t; C.ch = d_ch; C.slot_read = d_slot_read; C.fc = d_fc; C.P = CP; C.err = d_err; C.dbg = d_stat + 8; C.work = ba.zero<uint32_t>(1); HB_ALLOC_CHECK(ba);
			ProfScope ps(ctx, "k_chain");
			k_chain_warp<<<std::max(1u, std::min(nblk((uint64_t)h_dirn * 32, 128), (unsigned)ctx->sm_count * 16)), 128, 0, ctx->stream>>>(C);
		}
		HB_CUDA(cudaGetLastError());
	TRACE("chain");
		if (mode == 2) return HB_OK;
		ba.release(d_f); ba.release(d_p); ba.release(d_ii); ba.release(d_t);

		// post
		uint32_t *d_idx = ba.get<uint32_t>(n_slots + 1), *d_nol = ba.zero<uint32_t>(nb + 1), *d_cccap = ba.get<uint32_t>(nb + 1); uint8_t *d_keep = ba.zero<uint8_t>(n_slots + 1), *d_exact = ba.zero<uint8_t>(n_slots + 1);
		uint64_t *d_ccoff = ba.get<uint64_t>(nb + 2);
		HB_ALLOC_CHECK(ba);
		k_cc_per_read<<<nblk(nb, 256), 256, 0, ctx->stream>>>(nb, r0 + b0, ctx->d_rlen, CP.ocv_w, d_cccap);
		rc = hb_scan_u32_to_u64(ctx, d_cccap, d_ccoff, nb - 0); if (rc) return rc;
		uint64_t cc_tot = 0; for (uint64_t i = r0 + b0; i < r0 + b1; i++) cc_tot += ctx->h_rlen[i] / CP.ocv_w + 2;
		uint64_t *d_cc = ba.get<uint64_t>(cc_tot + 1); HB_ALLOC_CHECK(ba);
		{
			PostArgs Pa; Pa.R = R; Pa.r0 = r0 + b0; Pa.nR = nb; Pa.c_off = d_coff; Pa.ch = d_ch; Pa.chits = d_chits; Pa.ghits = d_hits; Pa.idx = d_idx; Pa.n_ol = d_nol; Pa.keep = d_keep; Pa.cc = d_cc; Pa.cc_off = d_ccoff; Pa.P = CP;
			ProfScope ps(ctx, "k_post");
			k_post_warp<<<nblk(nb, POST_WARPS), POST_WARPS * 32, POST_WARPS * POST_SMEM_PER_WARP, ctx->stream>>>(Pa);
		}
		HB_CUDA(cudaGetLastError());
	TRACE("post");
		ctx->counters[5] += n_slots;

		if (mode >= 4) { // window pass of an EC round over the chains of this batch
			uint32_t *d_wc = ba.zero<uint32_t>(nb + 1); uint64_t *d_woff = ba.get<uint64_t>(nb + 2), *d_fcb = ba.get<uint64_t>(n_slots + 1);
			HB_ALLOC_CHECK(ba);
			const int32_t w_l = so->w_l;
			k_win_count<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_coff, d_ch, d_idx, d_nol, w_l, d_wc);
			k_fc_grp_base<<<std::max(1u, std::min(nblk(h_dirn, 128), (unsigned)ctx->sm_count * 16)), 128, 0, ctx->stream>>>(d_dir, d_dirn, d_aoff + b0, a_base, d_coff, CP.mcopy_num, CP.mcopy_khit_cutoff, d_fcb);
			if ((rc = hb_scan_u32_to_u64(ctx, d_wc, d_woff, nb))) return rc;
			std::vector<uint64_t> h_woff(nb + 1);
			HB_CUDA(cudaMemcpyAsync(h_woff.data(), d_woff, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
			const uint64_t n_win = h_woff[nb];
			WinDesc *d_desc = ba.get<WinDesc>(n_win + 1); hb_win_t *d_wout = ba.get<hb_win_t>(n_win + 1);
			HB_ALLOC_CHECK(ba);
			k_win_desc<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_coff, d_ch, d_idx, d_nol, w_l, d_woff, d_desc);
			// row a12 first (as gen_hc_r_alin_ea does, ecovlp.cpp:2810): an overlap the previous round's exact record still covers needs no window alignment at all —
			// in rounds 2 and 3 that is most of them — so the window pass skips its windows
			uint64_t *d_ooff = 0; OvDesc *d_od = 0; uint8_t *d_ea = 0; uint64_t n_ov = 0; std::vector<uint64_t> h_ooff;
			if (mode >= 5) {
				d_ooff = ba.get<uint64_t>(nb + 2); HB_ALLOC_CHECK(ba);
				if ((rc = hb_scan_u32_to_u64(ctx, d_nol, d_ooff, nb))) return rc;
				h_ooff.resize(nb + 1);
				HB_CUDA(cudaMemcpyAsync(h_ooff.data(), d_ooff, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
				n_ov = h_ooff[nb];
				d_od = ba.get<OvDesc>(n_ov + 1); HB_ALLOC_CHECK(ba);
				k_ov_desc<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_coff, d_ch, d_idx, d_nol, w_l, d_woff, d_ooff, d_od);
				if (so->use_prev) {
					if (!ctx->d_prev0_off) { hb_set_err(ctx, HB_E_STATE, "previous overlaps are not staged (hb_ec_stage_prev)"); return HB_E_STATE; }
					d_ea = ba.zero<uint8_t>(n_ov + 1); HB_ALLOC_CHECK(ba);
					ProfScope ps(ctx, "k_ec_ea");
					if (n_ov) k_ec_ea_w<<<nblk(n_ov * 32, 128), 128, 0, ctx->stream>>>(R, r0 + b0, n_ov, d_od, d_ch, ctx->d_prev0, ctx->d_prev0_off, d_ea);
				}
			}
			{
				WinArgs W; W.R = R; W.r0 = r0 + b0; W.n_win = n_win; W.desc = d_desc; W.ch = d_ch; W.fc = d_fc; W.fc_grp_base = d_fcb; W.e_rate = so->e_rate; W.w_l = w_l; W.out = d_wout; W.err = d_err; W.ea = d_ea; W.o_off = d_ooff;
				ProfScope ps(ctx, "k_windows");
				if (n_win) k_windows<<<nblk(n_win, 128), 128, 0, ctx->stream>>>(W);
			}
			HB_CUDA(cudaGetLastError());
			ctx->counters[8] += n_win;
			int h_err2 = 0; HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
			if (h_err2 & 32) { hb_set_err(ctx, HB_E_STATE, "a window start fell outside its chain's fake cigar"); return HB_E_STATE; }
			if (mode >= 5) { // step A of the alignment stage: one thread per overlap consumes the window records
				// the queued (aligner) pass of step A: its grid is bounded by the per-thread trace scratch (31 KB); up to 8 GB of it, so that in the usual case every queued overlap has a thread
				const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_ov + 63) / 64, std::max<uint64_t>((uint64_t)ctx->sm_count * 4, (8ull << 30) / ((uint64_t)w_l * 40 + 2 * HB_EC_CIG_TMP) / 64))); const uint64_t nthr = (uint64_t)blocks * 64;
				hb_aln_t *d_aln = ba.get<hb_aln_t>(n_ov + 1); hb_wl_t *d_wl = ba.zero<hb_wl_t>(n_win + 1);
				uint64_t *d_path = ba.get<uint64_t>(nthr * (uint64_t)w_l * 5); uint16_t *d_ctmp = ba.get<uint16_t>(nthr * HB_EC_CIG_TMP);
				unsigned long long *d_pused = ba.zero<unsigned long long>(1);
				HB_ALLOC_CHECK(ba);
				uint64_t pool_cap = n_win * 4 + 4096, pool_used = 0; uint16_t *d_pool = 0;
				for (int attempt = 0;; attempt++) {
					d_pool = ba.get<uint16_t>(pool_cap); HB_ALLOC_CHECK(ba);
					HB_CUDA(cudaMemsetAsync(d_pused, 0, 8, ctx->stream));
					EcAlnArgs E; E.ea = d_ea; E.R = R; E.r0 = r0 + b0; E.n_ov = n_ov; E.desc = d_od; E.ch = d_ch; E.fc = d_fc; E.fc_grp_base = d_fcb; E.win = d_wout; E.e_rate = so->e_rate; E.w_l = w_l;
					E.wl = d_wl; E.out = d_aln; E.path = d_path; E.cig_tmp = d_ctmp; E.pool = d_pool; E.pool_used = d_pused; E.pool_cap = pool_cap; E.err = d_err;
					E.q = ba.get<uint32_t>(n_ov + 1); E.q_n = ba.zero<uint32_t>(1); HB_ALLOC_CHECK(ba);
					if (attempt) HB_CUDA(cudaMemsetAsync(d_wl, 0, (n_win + 1) * sizeof(hb_wl_t), ctx->stream)); // (a rerun with a larger pool starts from clean window lists)
					{
						ProfScope ps(ctx, "k_ec_overlap_fast");
						if (n_ov) hb_k_ecaln(0, nblk(n_ov, 128), ctx->stream, E);
					}
					{
						ProfScope ps(ctx, "k_ec_overlap");
						if (n_ov) hb_k_ecaln(1, blocks, ctx->stream, E);
					}
					HB_CUDA(cudaGetLastError());
					HB_CUDA(cudaMemcpyAsync(&pool_used, d_pused, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
					HB_CUDA(cudaStreamSynchronize(ctx->stream));
					if (h_err2 & (32 | 64)) { hb_set_err(ctx, HB_E_STATE, "EC alignment: fake-cigar lookup or cigar scratch overflow (flags %d)", h_err2); return HB_E_STATE; }
					if (pool_used <= pool_cap) break;
					if (attempt >= 2) { hb_set_err(ctx, HB_E_OVERFLOW, "EC alignment: cigar pool"); return HB_E_OVERFLOW; }
					ba.release(d_pool); pool_cap = pool_used + 4096; // the exact need is known now
				}
				ctx->counters[9] += n_ov;
				if (mode >= 6) { // step B: base-level CIGAR of the accepted overlaps (row a10)
					uint32_t *d_cap = ba.zero<uint32_t>(n_ov + 1); uint64_t *d_wboff = ba.get<uint64_t>(n_ov + 2); hb_alnb_t *d_alnb = ba.get<hb_alnb_t>(n_ov + 1), *d_alnb2 = ba.get<hb_alnb_t>(n_ov + 1);
					unsigned int *d_ndef = (unsigned int *)ba.zero<uint32_t>(1);
					int64_t *d_dpt = ba.get<int64_t>(2 * (B + 1)), *d_dpp = ba.get<int64_t>(2 * (B + 1)); int32_t *d_dpf = ba.get<int32_t>(2 * (B + 1));
					HB_ALLOC_CHECK(ba);
					k_ecb_cap<<<nblk(n_ov, 256), 256, 0, ctx->stream>>>(n_ov, d_od, d_ch, d_aln, d_cap);
					if ((rc = hb_scan_u32_to_u64(ctx, d_cap, d_wboff, n_ov))) return rc;
					uint64_t wb_tot = 0; HB_CUDA(cudaMemcpyAsync(&wb_tot, d_wboff + n_ov, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
					hb_wl_t *d_wlb = ba.zero<hb_wl_t>(wb_tot + 1); HB_ALLOC_CHECK(ba);
					const uint64_t path_words1 = ctx->ecb_path_words; const int32_t merge_cw0 = ctx->ecb_cig_words; // tier 1's trace words per warp (tier 2: 16 x); cigar words of the first merge launch
					EcPrep *d_prep = ba.get<EcPrep>(n_ov + 1); uint32_t *d_nseg = ba.zero<uint32_t>(n_ov + 1); uint64_t *d_segoff = ba.get<uint64_t>(n_ov + 2);
					uint32_t *d_qn = ba.zero<uint32_t>(8); unsigned long long *d_spused = ba.zero<unsigned long long>(1);
					HB_ALLOC_CHECK(ba);
					EcCigArgs G; memset(&G, 0, sizeof(G));
					G.R = R; G.r0 = r0 + b0; G.n_ov = n_ov; G.desc = d_od; G.ch = d_ch; G.fc = d_fc; G.fc_grp_base = d_fcb; G.aln = d_aln; G.wlA = d_wl;
					G.chits = d_chits; G.ghits = d_hits; G.dp_half = B + 1; G.dp_t = d_dpt; G.dp_p = d_dpp; G.dp_f = d_dpf; G.e_rate = so->e_rate; G.w_l = w_l; G.gaps = so->gaps;
					G.prep = d_prep; G.nseg = d_nseg; G.seg_off = d_segoff; G.out = d_alnb; G.wl = d_wlb; G.wl_off = d_wboff; G.n_deferred = d_ndef; G.err = d_err;
					{
						ProfScope ps(ctx, "k_ecb_prep");
						if (n_ov) hb_k_ecb(HB_K_ECB_PREP, nblk(n_ov, 128), 128, ctx->stream, G);
					}
					HB_CUDA(cudaGetLastError());
					if ((rc = hb_scan_u32_to_u64(ctx, d_nseg, d_segoff, n_ov))) return rc;
					uint64_t n_seg = 0; HB_CUDA(cudaMemcpyAsync(&n_seg, d_segoff + n_ov, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
					if (n_seg >= (1ull << 32)) { hb_set_err(ctx, HB_E_OVERFLOW, "EC base alignment: >= 2^32 segments in one batch"); return HB_E_OVERFLOW; }
					ctx->counters[11] += n_seg;
					EcSeg *d_segs = ba.get<EcSeg>(n_seg + 1); uint32_t *d_q1 = ba.get<uint32_t>(n_seg + 1), *d_q2 = ba.get<uint32_t>(n_seg + 1), *d_qkey = ba.get<uint32_t>(n_seg + 1);
					HB_ALLOC_CHECK(ba);
					G.n_seg = n_seg; G.segs = d_segs;
					// ---- segments: tier 0 (private scratch) -> queue -> tier 1 (16 K trace words) -> queue -> tier 2 (the largest alignment the reference allows)
					uint64_t spool_cap = n_seg / 2 + 65536, spool_used = 0; uint16_t *d_spool = 0; uint32_t h_q[5] = { 0, 0, 0, 0, 0 };
					for (int attempt = 0;; attempt++) {
						d_spool = ba.get<uint16_t>(spool_cap); HB_ALLOC_CHECK(ba);
						HB_CUDA(cudaMemsetAsync(d_spused, 0, 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_qn, 0, 32, ctx->stream));
						G.spool = d_spool; G.spool_used = d_spused; G.spool_cap = spool_cap;
						// pre-pass: everything that needs no alignment; the rest goes to queue 0
						G.q_in = 0; G.q_in_n = 0; G.q_out = d_q2; G.q_out_n = d_qn + 0; G.q_key = d_qkey;
						{
							ProfScope ps(ctx, "k_ecb_seg_fast");
							if (n_seg) hb_k_ecb(HB_K_ECB_SEG_FAST, nblk(n_seg, 128), 128, ctx->stream, G);
						}
						HB_CUDA(cudaGetLastError());
						HB_CUDA(cudaMemcpyAsync(h_q, d_qn, 20, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
						G.q_key = 0;
						if (h_q[0]) { // the queue ordered by segment length (12-bit key): the 32 segments of a warp of tier 0 need about the same number of columns
							Arena sa(ctx);
							if ((rc = hb_sort_pairs_u32(ctx, d_qkey, d_q2, d_q1, h_q[0], 12))) return rc;
						}
						// alignment tiers 0..3: {trace words, band words, cigar runs, blocks of 128 threads}; tier 0 keeps its scratch private (local memory);
						// a segment that overflows one tier queues for the next (queues ping-pong between two arrays)
						// tier 1: thread / segment with global scratch (trace words per thread, 8-word band); tiers 2..3 give a WARP to a segment (k_ecb_seg_w): {trace words per
						// unit, cigar runs, units}; the largest holds the longest alignment the reference attempts (HB_MAX_SIN_L columns x 64 band words x 3 trace words)
						const struct { uint64_t pw; int32_t cw; unsigned units; const char *name; } TIER[4] = {
							{ 0, 0, 0, "k_ecb_seg" }, { path_words1, 256, (unsigned)ctx->sm_count * 8 * 128, "k_ecb_seg_tier1" }, { path_words1 * 64, 8192, (unsigned)ctx->sm_count * 8, "k_ecb_seg_tier2" },
							{ (uint64_t)HB_MW_MAXW * (2 + 3 * (uint64_t)HB_MAX_SIN_L), 65535, (unsigned)ctx->sm_count * 2, "k_ecb_seg_tier3" } };
						for (int tier = 0; tier <= 3 && h_q[tier]; tier++) {
							Arena sa(ctx);
							unsigned bl = (unsigned)(((uint64_t)h_q[tier] + 127) / 128);
							if (tier == 1) {
								bl = std::max(1u, std::min(bl, TIER[1].units / 128)); const uint64_t nt = (uint64_t)bl * 128;
								G.path = sa.get<uint64_t>(nt * TIER[1].pw); G.path_words = TIER[1].pw; G.vec = sa.get<uint64_t>(nt * 11 * 8); G.vstride = 8;
								G.cig_tmp = sa.get<uint16_t>(nt * (uint64_t)TIER[1].cw); G.cig_words = TIER[1].cw;
								if (sa.failed) return HB_E_WS;
							} else if (tier > 1) {
								const uint64_t nw = std::max<uint64_t>(4, std::min<uint64_t>(((uint64_t)h_q[tier] + 3) & ~3ull, TIER[tier].units)); bl = (unsigned)(nw / 4);
								G.path = sa.get<uint64_t>(nw * TIER[tier].pw); G.path_words = TIER[tier].pw; G.vec = sa.get<uint64_t>(nw * 2 * (uint64_t)HB_MW_MAXW); G.vstride = HB_MW_MAXW;
								G.cig_tmp = sa.get<uint16_t>(nw * (uint64_t)TIER[tier].cw); G.cig_words = TIER[tier].cw; G.work = sa.zero<uint32_t>(1);
								if (sa.failed) return HB_E_WS;
							}
							G.q_in = (tier & 1) ? d_q2 : d_q1; G.q_in_n = d_qn + tier; G.q_out = tier < 3 ? ((tier & 1) ? d_q1 : d_q2) : 0; G.q_out_n = d_qn + tier + 1;
							{
								ProfScope ps(ctx, TIER[tier].name);
								hb_k_ecb(tier == 0 ? HB_K_ECB_SEG : tier == 1 ? HB_K_ECB_SEG_G : HB_K_ECB_SEG_W, bl, 128, ctx->stream, G);
							}
							HB_CUDA(cudaGetLastError());
							HB_CUDA(cudaMemcpyAsync(h_q, d_qn, 20, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
						}
						HB_CUDA(cudaMemcpyAsync(&spool_used, d_spused, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
						HB_CUDA(cudaStreamSynchronize(ctx->stream));
						if (h_err2 & 32) { hb_set_err(ctx, HB_E_STATE, "EC base alignment: fake-cigar lookup failed"); return HB_E_STATE; }
						if (spool_used <= spool_cap && spool_used < (1ull << 32)) break;
						if (attempt >= 2 || spool_used >= (1ull << 32)) { hb_set_err(ctx, HB_E_OVERFLOW, "EC base alignment: segment cigar pool"); return HB_E_OVERFLOW; }
						spool_cap = spool_used + 65536; // the need is known now; the segments are recomputed (the chains stay refined)
					}
					ctx->counters[10] += h_q[3]; // segments that needed the largest scratch tier
					if (ctx->trace_ec) fprintf(stderr, "[hb] EC base alignment: %llu overlaps, %llu segments; queued for alignment %u, past tier 0 %u, past tier 1 %u, past tier 2 %u; segment cigar pool %llu\n",
					                                    (unsigned long long)n_ov, (unsigned long long)n_seg, h_q[0], h_q[1], h_q[2], h_q[3], (unsigned long long)spool_used);
					// ---- merge
					uint64_t poolb_cap = n_seg / 4 + 16 * n_ov + 65536, poolb_used = 0; uint16_t *d_poolb = 0; unsigned int n_def = 0;
					for (int attempt = 0;; attempt++) {
						Arena pa(ctx);
						d_poolb = pa.get<uint16_t>(poolb_cap);
						HB_CUDA(cudaMemsetAsync(d_pused, 0, 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_ndef, 0, 4, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_qn + 6, 0, 4, ctx->stream));
						G.pool = d_poolb; G.pool_used = d_pused; G.pool_cap = poolb_cap;
						G.rc_q = d_q1; G.rc_n = d_qn + 6; // the segment queues are free now: the merge kernel queues the overlaps that need the re-seeding rescue
						for (int pass = 0; pass < 2; pass++) {
							if (pass == 1 && !n_def) break;
							const unsigned bl = pass == 0 ? (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((n_ov + 63) / 64, (uint64_t)ctx->sm_count * 32)) : 4u; // 24 KB of cigar scratch per thread: 7 GB at most
							const uint64_t nt = (uint64_t)bl * 64; const int32_t cw = pass == 0 ? merge_cw0 : (1 << 20);
							Arena sa(ctx);
							G.cig_tmp = sa.get<uint16_t>(nt * 3 * (uint64_t)cw); G.cig_words = cw; G.pass = pass;
							if (sa.failed || pa.failed) return HB_E_WS;
							if (pass == 1) HB_CUDA(cudaMemsetAsync(d_ndef, 0, 4, ctx->stream));
							{
								ProfScope ps(ctx, pass == 0 ? "k_ecb_merge" : "k_ecb_merge_deferred");
								if (n_ov) hb_k_ecb(HB_K_ECB_MERGE, bl, 64, ctx->stream, G);
							}
							HB_CUDA(cudaGetLastError());
							HB_CUDA(cudaMemcpyAsync(&n_def, d_ndef, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
							if (pass == 1 && n_def) { hb_set_err(ctx, HB_E_OVERFLOW, "EC base alignment: %u overlaps exceed the largest scratch", n_def); return HB_E_OVERFLOW; }
						}
						// ---- the re-seeding rescue (rechain_aln_hc, Correct.cpp:17669): overlaps that keep an unaligned window of >= 512 bp on both reads.
						// Rare (a structural difference inside an accepted overlap); one thread per overlap with the largest aligner scratch, few threads.
						unsigned int n_rc = 0;
						HB_CUDA(cudaMemcpyAsync(&n_rc, d_qn + 6, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
						if (n_rc) {
							Arena sa(ctx);
							const int32_t cw = 65535, zcap = 2048, hcap = 16384; const uint64_t zc_cap = 1ull << 17;
							RcLaunch L; memset(&L, 0, sizeof(L));
							L.blocks = std::min<unsigned>((n_rc + 31) / 32, 2u); const uint64_t nt = (uint64_t)L.blocks * 32;
							L.R = R; L.r0 = r0 + b0; L.desc = d_od; L.ch = d_ch; L.fc = d_fc; L.fc_grp_base = d_fcb; L.aln = d_aln; L.wlA = d_wl; L.poolA = d_pool;
							L.e_rate = so->e_rate; L.w_l = w_l; L.gaps = so->gaps;
							L.out = d_alnb; L.wl = d_wlb; L.wl_off = d_wboff; L.pool = d_poolb; L.pool_used = d_pused; L.pool_cap = poolb_cap; L.err = d_err; L.rc_q = d_q1; L.rc_n = d_qn + 6;
							L.path_words = (uint64_t)HB_MW_MAXW * HB_MAX_SIN_L * 5; L.path = sa.get<uint64_t>(nt * L.path_words); L.vec = sa.get<uint64_t>(nt * 11 * (uint64_t)HB_MW_MAXW);
							L.cig_words = cw; L.cig3 = sa.get<uint16_t>(nt * 3 * (uint64_t)cw);
							L.zcap = zcap; L.zw = sa.get<hb_wl_t>(nt * (uint64_t)zcap); L.zc_cap = zc_cap; L.zc = sa.get<uint16_t>(nt * zc_cap); L.zc_used = sa.zero<unsigned long long>(nt);
							L.path1 = sa.get<uint64_t>(nt * (uint64_t)w_l * 5); L.cig1 = sa.get<uint16_t>(nt * HB_EC_CIG_TMP);
							L.hcap = hcap; L.h = sa.get<hb_hit_t>(nt * (uint64_t)hcap); L.t = sa.get<int64_t>(nt * (uint64_t)hcap); L.p = sa.get<int64_t>(nt * (uint64_t)hcap); L.f = sa.get<int32_t>(nt * (uint64_t)hcap);
							L.rs_b = sa.get<int32_t>(nt * 512); L.rs_f = sa.get<RsFrame>(nt * HB_RS_STACK);
							if (sa.failed || pa.failed) return HB_E_WS;
							{ const double tmp = expf((float)(-0.01 * (double)31 /* E_KHIT, ecovlp.cpp:10 */)); L.pen_gap = 0.5f; L.pen_skip = 0.0005f; L.pen_gap *= tmp; L.pen_skip *= tmp; } // set_lchain_dp_op, anchor.cpp:2272-2285
							{
								ProfScope ps(ctx, "k_ecb_rechain");
								HB_CUDA(hb_launch_ecb_rechain(L, ctx->stream));
							}
							HB_CUDA(cudaStreamSynchronize(ctx->stream));
							if (ctx->trace_ec) fprintf(stderr, "[hb] EC base alignment: %u overlaps through the re-seeding rescue\n", n_rc);
						}
						HB_CUDA(cudaMemcpyAsync(&poolb_used, d_pused, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream));
						HB_CUDA(cudaStreamSynchronize(ctx->stream));
						if (h_err2 & 32) { hb_set_err(ctx, HB_E_STATE, "EC base alignment: fake-cigar lookup failed"); return HB_E_STATE; }
						if (poolb_used <= poolb_cap && mode >= 7) { // phasing of the batch's reads over the step-C state that stays in HBM (row a13)
							std::vector<uint64_t> h_boff(nb + 1, 0);
							for (uint64_t i = 0; i < nb; i++) h_boff[i + 1] = h_boff[i] + hb_ph_bits_bytes(ctx->h_rlen[r0 + b0 + i]);
							uint64_t *d_boff = ba.get<uint64_t>(nb + 1); uint8_t *d_cnt = ba.zero<uint8_t>(h_boff[nb] + 8); PhOv *d_phov = ba.get<PhOv>(n_ov + 1);
							uint32_t *d_nacc = ba.zero<uint32_t>(nb + 1), *d_nsite = ba.zero<uint32_t>(nb + 1), *d_nev = ba.zero<uint32_t>(nb + 1); uint64_t *d_soff = ba.get<uint64_t>(nb + 2), *d_eoff = ba.get<uint64_t>(nb + 2);
							hb_phase_t *d_ph = ba.get<hb_phase_t>(n_ov + 1);
							HB_ALLOC_CHECK(ba);
							HB_CUDA(cudaMemcpyAsync(d_boff, h_boff.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
							PhArgs P; memset(&P, 0, sizeof(P));
							P.R = R; P.r0 = r0 + b0; P.nR = nb; P.o_off = d_ooff; P.desc = d_od; P.ch = d_ch; P.aln = d_aln; P.alnb = d_alnb; P.wl = d_wlb; P.pool = d_poolb;
							P.ov = d_phov; P.n_acc = d_nacc; P.b_off = d_boff; P.cnt = d_cnt; P.n_site = d_nsite; P.n_ev = d_nev; P.out = d_ph; P.err = d_err; P.work = ba.zero<uint32_t>(2); HB_ALLOC_CHECK(ba);
							const unsigned ph_blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nb + PH_WARPS - 1) / PH_WARPS, (uint64_t)ctx->sm_count * 8));
							{
								ProfScope ps(ctx, "k_ph_count");
								hb_k_ph(0, ph_blocks, ctx->stream, P);
							}
							HB_CUDA(cudaGetLastError());
							if ((rc = hb_scan_u32_to_u64(ctx, d_nsite, d_soff, nb)) || (rc = hb_scan_u32_to_u64(ctx, d_nev, d_eoff, nb))) return rc;
							uint64_t tot_s = 0, tot_e = 0;
							HB_CUDA(cudaMemcpyAsync(&tot_s, d_soff + nb, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&tot_e, d_eoff + nb, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
							P.site_off = d_soff; P.ev_off = d_eoff;
							P.site_pos = ba.get<uint32_t>(tot_s + 1); P.site_o = ba.get<uint32_t>(tot_s + nb + 2); P.ev = ba.get<PhEv>(tot_e + 2); P.ev2 = ba.get<PhEv>(tot_e + 2); P.snp = ba.get<PhSnp>(4 * tot_s + 1);
							P.ord = ba.get<uint64_t>(n_ov + 1); P.ov_o = ba.get<uint32_t>(n_ov + nb + 2);
							HB_ALLOC_CHECK(ba);
							{
								ProfScope ps(ctx, "k_ph_decide");
								hb_k_ph(1, ph_blocks, ctx->stream, P);
							}
							HB_CUDA(cudaGetLastError());
							HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
							if (h_err2 & 128) { hb_set_err(ctx, HB_E_OVERFLOW, "phasing: radix-sort stack"); return HB_E_OVERFLOW; }
							if (mode == 9) { // the round's paf[] / reverse_paf[] and the two read flags of the batch (rows a15, with the staged edit scripts)
								hb_ma_hit_t *d_rp = ba.get<hb_ma_hit_t>(n_ov + 1), *d_sp = ba.get<hb_ma_hit_t>(n_ov + 1); uint32_t *d_nrp = ba.zero<uint32_t>(nb + 1), *d_nsp = ba.zero<uint32_t>(nb + 1);
								uint64_t *d_srt = ba.get<uint64_t>(2 * n_ov + 2); uint8_t *d_fl = ba.zero<uint8_t>(2 * nb + 2); HB_ALLOC_CHECK(ba);
								{
									ProfScope ps(ctx, "k_ec_rpaf");
									k_ec_rpaf<<<nblk(nb, 64), 64, 0, ctx->stream>>>(R, r0 + b0, nb, d_ooff, d_ph, P.ord, d_rp, d_nrp, d_err);
								}
								const uint16_t *d_sc_use = ctx->d_scc; const uint64_t *d_scoff_use = ctx->d_scc_off; uint64_t sc_rid0 = 0;
								std::vector<uint16_t> h_scc_b; std::vector<uint64_t> h_scoff_b; unsigned long long h_nec_b = 0;
								if (so->cns) { // row a14: the edit scripts of the batch's reads by window consensus, on the device
									uint32_t *d_entcap = ba.zero<uint32_t>(nb + 1), *d_scn = ba.zero<uint32_t>(nb + 1); uint64_t *d_entoff = ba.get<uint64_t>(nb + 2), *d_slot = ba.get<uint64_t>(nb + 2), *d_scoff_b = ba.get<uint64_t>(nb + 2);
									uint8_t *d_status = ba.zero<uint8_t>(nb + 8); unsigned long long *d_nec = ba.zero<unsigned long long>(1); CnsOv *d_cov = ba.get<CnsOv>(n_ov + 1);
									HB_ALLOC_CHECK(ba);
									k_cns_cap<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_ooff, d_alnb, d_entcap);
									if ((rc = hb_scan_u32_to_u64(ctx, d_entcap, d_entoff, nb))) return rc;
									uint64_t tot_ent = 0; HB_CUDA(cudaMemcpyAsync(&tot_ent, d_entoff + nb, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
									const unsigned cblocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((nb + CNS_WARPS - 1) / CNS_WARPS, (uint64_t)ctx->sm_count * 3)); const uint64_t cwarps = (uint64_t)cblocks * CNS_WARPS;
									CnsEnt *d_ent = ba.get<CnsEnt>(tot_ent + 1); uint32_t *d_csrt = ba.get<uint32_t>(tot_ent + 1), *d_acta = ba.get<uint32_t>(tot_ent + 1), *d_actb = ba.get<uint32_t>(tot_ent + 1), *d_b32 = ba.get<uint32_t>(tot_ent + 1);
									uint64_t *d_key = ba.get<uint64_t>(tot_ent + 1); int32_t *d_rs = ba.get<int32_t>(cwarps * HB_CNS_RS_WORDS); uint32_t *d_cwork = ba.zero<uint32_t>(2);
									HB_ALLOC_CHECK(ba);
									std::vector<uint64_t> h_slot(nb + 1, 0); std::vector<uint32_t> h_scn(nb + 1); std::vector<uint64_t> h_scoff(nb + 1);
									uint16_t *d_scslot = 0; uint64_t mult = 1;
									for (int attempt = 0;; attempt++) {
										for (uint64_t i = 0; i < nb; i++) h_slot[i + 1] = h_slot[i] + (128 + ctx->h_rlen[r0 + b0 + i] / 16) * mult;
										d_scslot = ba.get<uint16_t>(h_slot[nb] + 1); HB_ALLOC_CHECK(ba);
										HB_CUDA(cudaMemcpyAsync(d_slot, h_slot.data(), (nb + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
										HB_CUDA(cudaMemsetAsync(d_nec, 0, 8, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_cwork, 0, 8, ctx->stream)); if (attempt) HB_CUDA(cudaMemsetAsync(d_err, 0, 4, ctx->stream)); // (bit 256 of the attempt before; no other bit was set)
										CnsArgs CA; memset(&CA, 0, sizeof(CA));
										CA.R = R; CA.r0 = r0 + b0; CA.nR = nb; CA.o_off = d_ooff; CA.ph = d_ph; CA.alnb = d_alnb; CA.wl = d_wlb; CA.pool = d_poolb; CA.ord = P.ord; CA.cov = d_cov; CA.ent_off = d_entoff; CA.ent = d_ent;
										CA.srt = d_csrt; CA.act_a = d_acta; CA.act_b = d_actb; CA.b32 = d_b32; CA.key = d_key; CA.rs = d_rs; CA.work = d_cwork; CA.out_off = d_slot; CA.out = d_scslot; CA.out_n = d_scn; CA.status = d_status; CA.nec = d_nec; CA.err = d_err;
										{
											ProfScope ps(ctx, "k_ec_cns");
											if (hb_k_cns(0, cblocks, ctx->stream, CA)) { hb_set_err(ctx, HB_E_CUDA, "window consensus: shared-memory attribute"); return HB_E_CUDA; }
										}
										HB_CUDA(cudaGetLastError());
										std::vector<uint8_t> h_st(nb + 1);
										HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_st.data(), d_status, nb, cudaMemcpyDeviceToHost, ctx->stream));
										HB_CUDA(cudaStreamSynchronize(ctx->stream));
										if (h_err2 & 128) { hb_set_err(ctx, HB_E_OVERFLOW, "dedup_chains: radix-sort stack"); return HB_E_OVERFLOW; }
										std::vector<uint32_t> h_queue; for (uint64_t i = 0; i < nb; i++) if (h_st[i] & 1) h_queue.push_back((uint32_t)i);
										if (!h_queue.empty()) { // second launch: the reads whose stretches need the graph consensus (cns_gen_full), each warp with a graph arena
											Arena sa(ctx);
											const unsigned gblocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((h_queue.size() + CNS_WARPS - 1) / CNS_WARPS, (uint64_t)ctx->sm_count * 2)); const uint64_t gthr = (uint64_t)gblocks * CNS_WARPS;
											CnsArgs CB = CA; CB.n_queue = (uint32_t)h_queue.size(); CB.g_nodes = ctx->cns_g_nodes; CB.g_arcs = ctx->cns_g_arcs; CB.g_nseq = 2048; CB.g_pcap = 8192; CB.g_ccap = 2048;
											uint32_t *d_queue = sa.get<uint32_t>(h_queue.size()); CB.queue = d_queue; CB.work = d_cwork + 1;
											CB.g_nd = sa.get<CnsNode>(gthr * CB.g_nodes); CB.g_arc = sa.get<CnsArc>(gthr * CB.g_arcs); CB.g_q = sa.get<uint32_t>(gthr * CB.g_nodes); CB.g_b32 = sa.get<uint32_t>(gthr * CB.g_arcs);
											CB.g_np = sa.get<uint32_t>(gthr * CB.g_nseq); CB.g_ns = sa.get<uint8_t>(gthr * CB.g_nseq); CB.g_path = sa.get<uint64_t>(gthr * CB.g_pcap); CB.g_vec = sa.get<uint64_t>(gthr * 32); CB.g_cig = sa.get<uint16_t>(gthr * CB.g_ccap);
											CB.rs = sa.get<int32_t>(gthr * HB_CNS_RS_WORDS);
											if (sa.failed) return HB_E_WS;
											HB_CUDA(cudaMemcpyAsync(d_queue, h_queue.data(), h_queue.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
											{
												ProfScope ps(ctx, "k_ec_cns_graph");
												if (hb_k_cns(1, gblocks, ctx->stream, CB)) { hb_set_err(ctx, HB_E_CUDA, "window consensus: shared-memory attribute"); return HB_E_CUDA; }
											}
											HB_CUDA(cudaGetLastError());
											HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
										}
										if (!(h_err2 & 256)) break;
										if (attempt >= 3) { hb_set_err(ctx, HB_E_OVERFLOW, "window consensus: edit-script buffer"); return HB_E_OVERFLOW; }
										mult *= 8; // some script outgrew its slot: all reads of the batch again with larger slots
									}
									if ((rc = hb_scan_u32_to_u64(ctx, d_scn, d_scoff_b, nb))) return rc;
									unsigned long long h_nec = 0;
									HB_CUDA(cudaMemcpyAsync(h_scoff.data(), d_scoff_b, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(&h_nec, d_nec, 8, cudaMemcpyDeviceToHost, ctx->stream));
									if (so->status) HB_CUDA(cudaMemcpyAsync(so->status + b0, d_status, nb, cudaMemcpyDeviceToHost, ctx->stream));
									HB_CUDA(cudaStreamSynchronize(ctx->stream));
									uint16_t *d_scd = ba.get<uint16_t>(h_scoff[nb] + 1); HB_ALLOC_CHECK(ba);
									k_sc_compact<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_slot, d_scoff_b, d_scslot, d_scd);
									HB_CUDA(cudaGetLastError());
									h_scc_b.resize(h_scoff[nb] + 1); // the batch's scripts; added to the pass's in batch order below
									if (h_scoff[nb]) HB_CUDA(cudaMemcpyAsync(h_scc_b.data(), d_scd, h_scoff[nb] * 2, cudaMemcpyDeviceToHost, ctx->stream));
									HB_CUDA(cudaStreamSynchronize(ctx->stream));
									h_scoff_b.swap(h_scoff); h_nec_b = h_nec;
									d_sc_use = d_scd; d_scoff_use = d_scoff_b; sc_rid0 = r0 + b0;
								}
								{
									ProfScope ps(ctx, "k_ec_spaf");
									k_ec_spaf<<<nblk(nb, 64), 64, 0, ctx->stream>>>(R, r0 + b0, nb, d_ooff, d_ph, d_alnb, d_wlb, d_poolb, d_sc_use, d_scoff_use, sc_rid0, P.ord, d_srt, d_sp, d_nsp, d_fl, d_err);
								}
								HB_CUDA(cudaGetLastError());
								std::vector<hb_ma_hit_t> h_rp(n_ov + 1), h_sp(n_ov + 1); std::vector<uint32_t> h_nrp(nb + 1), h_nsp(nb + 1);
								HB_CUDA(cudaMemcpyAsync(h_rp.data(), d_rp, n_ov * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_nrp.data(), d_nrp, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
								HB_CUDA(cudaMemcpyAsync(h_sp.data(), d_sp, n_ov * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_nsp.data(), d_nsp, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
								if (so->flags) HB_CUDA(cudaMemcpyAsync(so->flags + 2 * b0, d_fl, 2 * nb, cudaMemcpyDeviceToHost, ctx->stream));
								HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
								if (h_err2 & 128) { hb_set_err(ctx, HB_E_OVERFLOW, "dedup_chains: radix-sort stack"); return HB_E_OVERFLOW; }
								if ((rc = order.commit(bk, [&]() -> int { // the pass's lists and scripts grow in batch order
									if (so->cns) {
										const size_t base = so->h_scc->size(); so->h_scc->resize(base + h_scoff_b[nb]);
										if (h_scoff_b[nb]) memcpy(so->h_scc->data() + base, h_scc_b.data(), h_scoff_b[nb] * 2);
										for (uint64_t i = 0; i <= nb; i++) (*so->h_scc_off)[b0 + i] = base + h_scoff_b[i];
										so->n_corrected += h_nec_b;
									}
									for (uint64_t i = 0; i < nb; i++) {
										st3_off[b0 + i] = st3_n; st3_hit_off[b0 + i] = st3_nh;
										if ((so->rec && st3_n + h_nsp[i] > so->rec_cap) || (so->rec2 && st3_nh + h_nrp[i] > so->rec2_cap)) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity"); return HB_E_OVERFLOW; }
										if (so->rec) memcpy((hb_ma_hit_t *)so->rec + st3_n, h_sp.data() + h_ooff[i], h_nsp[i] * sizeof(hb_ma_hit_t));
										if (so->rec2) memcpy(so->rec2 + st3_nh, h_rp.data() + h_ooff[i], h_nrp[i] * sizeof(hb_ma_hit_t));
										st3_n += h_nsp[i]; st3_nh += h_nrp[i];
									}
									st3_off[b0 + nb] = st3_n; st3_hit_off[b0 + nb] = st3_nh;
									return HB_OK; }))) return rc;
								break;
							}
							if (mode == 8) { // the round's reverse_paf lists of the batch
								hb_ma_hit_t *d_rp = ba.get<hb_ma_hit_t>(n_ov + 1); uint32_t *d_nrp = ba.zero<uint32_t>(nb + 1); HB_ALLOC_CHECK(ba);
								{
									ProfScope ps(ctx, "k_ec_rpaf");
									k_ec_rpaf<<<nblk(nb, 64), 64, 0, ctx->stream>>>(R, r0 + b0, nb, d_ooff, d_ph, P.ord, d_rp, d_nrp, d_err);
								}
								HB_CUDA(cudaGetLastError());
								std::vector<hb_ma_hit_t> h_rp(n_ov + 1); std::vector<uint32_t> h_nrp(nb + 1);
								HB_CUDA(cudaMemcpyAsync(h_rp.data(), d_rp, n_ov * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_nrp.data(), d_nrp, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
								HB_CUDA(cudaMemcpyAsync(&h_err2, d_err, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
								if (h_err2 & 128) { hb_set_err(ctx, HB_E_OVERFLOW, "dedup_chains: radix-sort stack"); return HB_E_OVERFLOW; }
								for (uint64_t i = 0; i < nb; i++) {
									st3_off[b0 + i] = st3_n;
									if (so->rec && st3_n + h_nrp[i] > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity"); return HB_E_OVERFLOW; }
									if (so->rec) memcpy((hb_ma_hit_t *)so->rec + st3_n, h_rp.data() + h_ooff[i], h_nrp[i] * sizeof(hb_ma_hit_t));
									st3_n += h_nrp[i];
								}
								st3_off[b0 + nb] = st3_n;
								break;
							}
							if (so->rec && st3_n + n_ov > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity"); return HB_E_OVERFLOW; }
							if (so->rec) HB_CUDA(cudaMemcpyAsync((hb_phase_t *)so->rec + st3_n, d_ph, n_ov * sizeof(hb_phase_t), cudaMemcpyDeviceToHost, ctx->stream));
							HB_CUDA(cudaStreamSynchronize(ctx->stream));
							for (uint64_t i = 0; i <= nb; i++) st3_off[b0 + i] = st3_n + h_ooff[i];
							st3_n += n_ov;
							break;
						}
						if (poolb_used <= poolb_cap) {
							// dense window lists for the host
							uint32_t *d_wn = ba.zero<uint32_t>(n_ov + 1); uint64_t *d_dense = ba.get<uint64_t>(n_ov + 2); HB_ALLOC_CHECK(ba);
							k_ecb_wn<<<nblk(n_ov, 256), 256, 0, ctx->stream>>>(n_ov, d_alnb, d_wn);
							if ((rc = hb_scan_u32_to_u64(ctx, d_wn, d_dense, n_ov))) return rc;
							uint64_t nwd = 0; HB_CUDA(cudaMemcpyAsync(&nwd, d_dense + n_ov, 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
							hb_wl_t *d_wld = ba.get<hb_wl_t>(nwd + 1); HB_ALLOC_CHECK(ba);
							k_gather_wl<<<nblk(n_ov, 256), 256, 0, ctx->stream>>>(n_ov, d_alnb, d_dense, d_wlb, d_alnb2, d_wld);
							HB_CUDA(cudaGetLastError());
							if (so->rec && st3_n + n_ov > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity"); return HB_E_OVERFLOW; }
							if (so->wl && so->n_wl + nwd > so->wl_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "window-list output capacity"); return HB_E_OVERFLOW; }
							if (so->cig && so->n_cig + poolb_used > so->cig_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "cigar output capacity"); return HB_E_OVERFLOW; }
							if (so->rec) HB_CUDA(cudaMemcpyAsync((hb_alnb_t *)so->rec + st3_n, d_alnb2, n_ov * sizeof(hb_alnb_t), cudaMemcpyDeviceToHost, ctx->stream));
							if (so->wl) HB_CUDA(cudaMemcpyAsync(so->wl + so->n_wl, d_wld, nwd * sizeof(hb_wl_t), cudaMemcpyDeviceToHost, ctx->stream));
							if (so->cig) HB_CUDA(cudaMemcpyAsync(so->cig + so->n_cig, d_poolb, poolb_used * 2, cudaMemcpyDeviceToHost, ctx->stream));
							HB_CUDA(cudaStreamSynchronize(ctx->stream));
							if (so->rec) for (uint64_t i = 0; i < n_ov; i++) ((hb_alnb_t *)so->rec)[st3_n + i].w_off += so->n_wl;
							if (so->wl && so->n_cig) for (uint64_t i = 0; i < nwd; i++) if (so->wl[so->n_wl + i].clen) so->wl[so->n_wl + i].cidx += (uint32_t)so->n_cig;
							for (uint64_t i = 0; i <= nb; i++) st3_off[b0 + i] = st3_n + h_ooff[i];
							st3_n += n_ov; so->n_wl += nwd; so->n_cig += poolb_used;
							break;
						}
						if (attempt >= 2) { hb_set_err(ctx, HB_E_OVERFLOW, "EC base alignment: cigar pool"); return HB_E_OVERFLOW; }
						poolb_cap = poolb_used + 65536; n_def = 0; // the need is known now: only the merge is repeated
					}
					return HB_OK;
				}
				if (so->rec && st3_n + n_ov > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity"); return HB_E_OVERFLOW; }
				if (so->wl && so->n_wl + n_win > so->wl_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "window-list output capacity"); return HB_E_OVERFLOW; }
				if (so->cig && so->n_cig + pool_used > so->cig_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "cigar output capacity"); return HB_E_OVERFLOW; }
				if (so->rec) HB_CUDA(cudaMemcpyAsync((hb_aln_t *)so->rec + st3_n, d_aln, n_ov * sizeof(hb_aln_t), cudaMemcpyDeviceToHost, ctx->stream));
				if (so->wl) HB_CUDA(cudaMemcpyAsync(so->wl + so->n_wl, d_wl, n_win * sizeof(hb_wl_t), cudaMemcpyDeviceToHost, ctx->stream));
				if (so->cig) HB_CUDA(cudaMemcpyAsync(so->cig + so->n_cig, d_pool, pool_used * 2, cudaMemcpyDeviceToHost, ctx->stream));
				HB_CUDA(cudaStreamSynchronize(ctx->stream));
				if (so->rec) for (uint64_t i = 0; i < n_ov; i++) ((hb_aln_t *)so->rec)[st3_n + i].w_off += so->n_wl; // batch-relative -> global
				if (so->wl && so->n_cig) for (uint64_t i = 0; i < n_win; i++) if (so->wl[so->n_wl + i].clen) so->wl[so->n_wl + i].cidx += (uint32_t)so->n_cig;
				for (uint64_t i = 0; i <= nb; i++) st3_off[b0 + i] = st3_n + h_ooff[i];
				st3_n += n_ov; so->n_wl += n_win; so->n_cig += pool_used;
				return HB_OK;
			}
			if (so->rec && st3_n + n_win > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "window output capacity"); return HB_E_OVERFLOW; }
			if (so->rec) HB_CUDA(cudaMemcpyAsync((hb_win_t *)so->rec + st3_n, d_wout, n_win * sizeof(hb_win_t), cudaMemcpyDeviceToHost, ctx->stream));
			HB_CUDA(cudaStreamSynchronize(ctx->stream));
			for (uint64_t i = 0; i <= nb; i++) st3_off[b0 + i] = st3_n + h_woff[i];
			st3_n += n_win;
			return HB_OK;
		}
		if (mode == 3) { // assemble the stage-3 view of this batch and copy it out
			uint32_t *d_nh = ba.get<uint32_t>(nb + 1), *d_nf = ba.get<uint32_t>(nb + 1); uint64_t *d_och = ba.get<uint64_t>(nb + 2), *d_oh = ba.get<uint64_t>(nb + 2), *d_of = ba.get<uint64_t>(nb + 2), *d_fcb = ba.get<uint64_t>(n_slots + 1);
			HB_ALLOC_CHECK(ba);
			k_stage3_counts<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_coff, d_ch, d_idx, d_nol, d_nh, d_nf);
			k_fc_grp_base<<<std::max(1u, std::min(nblk(h_dirn, 128), (unsigned)ctx->sm_count * 16)), 128, 0, ctx->stream>>>(d_dir, d_dirn, d_aoff + b0, a_base, d_coff, CP.mcopy_num, CP.mcopy_khit_cutoff, d_fcb);
			if ((rc = hb_scan_u32_to_u64(ctx, d_nol, d_och, nb)) || (rc = hb_scan_u32_to_u64(ctx, d_nh, d_oh, nb)) || (rc = hb_scan_u32_to_u64(ctx, d_nf, d_of, nb))) return rc;
			std::vector<uint64_t> h_och(nb + 1), h_oh(nb + 1), h_of(nb + 1);
			HB_CUDA(cudaMemcpyAsync(h_och.data(), d_och, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_oh.data(), d_oh, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
			HB_CUDA(cudaMemcpyAsync(h_of.data(), d_of, (nb + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
			hb_chain_t *d_o1 = ba.get<hb_chain_t>(h_och[nb] + 1); hb_hit_t *d_o2 = ba.get<hb_hit_t>(h_oh[nb] + 1); uint64_t *d_o3 = ba.get<uint64_t>(h_of[nb] + 1);
			HB_ALLOC_CHECK(ba);
			k_stage3_fill<<<nblk(nb, 128), 128, 0, ctx->stream>>>(nb, d_coff, d_aoff + b0, a_base, d_ch, 0, d_idx, d_nol, d_chits, d_hits, d_fc, d_och, d_oh, d_of, d_o1, d_o2, d_o3, d_fcb);
			HB_CUDA(cudaGetLastError());
			if (so->rec && st3_n + h_och[nb] > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "chain output capacity"); return HB_E_OVERFLOW; }
			if (so->hits && st3_nh + h_oh[nb] > so->hit_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "chain-anchor output capacity"); return HB_E_OVERFLOW; }
			if (so->fc && st3_nf + h_of[nb] > so->fc_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "fake-cigar output capacity"); return HB_E_OVERFLOW; }
			if (so->rec) HB_CUDA(cudaMemcpyAsync((hb_chain_t *)so->rec + st3_n, d_o1, h_och[nb] * sizeof(hb_chain_t), cudaMemcpyDeviceToHost, ctx->stream));
			if (so->hits) HB_CUDA(cudaMemcpyAsync(so->hits + st3_nh, d_o2, h_oh[nb] * sizeof(hb_hit_t), cudaMemcpyDeviceToHost, ctx->stream));
			if (so->fc) HB_CUDA(cudaMemcpyAsync(so->fc + st3_nf, d_o3, h_of[nb] * 8, cudaMemcpyDeviceToHost, ctx->stream));
			HB_CUDA(cudaStreamSynchronize(ctx->stream));
			for (uint64_t i = 0; i <= nb; i++) { st3_off[b0 + i] = st3_n + h_och[i]; st3_hit_off[b0 + i] = st3_nh + h_oh[i]; st3_fc_off[b0 + i] = st3_nf + h_of[i]; }
			st3_n += h_och[nb]; st3_nh += h_oh[nb]; st3_nf += h_of[nb];
			return HB_OK;
		}

		// exact + merge (final pass)
		{
			ExactArgs E; E.R = R; E.r0 = r0 + b0; E.n_slots = n_slots; E.ch = d_ch; E.slot_read = d_slot_read; E.keep = d_keep; E.exact = d_exact;
			ProfScope ps(ctx, "k_exact");
			if (n_slots) k_exact<<<nblk(n_slots * 32, 256), 256, 0, ctx->stream>>>(E);
		}
		HB_CUDA(cudaGetLastError());
	TRACE("exact");
		const uint64_t g0 = r0 + b0, p0b = ctx->h_prev0_off[g0], p0e = ctx->h_prev0_off[g0 + nb], p1b = ctx->h_prev1_off[g0], p1e = ctx->h_prev1_off[g0 + nb];
		hb_ma_hit_t *d_in0 = ba.get<hb_ma_hit_t>(p0e - p0b + 1); uint64_t *d_i0off = ba.get<uint64_t>(nb + 2), *d_i1off = ba.get<uint64_t>(nb + 2), *d_ooff = ba.get<uint64_t>(nb + 2);
		uint32_t *d_cap = ba.get<uint32_t>(nb + 1);
		HB_ALLOC_CHECK(ba);
		HB_CUDA(cudaMemcpyAsync(d_in0, ctx->d_prev0 + p0b, (p0e - p0b) * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToDevice, ctx->stream)); // in0 is consumed (el reset, ecovlp.cpp:5103)
		k_rebase<<<nblk(nb + 1, 256), 256, 0, ctx->stream>>>(nb + 1, ctx->d_prev0_off + g0, p0b, d_i0off);
		k_rebase<<<nblk(nb + 1, 256), 256, 0, ctx->stream>>>(nb + 1, ctx->d_prev1_off + g0, p1b, d_i1off);
		k_cap_per_read<<<nblk(nb, 256), 256, 0, ctx->stream>>>(nb, d_sc, d_i0off, d_cap);
		rc = hb_scan_u32_to_u64(ctx, d_cap, d_ooff, nb); if (rc) return rc;
		const uint64_t o_tot = n_slots + (p0e - p0b);
		FinOv *d_ov = ba.get<FinOv>(o_tot + 1); uint64_t *d_srt = ba.get<uint64_t>((p0e - p0b) + (p1e - p1b) + 1);
		hb_ma_hit_t *d_o0 = ba.hi<hb_ma_hit_t>(o_tot + 1), *d_o1 = ba.hi<hb_ma_hit_t>(o_tot + 1); uint64_t *d_ooff_keep = ba.hi<uint64_t>(nb + 2);
		HB_ALLOC_CHECK(ba);
		{
			MergeArgs M; M.R = R; M.r0 = g0; M.nR = nb; M.c_off = d_coff; M.ch = d_ch; M.idx = d_idx; M.n_ol = d_nol; M.exact = d_exact;
			M.in0 = d_in0; M.in0_off = d_i0off; M.in1 = ctx->d_prev1 + p1b; M.in1_off = d_i1off; M.ov = d_ov; M.srt = d_srt; M.o_off = d_ooff;
			M.out0 = d_o0; M.out1 = d_o1; M.m0 = d_m0 + b0; M.m1 = d_m1 + b0; M.stat = d_stat;
			ProfScope ps(ctx, "k_merge");
			k_merge_warp<<<nblk(nb, MERGE_WARPS), MERGE_WARPS * 32, MERGE_WARPS * MERGE_SMEM_PER_WARP, ctx->stream>>>(M);
		}
		HB_CUDA(cudaGetLastError());
	TRACE("merge");
		HB_CUDA(cudaMemcpyAsync(d_ooff_keep, d_ooff, (nb + 1) * 8, cudaMemcpyDeviceToDevice, ctx->stream));
		BatchRes br; br.o0 = d_o0; br.o1 = d_o1; br.ooff = d_ooff_keep; br.b0 = b0; br.b1 = b1;
		HB_CUDA(cudaStreamSynchronize(ctx->stream)); // (the gather below runs on the pass's stream)
		order.commit(bk, [&]() { bres.push_back(br); return HB_OK; });
		return HB_OK;
	};
	if ((rc = run_batches(pass_ctx, mode, batches.size(), order, batch_body))) return rc;

	TRACE("batches done");
	{ unsigned long long h_dbg[2] = { 0, 0 }; HB_CUDA(cudaMemcpyAsync(h_dbg, d_stat + 8, 16, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream)); ctx->counters[6] = h_dbg[0]; ctx->counters[7] = h_dbg[1]; }
	if (mode == 2) {
		if (so->off) for (uint64_t i = 0; i <= nR; i++) so->off[i] = h_aoff[i];
		if (so->rec) {
			if (h_aoff[nR] > so->rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "anchor output capacity"); return HB_E_OVERFLOW; }
			HB_CUDA(cudaMemcpyAsync(so->rec, d_all_hits, h_aoff[nR] * sizeof(hb_hit_t), cudaMemcpyDeviceToHost, ctx->stream));
		}
		HB_CUDA(cudaStreamSynchronize(ctx->stream));
		return HB_OK;
	}
	if (mode >= 4) { if (so->off) memcpy(so->off, st3_off.data(), (nR + 1) * 8); if (mode == 9 && so->off2) memcpy(so->off2, st3_hit_off.data(), (nR + 1) * 8); return HB_OK; }
	if (mode == 3) {
		if (so->off) memcpy(so->off, st3_off.data(), (nR + 1) * 8);
		if (so->hit_off) memcpy(so->hit_off, st3_hit_off.data(), (nR + 1) * 8);
		if (so->fc_off) memcpy(so->fc_off, st3_fc_off.data(), (nR + 1) * 8);
		return HB_OK;
	}

	// ---- final pass: dense result arrays, kept resident in the context
	if (ctx->outoff_cap < nR + 2) {
		cudaFree(ctx->d_out0_off); cudaFree(ctx->d_out1_off); ctx->d_out0_off = ctx->d_out1_off = 0; ctx->outoff_cap = 0;
		HB_CUDA(cudaMalloc((void **)&ctx->d_out0_off, (nR + 2) * 8)); HB_CUDA(cudaMalloc((void **)&ctx->d_out1_off, (nR + 2) * 8)); ctx->outoff_cap = nR + 2;
	}
	if ((rc = hb_scan_u32_to_u64(ctx, d_m0, ctx->d_out0_off, nR)) || (rc = hb_scan_u32_to_u64(ctx, d_m1, ctx->d_out1_off, nR))) return rc;
	HB_CUDA(cudaMemcpyAsync(&ctx->n_out0, ctx->d_out0_off + nR, 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(&ctx->n_out1, ctx->d_out1_off + nR, 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	if (ctx->out0_cap < ctx->n_out0 + 1) { cudaFree(ctx->d_out0); ctx->d_out0 = 0; ctx->out0_cap = 0; HB_CUDA(cudaMalloc((void **)&ctx->d_out0, (ctx->n_out0 + 1 + (ctx->n_out0 >> 3)) * sizeof(hb_ma_hit_t))); ctx->out0_cap = ctx->n_out0 + 1 + (ctx->n_out0 >> 3); }
	if (ctx->out1_cap < ctx->n_out1 + 1) { cudaFree(ctx->d_out1); ctx->d_out1 = 0; ctx->out1_cap = 0; HB_CUDA(cudaMalloc((void **)&ctx->d_out1, (ctx->n_out1 + 1 + (ctx->n_out1 >> 3)) * sizeof(hb_ma_hit_t))); ctx->out1_cap = ctx->n_out1 + 1 + (ctx->n_out1 >> 3); }
	for (auto &br : bres) {
		uint64_t nb = br.b1 - br.b0;
		ProfScope ps(ctx, "k_gather_ma");
		k_gather_ma<<<nblk(nb * 32, 256), 256, 0, ctx->stream>>>(nb, br.ooff, ctx->d_out0_off + br.b0, br.o0, ctx->d_out0);
		k_gather_ma<<<nblk(nb * 32, 256), 256, 0, ctx->stream>>>(nb, br.ooff, ctx->d_out1_off + br.b0, br.o1, ctx->d_out1);
	}
	HB_CUDA(cudaGetLastError());
	TRACE("gather");
	ctx->out_reads = nR;
	unsigned long long h_stat[16] = { 0 };
	HB_CUDA(cudaMemcpyAsync(h_stat, d_stat, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->counters[6] = h_stat[8]; ctx->counters[7] = h_stat[9];
	if (h_stat[6]) { hb_set_err(ctx, HB_E_OVERFLOW, "a read has more overlaps than the in-kernel radix sort stack supports"); return HB_E_OVERFLOW; }
	if (stat_out) { // forward, reverse, strong, weak, exact, no_l_indel, inexact (ecovlp.cpp:6173-6179)
		for (int i = 0; i < 6; i++) stat_out[i] = h_stat[i];
		stat_out[6] = h_stat[0] - h_stat[4];
	}
	return HB_OK;
}

// k_gather_ma reads d_off relative to the batch: d_off + b0 gives absolute dense offsets; fine since it subtracts nothing.

// ---------------------------------------------------------------------------
// public stage APIs
// ---------------------------------------------------------------------------
extern "C" int hb_sketch(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *off, hb_mz_t *rec, uint64_t rec_cap)
{
	cudaSetDevice(ctx->device);
	if (r1 > ctx->n_reads || r0 > r1) { hb_set_err(ctx, HB_E_ARG, "read range out of bounds"); return HB_E_ARG; }
	uint64_t nR = r1 - r0; DevSketch sk; Arena ar(ctx);
	if (nR == 0) { off[0] = 0; return HB_OK; }
	int rc = hb_run_sketch(ctx, r0, r1, 0, &sk); if (rc) return rc;
	HB_CUDA(cudaMemcpyAsync(off, sk.off, (nR + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	if (rec) {
		if (sk.total > rec_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "minimizer output capacity"); return HB_E_OVERFLOW; }
		HB_CUDA(cudaMemcpyAsync(rec, sk.mz, sk.total * sizeof(hb_mz_t), cudaMemcpyDeviceToHost, ctx->stream));
	}
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	hb_ws_reset(ctx);
	return HB_OK;
}
extern "C" int hb_anchors(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *off, hb_hit_t *rec, uint64_t rec_cap)
{
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap;
	return run_pass(ctx, r0, r1, 2, 0.02, &so, 0);
}
extern "C" int hb_chains(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, uint64_t *off, hb_chain_t *rec, uint64_t rec_cap,
                         uint64_t *hit_off, hb_hit_t *hits, uint64_t hit_cap, uint64_t *fc_off, uint64_t *fc, uint64_t fc_cap)
{
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.hit_off = hit_off; so.hits = hits; so.hit_cap = hit_cap; so.fc_off = fc_off; so.fc = fc; so.fc_cap = fc_cap;
	return run_pass(ctx, r0, r1, 3, bw_thres, &so, 0);
}

extern "C" int hb_windows(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, uint64_t *off, hb_win_t *rec, uint64_t rec_cap)
{
	if (w_l < 8 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be >= 8 and e_rate in [0,1]"); return HB_E_ARG; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.e_rate = e_rate; so.w_l = w_l;
	return run_pass(ctx, r0, r1, 4, bw_thres, &so, 0);
}

static int stage_prev(hb_ctx *ctx, const hb_ma_hit_t *p0, const uint64_t *o0, const hb_ma_hit_t *p1, const uint64_t *o1);
extern "C" int hb_ec_align(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l,
                           uint64_t *off, hb_aln_t *rec, uint64_t rec_cap, hb_wl_t *wl, uint64_t wl_cap, uint16_t *cig, uint64_t cig_cap,
                           uint64_t *n_wl, uint64_t *n_cig)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.e_rate = e_rate; so.w_l = w_l;
	so.wl = wl; so.wl_cap = wl_cap; so.cig = cig; so.cig_cap = cig_cap;
	int rc = run_pass(ctx, r0, r1, 5, bw_thres, &so, 0);
	if (n_wl) *n_wl = so.n_wl;
	if (n_cig) *n_cig = so.n_cig;
	return rc;
}

extern "C" int hb_ec_cigar(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t gaps,
                           uint64_t *off, hb_alnb_t *rec, uint64_t rec_cap, hb_wl_t *wl, uint64_t wl_cap, uint16_t *cig, uint64_t cig_cap,
                           uint64_t *n_wl, uint64_t *n_cig)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.e_rate = e_rate; so.w_l = w_l;
	so.wl = wl; so.wl_cap = wl_cap; so.cig = cig; so.cig_cap = cig_cap; so.gaps = gaps & 1; so.use_prev = (gaps >> 1) & 1;
	int rc = run_pass(ctx, r0, r1, 6, bw_thres, &so, 0);
	if (n_wl) *n_wl = so.n_wl;
	if (n_cig) *n_cig = so.n_cig;
	return rc;
}

extern "C" int hb_ec_reverse_paf(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, uint64_t *off, hb_ma_hit_t *rec, uint64_t rec_cap)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.e_rate = e_rate; so.w_l = w_l; so.gaps = 1;
	return run_pass(ctx, r0, r1, 8, bw_thres, &so, 0);
}
extern "C" int hb_ec_round_lists(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t use_prev,
                                 uint64_t *src_off, hb_ma_hit_t *src, uint64_t src_cap, uint64_t *rev_off, hb_ma_hit_t *rev, uint64_t rev_cap, uint8_t *flags)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	if (!ctx->n_reads || ctx->scc_reads != ctx->n_reads) { hb_set_err(ctx, HB_E_STATE, "edit scripts are not staged for the resident reads (hb_ec_stage_scc)"); return HB_E_STATE; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = src_off; so.rec = src; so.rec_cap = src_cap; so.off2 = rev_off; so.rec2 = rev; so.rec2_cap = rev_cap; so.flags = flags;
	so.e_rate = e_rate; so.w_l = w_l; so.gaps = 1; so.use_prev = use_prev ? 1 : 0;
	return run_pass(ctx, r0, r1, 9, bw_thres, &so, 0);
}
extern "C" int hb_ec_stage_scc(hb_ctx_t *ctx, const uint16_t *scc, const uint64_t *scc_off);
extern "C" int hb_ec_round(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t use_prev,
                           uint64_t *src_off, hb_ma_hit_t *src, uint64_t src_cap, uint64_t *rev_off, hb_ma_hit_t *rev, uint64_t rev_cap, uint8_t *flags,
                           uint64_t *scc_off, uint16_t *scc, uint64_t scc_cap, uint8_t *status, uint64_t *n_corrected)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	std::vector<uint16_t> h_scc; std::vector<uint64_t> h_off;
	StageOut so; memset(&so, 0, sizeof(so)); so.off = src_off; so.rec = src; so.rec_cap = src_cap; so.off2 = rev_off; so.rec2 = rev; so.rec2_cap = rev_cap; so.flags = flags;
	so.e_rate = e_rate; so.w_l = w_l; so.gaps = 1; so.use_prev = use_prev ? 1 : 0; so.cns = 1; so.h_scc = &h_scc; so.h_scc_off = &h_off; so.status = status;
	int rc = run_pass(ctx, r0, r1, 9, bw_thres, &so, 0); if (rc) return rc;
	const uint64_t nR = r1 - r0;
	if (h_off.size() != nR + 1) h_off.assign(nR + 1, 0);
	if (scc_off) memcpy(scc_off, h_off.data(), (nR + 1) * 8);
	if (scc) {
		if (h_scc.size() > scc_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "edit-script output capacity: need %llu", (unsigned long long)h_scc.size()); return HB_E_OVERFLOW; }
		if (h_scc.size()) memcpy(scc, h_scc.data(), h_scc.size() * 2);
	}
	if (n_corrected) *n_corrected = so.n_corrected;
	if (r0 == 0 && r1 == ctx->n_reads && nR) return hb_ec_stage_scc(ctx, h_scc.data(), h_off.data()); // the whole store: the scripts stay in HBM for hb_ec_apply / hb_ec_update_paf
	return HB_OK;
}
extern "C" int hb_ec_stage_prev(hb_ctx_t *ctx, const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off)
{ // R_INF.paf[] of the previous EC round, flattened; only the source list matters to gen_hc_r_alin_ea
	static const hb_ma_hit_t none = {}; std::vector<uint64_t> z(ctx->n_reads + 1, 0);
	return stage_prev(ctx, prev_src, prev_src_off, &none, z.data());
}
extern "C" int hb_ec_phase(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, uint64_t *off, hb_phase_t *rec, uint64_t rec_cap)
{
	if (w_l < 8 || w_l > 4096 || e_rate < 0 || e_rate > 1) { hb_set_err(ctx, HB_E_ARG, "window length must be in [8,4096] and e_rate in [0,1]"); return HB_E_ARG; }
	StageOut so; memset(&so, 0, sizeof(so)); so.off = off; so.rec = rec; so.rec_cap = rec_cap; so.e_rate = e_rate; so.w_l = w_l; so.gaps = 1;
	return run_pass(ctx, r0, r1, 7, bw_thres, &so, 0);
}

// ---------------------------------------------------------------------------
// final pass
// ---------------------------------------------------------------------------
static int stage_prev(hb_ctx *ctx, const hb_ma_hit_t *p0, const uint64_t *o0, const hb_ma_hit_t *p1, const uint64_t *o1)
{
	uint64_t n = ctx->n_reads;
	free_prev(ctx);
	ctx->h_prev0_off.assign(o0, o0 + n + 1); ctx->h_prev1_off.assign(o1, o1 + n + 1);
	ctx->n_prev0 = o0[n]; ctx->n_prev1 = o1[n];
	HB_CUDA(cudaMalloc((void **)&ctx->d_prev0, (ctx->n_prev0 + 1) * sizeof(hb_ma_hit_t))); HB_CUDA(cudaMalloc((void **)&ctx->d_prev1, (ctx->n_prev1 + 1) * sizeof(hb_ma_hit_t)));
	HB_CUDA(cudaMalloc((void **)&ctx->d_prev0_off, (n + 2) * 8)); HB_CUDA(cudaMalloc((void **)&ctx->d_prev1_off, (n + 2) * 8));
	HB_CUDA(cudaMemcpyAsync(ctx->d_prev0, p0, ctx->n_prev0 * sizeof(hb_ma_hit_t), cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(ctx->d_prev1, p1, ctx->n_prev1 * sizeof(hb_ma_hit_t), cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(ctx->d_prev0_off, o0, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(ctx->d_prev1_off, o1, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

extern "C" int hb_cal_ov_r_resident(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *n_src, uint64_t *n_rev, uint64_t *stat)
{
	int rc = run_pass(ctx, r0, r1, 0, 0.001, 0, stat); // bw 0.001: ecovlp.cpp:3957
	if (rc) return rc;
	if (n_src) *n_src = ctx->n_out0;
	if (n_rev) *n_rev = ctx->n_out1;
	return HB_OK;
}

extern "C" int hb_cal_ov_r(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off, const hb_ma_hit_t *prev_rev, const uint64_t *prev_rev_off,
                           hb_ma_hit_t *out_src, uint64_t *out_src_off, uint64_t out_src_cap, hb_ma_hit_t *out_rev, uint64_t *out_rev_off, uint64_t out_rev_cap, uint64_t *stat)
{
	cudaSetDevice(ctx->device);
	if (prev_src_off) { int rc = stage_prev(ctx, prev_src, prev_src_off, prev_rev, prev_rev_off); if (rc) return rc; }
	int rc = run_pass(ctx, r0, r1, 0, 0.001, 0, stat); if (rc) return rc;
	uint64_t nR = r1 - r0;
	if (ctx->n_out0 > out_src_cap || ctx->n_out1 > out_rev_cap) { hb_set_err(ctx, HB_E_OVERFLOW, "overlap output capacity: need %llu / %llu", (unsigned long long)ctx->n_out0, (unsigned long long)ctx->n_out1); return HB_E_OVERFLOW; }
	if (nR == 0) { out_src_off[0] = out_rev_off[0] = 0; return HB_OK; }
	HB_CUDA(cudaMemcpyAsync(out_src_off, ctx->d_out0_off, (nR + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(out_rev_off, ctx->d_out1_off, (nR + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(out_src, ctx->d_out0, ctx->n_out0 * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(out_rev, ctx->d_out1, ctx->n_out1 * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

// ---------------------------------------------------------------------------
// window alignment
// ---------------------------------------------------------------------------
extern "C" int hb_ed_semi_64(hb_ctx_t *ctx, uint64_t n, const char *pat, const uint64_t *pat_off, const char *txt, const uint64_t *txt_off,
                             const int32_t *thre, const int32_t *abs_diag, int32_t *err, int32_t *pe)
{
	cudaSetDevice(ctx->device);
	if (n == 0) return HB_OK;
	size_t need = pat_off[n] + txt_off[n] + n * 48 + (1 << 20);
	if (ctx->ws_cap < need) { ctx->ws_need = need; int rc = hb_ws_grow(ctx); if (rc) return rc; }
	hb_ws_reset(ctx);
	Arena ar(ctx);
	for (uint64_t i = 0; i < n; i++) if (thre[i] < 0 || thre[i] > 31 || abs_diag[i] < 0 || abs_diag[i] > 2 * thre[i]) { hb_set_err(ctx, HB_E_ARG, "case %llu: thre must be in [0,31], abs_diag in [0,2*thre]", (unsigned long long)i); return HB_E_ARG; }
	char *d_p = ar.get<char>(pat_off[n] + 1), *d_t = ar.get<char>(txt_off[n] + 1); uint64_t *d_po = ar.get<uint64_t>(n + 1), *d_to = ar.get<uint64_t>(n + 1);
	int32_t *d_th = ar.get<int32_t>(n), *d_ab = ar.get<int32_t>(n), *d_e = ar.get<int32_t>(n), *d_pe = ar.get<int32_t>(n);
	HB_ALLOC_CHECK(ar);
	HB_CUDA(cudaMemcpyAsync(d_p, pat, pat_off[n], cudaMemcpyHostToDevice, ctx->stream)); HB_CUDA(cudaMemcpyAsync(d_t, txt, txt_off[n], cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(d_po, pat_off, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream)); HB_CUDA(cudaMemcpyAsync(d_to, txt_off, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(d_th, thre, n * 4, cudaMemcpyHostToDevice, ctx->stream)); HB_CUDA(cudaMemcpyAsync(d_ab, abs_diag, n * 4, cudaMemcpyHostToDevice, ctx->stream));
	{
		ProfScope ps(ctx, "k_ed_semi64");
		k_ed_semi64<<<nblk(n, 128), 128, 0, ctx->stream>>>(n, d_p, d_po, d_t, d_to, d_th, d_ab, d_e, d_pe);
	}
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaMemcpyAsync(err, d_e, n * 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(pe, d_pe, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	return HB_OK;
}

// ---------------------------------------------------------------------------
// instrumentation
// ---------------------------------------------------------------------------
extern "C" int hb_profile(const hb_ctx_t *ctx, const char **names, uint64_t *launches, double *ms, int cap)
{
	int n = 0;
	for (auto &p : ctx->prof) { if (n >= cap) break; names[n] = p.name; launches[n] = p.launches; ms[n] = p.ms; n++; }
	return n;
}
extern "C" void hb_profile_reset(hb_ctx_t *ctx) { ctx->prof.clear(); }
extern "C" int hb_last_pass_ms(const hb_ctx_t *ctx, double *ms) { *ms = ctx->last_pass_ms; return HB_OK; }
extern "C" int hb_stage_profile(const hb_ctx_t *ctx, const char **names, uint64_t *launches, double *ms, int cap, uint64_t *counters, int ccap)
{ // per-kernel launch counts / device time and the counters summed over every pass of the last hb_stage_run
	int n = 0;
	for (auto &p : ctx->prof_stage) { if (n >= cap) break; names[n] = p.name; launches[n] = p.launches; ms[n] = p.ms; n++; }
	for (int i = 0; i < ccap && i < 12; i++) counters[i] = ctx->stage_counters[i];
	return n;
}
extern "C" int hb_counters(const hb_ctx_t *ctx, uint64_t *c, int cap)
{
	int n = cap < 12 ? cap : 12;
	for (int i = 0; i < n; i++) c[i] = ctx->counters[i];
	return n;
}
