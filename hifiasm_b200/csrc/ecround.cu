// ecround.cu — the steps that close an error-correction round on the resident read store:
//   hb_ec_stage_scc    the round's edit scripts (scc.a[i], ecovlp.cpp:101) -> HBM
//   hb_ec_apply        sl_ec_r / worker_sl_ec (ecovlp.cpp:6402, 5965): every read with its script applied -> a new read store in HBM
//   hb_ec_update_paf   cal_update_ec_multiple / worker_update_dc_ec (6095, 3808): exact intervals remapped and re-checked
//   hb_ec_post_rev     worker_hap_post_rev (3866): reads reverse-complemented, both lists flipped
//   hb_reads_download  the resident read store back in the All_reads layout
// Host code is plumbing (allocation, copies, launches); the per-read / per-record work is in hb_ecround.cuh.
#include <algorithm>
#include "hb_internal.h"
#include <chrono>
#include "hb_ecround.cuh"

static inline unsigned nblk(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

__global__ void k_sl_len(uint64_t n, const uint16_t *__restrict__ sc, const uint64_t *__restrict__ sc_off, const uint32_t *__restrict__ old_len, uint32_t *__restrict__ new_len, uint32_t *__restrict__ changed)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
	hb_sl_len(sc + sc_off[i], (uint32_t)(sc_off[i + 1] - sc_off[i]), old_len[i], &new_len[i], &changed[i]);
}
// thread / read: a changed read is rebuilt base by base through its script; an unchanged one is copied.  N positions go to tmp_npos at the
// read's OLD list offset (a corrected read never has more Ns than before: edits only write A/C/G/T)
__global__ void k_sl_apply(DevReads R, const uint16_t *__restrict__ sc, const uint64_t *__restrict__ sc_off, const uint32_t *__restrict__ changed, const uint64_t *__restrict__ new_off,
                           uint8_t *__restrict__ new_packed, uint32_t *__restrict__ tmp_npos, uint32_t *__restrict__ nn_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= R.n) return;
	uint8_t *dst = new_packed + new_off[i]; uint32_t *np = tmp_npos + R.noff[i];
	if (changed[i]) { nn_out[i] = hb_sl_apply(hb_rd_view(R, i, 0), sc + sc_off[i], (uint32_t)(sc_off[i + 1] - sc_off[i]), dst, np); return; }
	const uint8_t *src = R.packed + R.off[i]; const uint32_t nb = R.len[i] / 4 + 1, nn = (uint32_t)(R.noff[i + 1] - R.noff[i]);
	for (uint32_t k = 0; k < nb; k++) dst[k] = src[k];
	for (uint32_t k = 0; k < nn; k++) np[k] = R.npos[R.noff[i] + k];
	nn_out[i] = nn;
}
__global__ void k_npos_compact(uint64_t n, const uint64_t *__restrict__ old_noff, const uint64_t *__restrict__ new_noff, const uint32_t *__restrict__ tmp, uint32_t *__restrict__ out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
	const uint64_t a = old_noff[i], b = new_noff[i], c = new_noff[i + 1] - b;
	for (uint64_t k = 0; k < c; k++) out[b + k] = tmp[a + k];
}
// block / read (grid-stride): one thread per output byte of the reverse complement; the N list is mirrored
__global__ void k_rc_reads(DevReads R, uint8_t *__restrict__ out_packed, uint32_t *__restrict__ out_npos)
{
	for (uint64_t i = blockIdx.x; i < R.n; i += gridDim.x) {
		const RdView v = hb_rd_view(R, i, 0); uint8_t *dst = out_packed + R.off[i]; const uint32_t nb = v.len / 4 + 1;
		for (uint32_t j = threadIdx.x; j < nb; j += blockDim.x) dst[j] = hb_rc_byte(v, j);
		for (uint32_t k = threadIdx.x; k < v.nn; k += blockDim.x) out_npos[R.noff[i] + k] = v.len - 1 - v.npos[v.nn - 1 - k];
	}
}
// thread / record of paf[]: the query read id travels in the record (qns >> 32)
__global__ void k_update_dc(DevReads R, uint64_t n_rec, hb_ma_hit_t *__restrict__ paf, const uint16_t *__restrict__ sc, const uint64_t *__restrict__ sc_off, unsigned long long *__restrict__ cnt)
{
	const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (k >= n_rec) return;
	hb_ma_hit_t z = paf[k];
	const uint32_t e = hb_update_dc(R, z.qns >> 32, &z, sc, sc_off);
	paf[k] = z;
	atomicAdd(&cnt[e ? 0 : 1], 1ull);
}
__global__ void k_flip_paf(uint64_t n, const uint32_t *__restrict__ rlen, hb_ma_hit_t *__restrict__ paf, const uint64_t *__restrict__ off, uint32_t *__restrict__ n_out)
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
	n_out[i] = hb_flip_paf(rlen, i, paf + off[i], (uint32_t)(off[i + 1] - off[i]));
}

struct DevBuf { // scoped device allocation (these calls run once per round: plain cudaMalloc is fine)
	void *p; DevBuf() : p(0) {} ~DevBuf() { if (p) cudaFree(p); }
	int alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 16) == cudaSuccess ? 0 : -1; }
	template <typename T> T *as() { return (T *)p; }
	void *take() { void *q = p; p = 0; return q; }
};
#define HB_DEVALLOC(buf, bytes) do { if ((buf).alloc(bytes)) { cudaGetLastError(); hb_set_err(ctx, HB_E_NOMEM, "%s:%d device allocation of %llu bytes failed", __FILE__, __LINE__, (unsigned long long)(bytes)); return HB_E_NOMEM; } } while (0)

extern "C" int hb_ec_stage_scc(hb_ctx_t *ctx, const uint16_t *scc, const uint64_t *scc_off)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads;
	if (!n || !scc_off) { hb_set_err(ctx, HB_E_ARG, "no reads resident / no offsets"); return HB_E_ARG; }
	for (uint64_t i = 0; i < n; i++) if (scc_off[i + 1] < scc_off[i]) { hb_set_err(ctx, HB_E_ARG, "edit-script offsets must ascend"); return HB_E_ARG; }
	cudaFree(ctx->d_scc); cudaFree(ctx->d_scc_off); ctx->d_scc = 0; ctx->d_scc_off = 0; ctx->scc_reads = 0;
	const uint64_t tot = scc_off[n];
	HB_CUDA(cudaMalloc((void **)&ctx->d_scc, (tot + 8) * 2)); HB_CUDA(cudaMalloc((void **)&ctx->d_scc_off, (n + 2) * 8));
	if (tot) HB_CUDA(cudaMemcpyAsync(ctx->d_scc, scc, tot * 2, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(ctx->d_scc_off, scc_off, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	ctx->scc_reads = n; ctx->scc_total = tot;
	return HB_OK;
}

extern "C" int hb_ec_apply(hb_ctx_t *ctx, uint64_t *n_changed, uint64_t *total_bases)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads;
	if (!n || ctx->scc_reads != n) { hb_set_err(ctx, HB_E_STATE, "edit scripts are not staged for the resident reads (hb_ec_stage_scc)"); return HB_E_STATE; }
	DevReads R = hb_dev_reads(ctx);
	DevBuf b_len, b_chg, b_nn, b_tmp;
	HB_DEVALLOC(b_len, (n + 1) * 4); HB_DEVALLOC(b_chg, (n + 1) * 4); HB_DEVALLOC(b_nn, (n + 1) * 4); HB_DEVALLOC(b_tmp, (ctx->n_npos + 1) * 4);
	{
		ProfScope ps(ctx, "k_sl_len");
		k_sl_len<<<nblk(n, 128), 128, 0, ctx->stream>>>(n, ctx->d_scc, ctx->d_scc_off, ctx->d_rlen, b_len.as<uint32_t>(), b_chg.as<uint32_t>());
	}
	HB_CUDA(cudaGetLastError());
	std::vector<uint32_t> h_len(n), h_chg(n), h_nn(n);
	HB_CUDA(cudaMemcpyAsync(h_len.data(), b_len.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_chg.data(), b_chg.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	std::vector<uint64_t> off(n + 1), noff(n + 1, 0); uint64_t o = 0, tb = 0, nc = 0;
	for (uint64_t i = 0; i < n; i++) {
		if (h_len[i] >= (1u << 27)) { hb_set_err(ctx, HB_E_OVERFLOW, "corrected read %llu longer than 2^27", (unsigned long long)i); return HB_E_OVERFLOW; }
		off[i] = o; o += (((uint64_t)h_len[i] / 4 + 1) + 31) & ~31ull; tb += h_len[i]; nc += h_chg[i];
	}
	off[n] = o;
	const uint64_t pc = o + 64 + (o >> 4), rcap = n + 1 + (n >> 4);
	DevBuf n_packed, n_roff, n_noff, n_npos;
	HB_DEVALLOC(n_packed, pc); HB_DEVALLOC(n_roff, (rcap + 1) * 8); HB_DEVALLOC(n_noff, (rcap + 1) * 8);
	HB_CUDA(cudaMemsetAsync(n_packed.p, 0, pc, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(n_roff.p, off.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	{
		ProfScope ps(ctx, "k_sl_apply");
		k_sl_apply<<<nblk(n, 64), 64, 0, ctx->stream>>>(R, ctx->d_scc, ctx->d_scc_off, b_chg.as<uint32_t>(), n_roff.as<uint64_t>(), n_packed.as<uint8_t>(), b_tmp.as<uint32_t>(), b_nn.as<uint32_t>());
	}
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaMemcpyAsync(h_nn.data(), b_nn.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	for (uint64_t i = 0; i < n; i++) noff[i + 1] = noff[i] + h_nn[i];
	const uint64_t npc = noff[n] + 1 + (noff[n] >> 2);
	HB_DEVALLOC(n_npos, npc * 4);
	HB_CUDA(cudaMemcpyAsync(n_noff.p, noff.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	k_npos_compact<<<nblk(n, 128), 128, 0, ctx->stream>>>(n, ctx->d_noff, n_noff.as<uint64_t>(), b_tmp.as<uint32_t>(), n_npos.as<uint32_t>());
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaMemcpyAsync(ctx->d_rlen, b_len.p, n * 4, cudaMemcpyDeviceToDevice, ctx->stream)); // d_rlen keeps its capacity (the number of reads does not change)
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	// swap the new store in
	cudaFree(ctx->d_packed); cudaFree(ctx->d_roff); cudaFree(ctx->d_noff); cudaFree(ctx->d_npos);
	ctx->d_packed = (uint8_t *)n_packed.take(); ctx->d_roff = (uint64_t *)n_roff.take(); ctx->d_noff = (uint64_t *)n_noff.take(); ctx->d_npos = (uint32_t *)n_npos.take();
	hb_pt_destroy(ctx); // the position index describes the reads as they were: a pass without a fresh hb_pt_gen now fails loudly ("no position index") instead of chaining stale positions
	ctx->packed_cap = pc; ctx->reads_cap = std::min<uint64_t>(ctx->reads_cap, rcap); ctx->npos_cap = npc;
	ctx->packed_bytes = o; ctx->n_npos = noff[n]; ctx->total_bases = tb;
	for (uint64_t i = 0; i < n; i++) ctx->h_rlen[i] = h_len[i];
	if (n_changed) *n_changed = nc;
	if (total_bases) *total_bases = tb;
	return HB_OK;
}

extern "C" int hb_reads_download(hb_ctx_t *ctx, uint64_t *read_length, uint8_t *packed, uint64_t packed_cap, uint64_t *n_off, uint64_t *n_pos, uint64_t n_pos_cap)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads;
	if (!n) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	uint64_t need = 0; for (uint64_t i = 0; i < n; i++) need += ctx->h_rlen[i] / 4 + 1;
	if (read_length) for (uint64_t i = 0; i < n; i++) read_length[i] = ctx->h_rlen[i];
	if (packed) {
		if (packed_cap < need) { hb_set_err(ctx, HB_E_OVERFLOW, "packed-read output capacity: need %llu bytes", (unsigned long long)need); return HB_E_OVERFLOW; }
		std::vector<uint8_t> h(ctx->packed_bytes + 16); std::vector<uint64_t> off(n + 1);
		HB_CUDA(cudaMemcpyAsync(h.data(), ctx->d_packed, ctx->packed_bytes, cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(cudaMemcpyAsync(off.data(), ctx->d_roff, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(cudaStreamSynchronize(ctx->stream));
		uint64_t o = 0;
		for (uint64_t i = 0; i < n; i++) { const uint64_t nb = ctx->h_rlen[i] / 4 + 1; memcpy(packed + o, h.data() + off[i], nb); o += nb; }
	}
	if (n_off) {
		HB_CUDA(cudaMemcpyAsync(n_off, ctx->d_noff, (n + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	if (n_pos) {
		if (n_pos_cap < ctx->n_npos) { hb_set_err(ctx, HB_E_OVERFLOW, "N-position output capacity: need %llu", (unsigned long long)ctx->n_npos); return HB_E_OVERFLOW; }
		std::vector<uint32_t> h(ctx->n_npos + 1);
		if (ctx->n_npos) HB_CUDA(cudaMemcpyAsync(h.data(), ctx->d_npos, ctx->n_npos * 4, cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(cudaStreamSynchronize(ctx->stream));
		for (uint64_t i = 0; i < ctx->n_npos; i++) n_pos[i] = h[i];
	}
	return HB_OK;
}

extern "C" int hb_ec_update_paf(hb_ctx_t *ctx, hb_ma_hit_t *paf, const uint64_t *paf_off, uint64_t *n_exact, uint64_t *n_inexact)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads;
	if (!n || ctx->scc_reads != n) { hb_set_err(ctx, HB_E_STATE, "edit scripts are not staged for the resident reads (hb_ec_stage_scc)"); return HB_E_STATE; }
	const uint64_t tot = paf_off[n];
	for (uint64_t i = 0; i < n; i++) for (uint64_t k = paf_off[i]; k < paf_off[i + 1]; k++)
		if ((paf[k].qns >> 32) != i || paf[k].tn >= n) { hb_set_err(ctx, HB_E_ARG, "paf[%llu]: record %llu names another query or an unknown target", (unsigned long long)i, (unsigned long long)k); return HB_E_ARG; }
	unsigned long long h_cnt[2] = { 0, 0 };
	if (tot) {
		DevBuf b_paf, b_cnt;
		HB_DEVALLOC(b_paf, tot * sizeof(hb_ma_hit_t)); HB_DEVALLOC(b_cnt, 16);
		HB_CUDA(cudaMemcpyAsync(b_paf.p, paf, tot * sizeof(hb_ma_hit_t), cudaMemcpyHostToDevice, ctx->stream)); HB_CUDA(cudaMemsetAsync(b_cnt.p, 0, 16, ctx->stream));
		{
			ProfScope ps(ctx, "k_update_dc");
			k_update_dc<<<nblk(tot, 128), 128, 0, ctx->stream>>>(hb_dev_reads(ctx), tot, b_paf.as<hb_ma_hit_t>(), ctx->d_scc, ctx->d_scc_off, b_cnt.as<unsigned long long>());
		}
		HB_CUDA(cudaGetLastError());
		HB_CUDA(cudaMemcpyAsync(paf, b_paf.p, tot * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaMemcpyAsync(h_cnt, b_cnt.p, 16, cudaMemcpyDeviceToHost, ctx->stream));
		HB_CUDA(cudaStreamSynchronize(ctx->stream));
	}
	if (n_exact) *n_exact = h_cnt[0];
	if (n_inexact) *n_inexact = h_cnt[1];
	return HB_OK;
}

static int flip_list(hb_ctx *ctx, hb_ma_hit_t *paf, uint64_t *off)
{
	const uint64_t n = ctx->n_reads, tot = off[n];
	for (uint64_t k = 0; k < tot; k++) if (paf[k].tn >= n) { hb_set_err(ctx, HB_E_ARG, "overlap record %llu names an unknown target", (unsigned long long)k); return HB_E_ARG; }
	DevBuf b_paf, b_off, b_no; std::vector<uint32_t> h_no(n);
	HB_DEVALLOC(b_paf, tot * sizeof(hb_ma_hit_t)); HB_DEVALLOC(b_off, (n + 1) * 8); HB_DEVALLOC(b_no, (n + 1) * 4);
	if (tot) HB_CUDA(cudaMemcpyAsync(b_paf.p, paf, tot * sizeof(hb_ma_hit_t), cudaMemcpyHostToDevice, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(b_off.p, off, (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
	{
		ProfScope ps(ctx, "k_flip_paf");
		k_flip_paf<<<nblk(n, 128), 128, 0, ctx->stream>>>(n, ctx->d_rlen, b_paf.as<hb_ma_hit_t>(), b_off.as<uint64_t>(), b_no.as<uint32_t>());
	}
	HB_CUDA(cudaGetLastError());
	if (tot) HB_CUDA(cudaMemcpyAsync(paf, b_paf.p, tot * sizeof(hb_ma_hit_t), cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaMemcpyAsync(h_no.data(), b_no.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	uint64_t w = 0; // lists only shrink: compact in place, front to back
	for (uint64_t i = 0; i < n; i++) {
		const uint64_t s = off[i]; off[i] = w;
		if (w != s) memmove(paf + w, paf + s, (size_t)h_no[i] * sizeof(hb_ma_hit_t));
		w += h_no[i];
	}
	off[n] = w;
	return HB_OK;
}

extern "C" int hb_ec_post_rev(hb_ctx_t *ctx, hb_ma_hit_t *paf, uint64_t *paf_off, hb_ma_hit_t *rpaf, uint64_t *rpaf_off)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads; int rc;
	if (!n) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	DevBuf n_packed, n_npos;
	HB_DEVALLOC(n_packed, ctx->packed_cap); HB_DEVALLOC(n_npos, ctx->npos_cap * 4); // (allocations first: a failure must not leave the lists flipped and the reads not)
	if (paf_off && (rc = flip_list(ctx, paf, paf_off))) return rc;   // flip_paf_rc only needs the read lengths, which the reverse complement keeps
	if (rpaf_off && (rc = flip_list(ctx, rpaf, rpaf_off))) return rc;
	HB_CUDA(cudaMemsetAsync(n_packed.p, 0, ctx->packed_cap, ctx->stream));
	{
		ProfScope ps(ctx, "k_rc_reads");
		k_rc_reads<<<(unsigned)std::max<uint64_t>(1, std::min<uint64_t>(n, (uint64_t)ctx->sm_count * 16)), 256, 0, ctx->stream>>>(hb_dev_reads(ctx), n_packed.as<uint8_t>(), n_npos.as<uint32_t>());
	}
	HB_CUDA(cudaGetLastError());
	HB_CUDA(cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_packed); cudaFree(ctx->d_npos);
	ctx->d_packed = (uint8_t *)n_packed.take(); ctx->d_npos = (uint32_t *)n_npos.take();
	hb_pt_destroy(ctx); // target positions and strands changed: the index must be rebuilt before the next pass
	return HB_OK;
}

// ---- cal_ec_r (ecovlp.h:13; ecovlp.cpp:6268-6309) as one call on the resident store: cal_ec_multiple -> sl_ec_r -> cal_update_ec_multiple ->
// worker_hap_post_rev, in that order, every step on the device.  Host code only passes buffers along.
extern "C" int hb_ec_stage_prev(hb_ctx_t *ctx, const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off);
extern "C" int hb_ec_round(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t use_prev,
                           uint64_t *src_off, hb_ma_hit_t *src, uint64_t src_cap, uint64_t *rev_off, hb_ma_hit_t *rev, uint64_t rev_cap, uint8_t *flags,
                           uint64_t *scc_off, uint16_t *scc, uint64_t scc_cap, uint8_t *status, uint64_t *n_corrected);
extern "C" int hb_cal_ec_r(hb_ctx_t *ctx, uint64_t round, uint64_t n_round, uint64_t is_sv, double e_rate, int32_t w_l,
                           const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off,
                           hb_ma_hit_t *out_src, uint64_t *out_src_off, uint64_t out_src_cap, hb_ma_hit_t *out_rev, uint64_t *out_rev_off, uint64_t out_rev_cap,
                           uint8_t *flags, uint8_t *status, uint64_t *tot_b, uint64_t *tot_e, uint64_t *n_exact, uint64_t *n_inexact)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads; int rc;
	if (!n) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	if (n_round) { hb_set_err(ctx, HB_E_ARG, "cal_sec_ec_multiple (number_of_pround > 0, CommandLines.cpp:281) is not supported"); return HB_E_ARG; }
	if (!out_src || !out_src_off || !out_rev || !out_rev_off) { hb_set_err(ctx, HB_E_ARG, "output buffers are required"); return HB_E_ARG; }
	// HB_TRACE: the host clock around the five steps of the round (stderr; the device is drained at each stamp)
	auto stamp = [&](const char *what) { static thread_local double t_prev = 0; if (!ctx->trace) return; cudaStreamSynchronize(ctx->stream); const double t = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); if (what) fprintf(stderr, "[hb trace] cal_ec_r %-12s %9.2f ms\n", what, t - t_prev); t_prev = t; };
	stamp(0);
	if ((rc = hb_ec_stage_prev(ctx, prev_src, prev_src_off))) return rc;                                   // gen_hc_r_alin_ea reads the previous paf[i] (ecovlp.cpp:3288)
	stamp("stage_prev");
	if (tot_b) *tot_b = ctx->total_bases;                                                                  // cnt[0]: bases of the reads as they enter the round (3276)
	uint64_t nec = 0;
	if ((rc = hb_ec_round(ctx, 0, n, ctx->opt.is_ont ? 0.05 : 0.02, e_rate, w_l, 1, out_src_off, out_src, out_src_cap, out_rev_off, out_rev, out_rev_cap, flags, 0, 0, 0, status, &nec))) return rc;
	if (tot_e) *tot_e = nec;
	stamp("ec_round");
	if ((rc = hb_ec_apply(ctx, 0, 0))) return rc;                                                          // sl_ec_r
	stamp("apply");
	if ((rc = hb_ec_update_paf(ctx, out_src, out_src_off, n_exact, n_inexact))) return rc;                // cal_update_ec_multiple
	stamp("update_paf");
	if (!is_sv || (round & 1)) if ((rc = hb_ec_post_rev(ctx, out_src, out_src_off, out_rev, out_rev_off))) return rc; // ecovlp.cpp:6293-6295
	stamp("post_rev");
	return HB_OK;
}
