// stage.cu — the whole overlap / error-correction stage as ONE C-ABI call (host side, C++): what ha_assemble() runs between reading the reads and
// building the string graph (Assembly.cpp:2076-2108): ha_ft_gen, number_of_round x ha_ec (ha_pt_gen + cal_ec_r, Assembly.cpp:996-1030), ha_ec_ff
// (ha_pt_gen + cal_ov_r, Assembly.cpp:1942-1959).  Nothing is computed here: every step is one of the library's own entry points on the resident store.
//
// More than one GPU (one process per GPU, reads + index replicated): the query reads of every pass are sharded contiguously over the
// ranks; an EC round ends with the ONE exchange step of the stage — every rank needs every read's edit script (to rebuild its replica of the store) and
// both overlap lists (the next round's exact shortcut, the final pass) — which is an all-gather of one variable-length blob per rank.  The transport is
// the caller's (hb_allgather_fn): NCCL through torch.distributed in bench.py / stage.py, ncclAllGather or MPI in a C++ host.
#include "hb_internal.h"
#include <chrono>

static double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

struct StageBuf { // the stage's host-side state between calls: lists of all reads, grown on demand
	std::vector<hb_ma_hit_t> src, rev, t_src, t_rev; std::vector<uint64_t> src_off, rev_off, t_so, t_ro, scc_off; std::vector<uint16_t> scc; std::vector<uint8_t> flags, status, fc, ab, blob, all;
	uint8_t *xbuf = 0; uint64_t xcap = 0; // page-locked exchange buffer (send | recv)
	~StageBuf() { if (xbuf) cudaFreeHost(xbuf); }
};
static StageBuf *stage_buf(hb_ctx *ctx) { if (!ctx->stage_buf) ctx->stage_buf = new StageBuf(); return (StageBuf *)ctx->stage_buf; }
void hb_stage_buf_free(hb_ctx *ctx) { delete (StageBuf *)ctx->stage_buf; ctx->stage_buf = 0; }

static void shard(uint64_t n, int rank, int world, uint64_t *r0, uint64_t *r1)
{ const uint64_t per = (n + world - 1) / world; *r0 = std::min<uint64_t>(n, (uint64_t)rank * per); *r1 = std::min<uint64_t>(n, (uint64_t)(rank + 1) * per); }

// all-gather of one variable-length byte string per rank -> all (concatenation in rank order), sizes[world]
static int gather_blobs(hb_ctx *ctx, int world, hb_allgather_fn ag, void *user, const std::vector<uint8_t> &mine, std::vector<uint8_t> &all, std::vector<uint64_t> &sizes)
{
	sizes.assign(world, 0); uint64_t sz = mine.size();
	if (ag(user, &sz, sizes.data(), 8)) { hb_set_err(ctx, HB_E_STATE, "the all-gather callback failed (sizes)"); return HB_E_STATE; }
	uint64_t mx = 8; for (int r = 0; r < world; r++) mx = std::max(mx, sizes[r]);
	mx = (mx + 15) & ~15ull;
	// the exchange buffers are page-locked (and kept): the transport copies them to and from the device
	StageBuf &B = *stage_buf(ctx); const uint64_t need = mx * (uint64_t)(world + 1);
	if (B.xcap < need) { if (B.xbuf) cudaFreeHost(B.xbuf); B.xbuf = 0; B.xcap = 0; if (cudaHostAlloc((void **)&B.xbuf, need + need / 4, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); hb_set_err(ctx, HB_E_NOMEM, "page-locked exchange buffer"); return HB_E_NOMEM; } B.xcap = need + need / 4; }
	uint8_t *send = B.xbuf, *recv = B.xbuf + mx; if (sz) memcpy(send, mine.data(), sz);
	if (ag(user, send, recv, mx)) { hb_set_err(ctx, HB_E_STATE, "the all-gather callback failed (payload)"); return HB_E_STATE; }
	uint64_t tot = 0; for (int r = 0; r < world; r++) tot += sizes[r];
	all.resize(tot); uint64_t o = 0;
	for (int r = 0; r < world; r++) { if (sizes[r]) memcpy(all.data() + o, recv + (uint64_t)r * mx, sizes[r]); o += sizes[r]; }
	return HB_OK;
}
template <typename T> static void put(std::vector<uint8_t> &b, const T *p, uint64_t n) { const size_t o = b.size(); b.resize(o + n * sizeof(T)); if (n) memcpy(b.data() + o, p, n * sizeof(T)); }
template <typename T> static const T *take(const uint8_t *&p, uint64_t n) { const T *q = (const T *)p; p += n * sizeof(T); return q; }

// one EC round on this rank's shard + the exchange + the closing steps on the replica (cal_ec_r, ecovlp.cpp:6268-6309)
static int ec_round_sharded(hb_ctx *ctx, StageBuf &B, uint64_t round, uint64_t is_sv, int rank, int world, hb_allgather_fn ag, void *user, uint64_t *tot_e, uint64_t *n_unfinished, double *ms_exchange)
{
	const uint64_t n = ctx->n_reads; uint64_t r0, r1; shard(n, rank, world, &r0, &r1); const uint64_t nl = r1 - r0; int rc;
	if ((rc = hb_ec_stage_prev(ctx, B.src.empty() ? (const hb_ma_hit_t *)B.src_off.data() : B.src.data(), B.src_off.data()))) return rc;
	uint64_t cap = 96 * nl + 4096, scc_cap = 0, nec = 0;
	for (uint64_t i = r0; i < r1; i++) scc_cap += ctx->h_rlen[i] / 8 + 64;
	std::vector<uint64_t> so(nl + 1), ro(nl + 1), co(nl + 1); std::vector<uint8_t> fl(2 * nl + 2), st(nl + 1);
	for (int attempt = 0;; attempt++) { // capacities are guesses: an overflow is answered by larger buffers, on every rank alike or not (the round itself has no collective)
		B.t_src.resize(cap); B.t_rev.resize(cap); B.scc.resize(scc_cap);
		rc = hb_ec_round(ctx, r0, r1, ctx->opt.is_ont ? 0.05 : 0.02, 0.04, 775, 1, so.data(), B.t_src.data(), cap, ro.data(), B.t_rev.data(), cap, fl.data(), co.data(), B.scc.data(), scc_cap, st.data(), &nec);
		if (rc != HB_E_OVERFLOW || attempt >= 4) break;
		cap *= 4; scc_cap *= 4;
	}
	uint64_t ok = rc == HB_OK ? 1 : 0; // a rank that failed must not leave the others waiting in the exchange: the first word of every blob says so
	std::vector<uint8_t> &blob = B.blob; blob.clear();
	put(blob, &ok, 1); put(blob, &nl, 1); put(blob, &nec, 1);
	if (ok) { put(blob, so.data(), nl + 1); put(blob, ro.data(), nl + 1); put(blob, co.data(), nl + 1); put(blob, fl.data(), 2 * nl); put(blob, st.data(), nl);
		while (blob.size() & 7) blob.push_back(0);
		put(blob, B.t_src.data(), so[nl]); put(blob, B.t_rev.data(), ro[nl]); put(blob, B.scc.data(), co[nl]); }
	std::vector<uint64_t> sizes; const double t0 = wall_ms();
	int grc = gather_blobs(ctx, world, ag, user, blob, B.all, sizes); *ms_exchange += wall_ms() - t0;
	if (grc) return grc;
	{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) { if (!*(const uint64_t *)p) { if (rc == HB_OK) { hb_set_err(ctx, HB_E_STATE, "EC round %llu failed on rank %d", (unsigned long long)round, r); rc = HB_E_STATE; } } p += sizes[r]; } }
	if (rc) return rc;
	// unpack in rank order = read order
	B.t_so.assign(n + 1, 0); B.t_ro.assign(n + 1, 0); B.scc_off.assign(n + 1, 0); B.fc.assign(n, 0); B.ab.assign(n, 0); B.status.assign(n, 0);
	uint64_t ts = 0, tr = 0, tc = 0; *tot_e = 0;
	{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) { const uint8_t *q = p; take<uint64_t>(q, 1); const uint64_t m = *take<uint64_t>(q, 1); take<uint64_t>(q, 1); const uint64_t *a = take<uint64_t>(q, m + 1), *b = take<uint64_t>(q, m + 1), *c = take<uint64_t>(q, m + 1); ts += a[m]; tr += b[m]; tc += c[m]; p += sizes[r]; } }
	B.src.resize(ts + 1); B.rev.resize(tr + 1); B.scc.resize(tc + 1);
	uint64_t rb = 0, os = 0, orv = 0, oc = 0;
	{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) {
		const uint8_t *q = p; take<uint64_t>(q, 1); const uint64_t m = *take<uint64_t>(q, 1); *tot_e += *take<uint64_t>(q, 1);
		const uint64_t *a = take<uint64_t>(q, m + 1), *b = take<uint64_t>(q, m + 1), *c = take<uint64_t>(q, m + 1); const uint8_t *f = take<uint8_t>(q, 2 * m), *s = take<uint8_t>(q, m);
		q = p + (((uint64_t)(q - p) + 7) & ~7ull);
		const hb_ma_hit_t *xs = take<hb_ma_hit_t>(q, a[m]), *xr = take<hb_ma_hit_t>(q, b[m]); const uint16_t *xc = take<uint16_t>(q, c[m]);
		for (uint64_t i = 0; i < m; i++) { B.t_so[rb + i] = os + a[i]; B.t_ro[rb + i] = orv + b[i]; B.scc_off[rb + i] = oc + c[i]; B.fc[rb + i] = f[2 * i]; B.ab[rb + i] = f[2 * i + 1]; B.status[rb + i] = s[i]; }
		if (a[m]) memcpy(B.src.data() + os, xs, a[m] * sizeof(hb_ma_hit_t)); if (b[m]) memcpy(B.rev.data() + orv, xr, b[m] * sizeof(hb_ma_hit_t)); if (c[m]) memcpy(B.scc.data() + oc, xc, c[m] * 2);
		rb += m; os += a[m]; orv += b[m]; oc += c[m]; p += sizes[r];
	} }
	if (rb != n) { hb_set_err(ctx, HB_E_STATE, "the shards of the EC round do not add up to the read set"); return HB_E_STATE; }
	B.t_so[n] = os; B.t_ro[n] = orv; B.scc_off[n] = oc; B.src_off = B.t_so; B.rev_off = B.t_ro;
	for (uint64_t i = 0; i < n; i++) if (B.status[i]) ++*n_unfinished;
	// closing steps on the replica: sl_ec_r, cal_update_ec_multiple, worker_hap_post_rev (rows a16-a18)
	if ((rc = hb_ec_stage_scc(ctx, B.scc.data(), B.scc_off.data())) || (rc = hb_ec_apply(ctx, 0, 0)) || (rc = hb_ec_update_paf(ctx, B.src.data(), B.src_off.data(), 0, 0))) return rc;
	if (!is_sv || (round & 1)) if ((rc = hb_ec_post_rev(ctx, B.src.data(), B.src_off.data(), B.rev.data(), B.rev_off.data()))) return rc;
	return HB_OK;
}

extern "C" int hb_stage_run(hb_ctx_t *ctx, int n_round, int rank, int world, hb_allgather_fn ag, void *user, hb_stage_result_t *res)
{
	cudaSetDevice(ctx->device);
	const uint64_t n = ctx->n_reads; int rc, hom = 0, het = 0;
	if (!n) { hb_set_err(ctx, HB_E_STATE, "no reads resident"); return HB_E_STATE; }
	if (n_round < 0 || n_round > 8 || world < 1 || rank < 0 || rank >= world || (world > 1 && !ag)) { hb_set_err(ctx, HB_E_ARG, "hb_stage_run: 0..8 rounds, 0 <= rank < world, an all-gather callback when world > 1"); return HB_E_ARG; }
	memset(res, 0, sizeof(*res));
	StageBuf &B = *stage_buf(ctx);
	hb_stage_prof_begin(ctx);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, ctx->stream);
	const double t_begin = wall_ms(); double t = t_begin;
	if ((rc = hb_ft_gen(ctx, &hom))) return rc;                                     // Assembly.cpp:2081-2085
	hb_opt_update_cov(&ctx->opt, hom);
	res->ms_ft = wall_ms() - t;
	B.src.clear(); B.rev.clear(); B.src_off.assign(n + 1, 0); B.rev_off.assign(n + 1, 0); B.fc.assign(n, 0); B.ab.assign(n, 0);
	for (int k = 0; k < n_round; k++) {                                              // ha_ec, Assembly.cpp:996-1030
		t = wall_ms();
		if ((rc = hb_pt_gen(ctx, &hom, &het))) return rc;
		ctx->opt.hom_cov = hom; ctx->opt.het_cov = het;
		res->ms_pt[k] = wall_ms() - t; t = wall_ms();
		const uint64_t is_sv = k == n_round - 1 ? 1 : 0; uint64_t tot_e = 0;
		if (world == 1) {
			uint64_t cap = std::max<uint64_t>(2 * B.src_off[n], 96 * n) + 4096, tb = 0, ne = 0, ni = 0;
			B.flags.assign(2 * n + 2, 0); B.status.assign(n + 1, 0); B.t_so.assign(n + 1, 0); B.t_ro.assign(n + 1, 0);
			for (int attempt = 0;; attempt++) {
				B.t_src.resize(cap); B.t_rev.resize(cap);
				rc = hb_cal_ec_r(ctx, (uint64_t)k, 0, is_sv, 0.04, 775, B.src.empty() ? (const hb_ma_hit_t *)B.src_off.data() : B.src.data(), B.src_off.data(), B.t_src.data(), B.t_so.data(), cap,
				                 B.t_rev.data(), B.t_ro.data(), cap, B.flags.data(), B.status.data(), &tb, &tot_e, &ne, &ni);
				if (rc != HB_E_OVERFLOW || attempt >= 4) break;
				cap *= 4; // (the round left the store untouched: the lists are emitted before sl_ec_r)
			}
			if (rc) return rc;
			B.src.swap(B.t_src); B.rev.swap(B.t_rev); B.src_off.swap(B.t_so); B.rev_off.swap(B.t_ro);
			for (uint64_t i = 0; i < n; i++) { B.fc[i] = B.flags[2 * i]; B.ab[i] = B.flags[2 * i + 1]; if (B.status[i]) res->n_unfinished++; }
		} else if ((rc = ec_round_sharded(ctx, B, (uint64_t)k, is_sv, rank, world, ag, user, &tot_e, &res->n_unfinished, &res->ms_exchange))) return rc;
		res->corrected_bases[k] = tot_e; res->ms_ec[k] = wall_ms() - t;
	}
	t = wall_ms();
	if ((rc = hb_pt_gen(ctx, &hom, &het))) return rc;                               // ha_ec_ff, Assembly.cpp:1942-1959
	ctx->opt.hom_cov = hom; ctx->opt.het_cov = het;
	res->ms_pt[n_round] = wall_ms() - t; t = wall_ms();
	uint64_t r0 = 0, r1 = n; if (world > 1) shard(n, rank, world, &r0, &r1);
	const uint64_t nl = r1 - r0; uint64_t cap = std::max<uint64_t>(4 * (B.src_off[n] + B.rev_off[n]) / world, 128 * nl) + 4096, stat[8];
	B.t_so.assign(nl + 1, 0); B.t_ro.assign(nl + 1, 0);
	for (int attempt = 0;; attempt++) {
		B.t_src.resize(cap); B.t_rev.resize(cap);
		rc = hb_cal_ov_r(ctx, r0, r1, B.src.empty() ? (const hb_ma_hit_t *)B.src_off.data() : B.src.data(), B.src_off.data(), B.rev.empty() ? (const hb_ma_hit_t *)B.rev_off.data() : B.rev.data(), B.rev_off.data(),
		                 B.t_src.data(), B.t_so.data(), cap, B.t_rev.data(), B.t_ro.data(), cap, stat);
		if (rc != HB_E_OVERFLOW || attempt >= 4) break;
		cap *= 4;
	}
	if (world == 1) { if (rc) return rc; B.src.swap(B.t_src); B.rev.swap(B.t_rev); B.src_off.swap(B.t_so); B.rev_off.swap(B.t_ro); }
	else { // the final lists of all reads on every rank
		uint64_t ok = rc == HB_OK ? 1 : 0; std::vector<uint8_t> &blob = B.blob; blob.clear();
		put(blob, &ok, 1); put(blob, &nl, 1);
		if (ok) { put(blob, B.t_so.data(), nl + 1); put(blob, B.t_ro.data(), nl + 1); put(blob, B.t_src.data(), B.t_so[nl]); put(blob, B.t_rev.data(), B.t_ro[nl]); }
		std::vector<uint64_t> sizes; const double tx = wall_ms();
		int grc = gather_blobs(ctx, world, ag, user, blob, B.all, sizes); res->ms_exchange += wall_ms() - tx;
		if (grc) return grc;
		{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) { if (!*(const uint64_t *)p && rc == HB_OK) { hb_set_err(ctx, HB_E_STATE, "the final pass failed on rank %d", r); rc = HB_E_STATE; } p += sizes[r]; } }
		if (rc) return rc;
		B.src_off.assign(n + 1, 0); B.rev_off.assign(n + 1, 0); uint64_t ts = 0, tr = 0;
		{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) { const uint8_t *q = p; take<uint64_t>(q, 1); const uint64_t m = *take<uint64_t>(q, 1); const uint64_t *a = take<uint64_t>(q, m + 1), *b = take<uint64_t>(q, m + 1); ts += a[m]; tr += b[m]; p += sizes[r]; } }
		B.src.resize(ts + 1); B.rev.resize(tr + 1); uint64_t rb = 0, os = 0, orv = 0;
		{ const uint8_t *p = B.all.data(); for (int r = 0; r < world; r++) {
			const uint8_t *q = p; take<uint64_t>(q, 1); const uint64_t m = *take<uint64_t>(q, 1); const uint64_t *a = take<uint64_t>(q, m + 1), *b = take<uint64_t>(q, m + 1);
			const hb_ma_hit_t *xs = take<hb_ma_hit_t>(q, a[m]), *xr = take<hb_ma_hit_t>(q, b[m]);
			for (uint64_t i = 0; i < m; i++) { B.src_off[rb + i] = os + a[i]; B.rev_off[rb + i] = orv + b[i]; }
			if (a[m]) memcpy(B.src.data() + os, xs, a[m] * sizeof(hb_ma_hit_t)); if (b[m]) memcpy(B.rev.data() + orv, xr, b[m] * sizeof(hb_ma_hit_t));
			rb += m; os += a[m]; orv += b[m]; p += sizes[r];
		} }
		if (rb != n) { hb_set_err(ctx, HB_E_STATE, "the shards of the final pass do not add up to the read set"); return HB_E_STATE; }
		B.src_off[n] = os; B.rev_off[n] = orv;
	}
	res->ms_final = wall_ms() - t;
	float dms = 0; cudaEventRecord(e1, ctx->stream); cudaEventSynchronize(e1); cudaEventElapsedTime(&dms, e0, e1); cudaEventDestroy(e0); cudaEventDestroy(e1);
	hb_stage_prof_end(ctx);
	res->device_ms = dms; res->ms_total = wall_ms() - t_begin;
	res->n_reads = n; res->src = B.src.data(); res->rev = B.rev.data(); res->src_off = B.src_off.data(); res->rev_off = B.rev_off.data();
	res->is_fully_corrected = B.fc.data(); res->is_abnormal = B.ab.data(); res->hom_cov = hom; res->het_cov = het; res->n_src = B.src_off[n]; res->n_rev = B.rev_off[n];
	return HB_OK;
}
