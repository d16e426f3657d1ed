// tests/hostemu/hostemu.cpp — TEST INFRASTRUCTURE ONLY.
//
// Compiles the bodies of the product's one-thread-per-unit kernels
// (hifiasm_b200/csrc/*.cuh, the HB_HD functions) as plain host C++ so their
// logic can be checked against the golden vectors in the GPU-less build
// container.  Nothing here is part of, linked into, or called by the product:
// libhifiasm_b200.so has no host path and fails loudly without a CUDA device.
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <math.h>
#include <map>
#include <algorithm>
#include "../../hifiasm_b200/csrc/hb_sketch.cuh"
#include "../../hifiasm_b200/csrc/hb_final.cuh"
#include "../../hifiasm_b200/csrc/hb_ecaln.cuh"
#include "../../hifiasm_b200/csrc/hb_ecrechain.cuh"
#include "../../hifiasm_b200/csrc/hb_bloom.cuh"
#include "../../hifiasm_b200/csrc/hb_ecphase.cuh"
#include "../../hifiasm_b200/csrc/hb_ecround.cuh"
#include "../../hifiasm_b200/csrc/hb_eccns.cuh"
#include "../../hifiasm_b200/csrc/hb_eccns_full.cuh"
#include "seq_ref.h"

struct EmuReads { DevReads d; std::vector<uint8_t> packed; std::vector<uint64_t> off, noff; std::vector<uint32_t> len, npos; };
struct EmuFt { DevFt d; std::vector<uint64_t> key; std::vector<int32_t> val; };
struct EmuPt { DevPt d; std::vector<ulonglong2> slot; std::vector<uint64_t> pos; };

extern "C" {

void *emu_reads_create(uint64_t n, const uint64_t *len, const uint8_t *packed, const uint64_t *byte_off, const uint64_t *n_pos, const uint64_t *n_off)
{
	EmuReads *r = new EmuReads();
	r->off.resize(n + 1); r->len.resize(n); r->noff.resize(n + 1);
	uint64_t o = 0;
	for (uint64_t i = 0; i < n; i++) { r->off[i] = o; r->len[i] = (uint32_t)len[i]; o += ((len[i] / 4 + 1) + 31) & ~31ULL; }
	r->off[n] = o; r->packed.assign(o + 64, 0);
	for (uint64_t i = 0; i < n; i++) memcpy(&r->packed[r->off[i]], packed + byte_off[i], len[i] / 4 + 1);
	for (uint64_t i = 0; i <= n; i++) r->noff[i] = n_off ? n_off[i] : 0;
	r->npos.resize(r->noff[n] + 1);
	for (uint64_t i = 0; i < r->noff[n]; i++) r->npos[i] = (uint32_t)n_pos[i];
	r->d.n = n; r->d.packed = r->packed.data(); r->d.off = r->off.data(); r->d.len = r->len.data(); r->d.noff = r->noff.data(); r->d.npos = r->npos.data();
	return r;
}
void emu_reads_destroy(void *p) { delete (EmuReads *)p; }

void *emu_ft_create(uint64_t n, const uint64_t *key, const int32_t *val)
{
	EmuFt *f = new EmuFt();
	if (n == 0) { f->d.mask = 0; f->d.key = 0; f->d.val = 0; return f; }
	uint64_t cap = 16; while (cap < 2 * n) cap <<= 1;
	f->key.assign(cap, 0); f->val.assign(cap, 0);
	for (uint64_t i = 0; i < n; i++) {
		uint64_t b = hb_bucket(key[i], cap - 1);
		while (f->val[b] != 0) b = (b + 1) & (cap - 1);
		f->key[b] = key[i]; f->val[b] = val[i];
	}
	f->d.mask = cap - 1; f->d.key = f->key.data(); f->d.val = f->val.data();
	return f;
}
void emu_ft_destroy(void *p) { delete (EmuFt *)p; }

void *emu_pt_create(uint64_t nk, const uint64_t *key, const uint64_t *off, const uint32_t *cnt, uint64_t npos, const uint64_t *pos)
{
	EmuPt *t = new EmuPt();
	uint64_t cap = 16; while (cap < 2 * nk) cap <<= 1;
	ulonglong2 z; z.x = 0; z.y = 0;
	t->slot.assign(cap, z);
	for (uint64_t i = 0; i < nk; i++) {
		uint64_t b = hb_bucket(key[i], cap - 1);
		while ((t->slot[b].y & 0xfff) != 0) b = (b + 1) & (cap - 1);
		t->slot[b].x = key[i]; t->slot[b].y = off[i] << 12 | cnt[i];
	}
	t->pos.assign(pos, pos + npos);
	t->d.mask = cap - 1; t->d.slot = t->slot.data(); t->d.pos = t->pos.data();
	return t;
}
void emu_pt_destroy(void *p) { delete (EmuPt *)p; }

int emu_sketch(void *reads, void *ft, int w, int k, int is_hpc, int sample_dist, int rewin, uint64_t rid, uint32_t rid_out,
               hb_mz_t *out, uint32_t cap, uint32_t *n_out)
{
	EmuReads *r = (EmuReads *)reads; EmuFt *f = (EmuFt *)ft;
	SketchPar P = { w, k, is_hpc, sample_dist, rewin };
	std::vector<uint64_t> rx(256), rm(256); std::vector<uint32_t> rl(256), l(cap + 1);
	SketchOut o; o.mz = out; o.l = l.data(); o.cap = cap; o.n = 0; o.ovf = 0;
	RingRef<uint64_t> RX = { rx.data(), 1 }, RM = { rm.data(), 1 }; RingRef<uint32_t> RL = { rl.data(), 1 };
	hb_sketch_read(r->d, f->d, P, rid, rid_out, RX, RM, RL, o);
	*n_out = o.n;
	return o.ovf;
}


// two-stage sketch (event stream + per-event window minimum), driven sequentially
int emu_sketch2(void *reads, void *ft, int w, int k, int is_hpc, int sample_dist, int rewin, uint64_t rid, uint32_t rid_out,
                hb_mz_t *out, uint32_t cap, uint32_t *n_out)
{
	EmuReads *r = (EmuReads *)reads; EmuFt *f = (EmuFt *)ft;
	SketchPar P = { w, k, is_hpc, sample_dist, rewin };
	uint32_t len = r->d.len[rid], T = 0, tl = 0;
	std::vector<ulonglong2> evs(len + 2); std::vector<uint64_t> ex(len + 1), em(len + 1); std::vector<uint32_t> el(len + 1), ol(cap + 1);
	SkEv ev = { evs.data() };
	hb_sketch_events(r->d, f->d, P, rid, ev, &T, &tl);
	{ // what k_sketch_select does when it stages a tile: split the records, l = events since the last N base
		int64_t lastN = -1;
		for (uint32_t t = 0; t < T; t++) {
			ex[t] = evs[t].x; em[t] = evs[t].y;
			if (em[t] == SK_NMARK) { lastN = t; el[t] = 0; em[t] = SK_DUMMY_META; } else el[t] = (uint32_t)(t - lastN);
		}
	}
	std::vector<int32_t> pre(T + 1), suf(T + 1);
	for (uint32_t b = 0; b < T; b += w) {
		uint32_t e = std::min<uint32_t>(T, b + w); int32_t cur = b;
		for (uint32_t u = b; u < e; u++) { if (sk_cmp(ex[u], em[u], ex[cur], em[cur]) <= 0) cur = u; pre[u] = cur; }
		cur = e - 1;
		for (int32_t u = e - 1; u >= (int32_t)b; u--) { if (sk_cmp(ex[u], em[u], ex[cur], em[cur]) < 0) cur = u; suf[u] = cur; }
	}
	SketchOut o; o.mz = out; o.l = ol.data(); o.cap = cap; o.n = 0; o.ovf = 0;
	const uint64_t *X = ex.data(), *M = em.data(); const uint32_t *L = el.data();
	int32_t mp = -1;
	for (int32_t t = 0; t < (int32_t)T; t++) {
		int32_t mc = sk2_window_min(X, M, pre.data(), suf.data(), 0, t, w);
		hb_mz_t tmp[600]; uint32_t tl_[600];
		uint32_t n = sk2_emit(X, M, L, 0, t, mp, mc, w, k, tmp, tl_);
		for (uint32_t i = 0; i < n; i++) o.push(tmp[i].x, tmp[i].info, tl_[i]);
		mp = mc;
	}
	if (mp >= 0 && ex[mp] != ~0ULL) o.push(ex[mp], em[mp], el[mp]);
	if (!o.ovf) {
		if (sample_dist > w && f->d.mask != 0) sk_select_mz_h(o, (int32_t)len, sample_dist, rewin, k, (int32_t)tl);
		for (uint32_t i = 0; i < o.n; i++) out[i].info = (out[i].info & ~0xfffffffULL) | (rid_out & 0xfffffff);
	}
	*n_out = o.n;
	return o.ovf;
}

// weight table of minimizers_qgen0 (anchor.cpp:1066-1075): w_tab[n_occ]
static void weight_tab(uint32_t *w_tab, uint32_t high_occ, uint32_t low_occ)
{
	uint64_t max_cnt = high_occ < 2 ? 2 : high_occ, min_cnt = low_occ < 2 ? 2 : low_occ;
	for (uint32_t n = 0; n < 4096; n++) {
		uint32_t w;
		if (n < max_cnt && n > min_cnt) w = 1;
		else if (n <= min_cnt) w = 2;
		else { w = (uint32_t)(1 + ((n + (max_cnt << 1) - 1) / (max_cnt << 1))); w = (uint32_t)pow((double)w, 1.1); }
		w_tab[n] = w;
	}
}

// sequential stand-in of the probe + group/scatter kernels (anchors.cu): same
// output contract — anchors grouped by key = tid<<1|rev ascending, inside a
// group in query-minimizer order then index-list order.
uint64_t emu_anchors(void *reads, void *pt, const hb_mz_t *mz, uint32_t n_mz, uint32_t high_occ, uint32_t low_occ, hb_hit_t *out, uint64_t cap)
{
	EmuReads *r = (EmuReads *)reads; EmuPt *t = (EmuPt *)pt;
	uint32_t w_tab[4096]; weight_tab(w_tab, high_occ, low_occ);
	std::map<uint32_t, uint64_t> cnt;
	uint64_t tot = 0;
	for (uint32_t i = 0; i < n_mz; i++) {
		uint64_t off; uint32_t n = hb_pt_lookup(t->d, mz[i].x, &off);
		for (uint32_t j = 0; j < n; j++) { uint64_t y = t->d.pos[off + j]; uint32_t key = HB_MZ_RID(y) << 1 | (HB_MZ_REV(mz[i].info) ^ HB_MZ_REV(y)); cnt[key]++; tot++; }
	}
	if (tot > cap) return tot;
	uint64_t base = 0;
	for (auto &kv : cnt) { uint64_t c = kv.second; kv.second = base; base += c; }
	for (uint32_t i = 0; i < n_mz; i++) {
		uint64_t off; uint32_t n = hb_pt_lookup(t->d, mz[i].x, &off);
		uint32_t zpos = HB_MZ_POS(mz[i].info), zspan = HB_MZ_SPAN(mz[i].info);
		for (uint32_t j = 0; j < n; j++) {
			uint64_t y = t->d.pos[off + j]; uint32_t tid = HB_MZ_RID(y), rev = HB_MZ_REV(mz[i].info) ^ HB_MZ_REV(y);
			hb_hit_t &h = out[cnt[tid << 1 | rev]++];
			uint32_t tl = r->d.len[tid];
			h.id_strand = tid | rev << 31; h.self_offset = zpos;
			h.offset = rev ? tl - 1 - (HB_MZ_POS(y) + 1 - HB_MZ_SPAN(y)) : HB_MZ_POS(y);
			h.cnt = w_tab[n] << 8 | (zspan <= 255 ? zspan : 255);
		}
	}
	return tot;
}

static ChainPar chain_par(double bw, int k, int max_n_chain)
{
	ChainPar P; double tmp = expf((float)(-0.01 * (double)k));
	P.pen_gap = 0.5f; P.pen_skip = 0.0005f; P.pen_gap *= tmp; P.pen_skip *= tmp; P.bw_rate = bw;
	P.max_skip = 25; P.max_iter = 5000; P.max_dis = 5000; P.mcopy_num = 3; P.mcopy_khit_cutoff = 32; P.mcopy_rate = 0.7;
	P.max_n_chain = max_n_chain; P.chain_cutoff = 2; P.ocv_w = 3072;
	return P;
}

struct EmuChains { std::vector<hb_chain_t> ch; std::vector<hb_hit_t> chits; std::vector<uint32_t> idx; uint32_t n_ol; std::vector<uint64_t> fc; };

// the chain kernel (thread per group) + post kernel (thread per read), run
// sequentially over one read's grouped anchors
static void run_chains(EmuReads *r, uint32_t rid, hb_hit_t *hits, uint64_t n_hits, const ChainPar &P, EmuChains &E, bool want_fc)
{
	std::vector<GroupDir> dir; uint32_t slot = 0;
	for (uint64_t l = 0, k = 1; k <= n_hits; k++)
		if (k == n_hits || HB_HIT_ID(hits[k]) != HB_HIT_ID(hits[l])) {
			GroupDir g; g.read = rid; g.start = (uint32_t)l; g.count = (uint32_t)(k - l); g.slot = slot;
			if (HB_HIT_ID(hits[l]) != rid) { slot += g.count >= (uint32_t)P.mcopy_khit_cutoff ? P.mcopy_num : 1; dir.push_back(g); }
			else hb_order_group(hits + l, (int32_t)(k - l)); // self hits get no chain slot but are ordered like every group (k_chain_warp, slot == GRP_EMPTY)
			l = k;
		}
	E.ch.assign(slot + 1, hb_chain_t()); E.chits.assign(n_hits + 1, hb_hit_t()); E.idx.assign(slot + 1, 0);
	E.fc.assign(want_fc ? n_hits + 2 * slot + 4 : 1, 0);
	std::vector<int32_t> f(n_hits + 1), p(n_hits + 1), ii(n_hits + 1); std::vector<int64_t> t(n_hits + 1);
	FcOut fc; fc.buf = want_fc ? E.fc.data() : 0; fc.n = 0; fc.cap = (uint32_t)E.fc.size(); fc.ovf = 0;
	for (auto &g : dir) {
		int32_t ns = g.count >= (uint32_t)P.mcopy_khit_cutoff ? P.mcopy_num : 1;
		hb_chain_group(hits + g.start, (int32_t)g.count, E.chits.data() + g.start, g.start, f.data() + g.start, p.data() + g.start, t.data() + g.start, ii.data() + g.start,
		               P, r->d.len[rid], r->d.len[HB_HIT_ID(hits[g.start])], E.ch.data() + g.slot, ns, fc);
	}
	E.fc.resize(fc.n);
	std::vector<uint64_t> cc(r->d.len[rid] / P.ocv_w + 2);
	E.n_ol = hb_chain_post(E.ch.data(), slot, E.chits.data(), hits, E.idx.data(), cc.data(), r->d.len[rid], P);
}

// -> chains in final order (first_hit = compacted cl->list index), compacted chain anchors, fake cigars
int emu_chains(void *reads, uint32_t rid, hb_hit_t *hits, uint64_t n_hits, double bw, int k, int max_n_chain,
               hb_chain_t *out, uint32_t *n_out, hb_hit_t *chits_out, uint64_t *n_chits, uint64_t *fc_out, uint64_t *n_fc)
{
	EmuReads *r = (EmuReads *)reads; ChainPar P = chain_par(bw, k, max_n_chain); EmuChains E;
	run_chains(r, rid, hits, n_hits, P, E, true);
	uint64_t m = 0; uint32_t ord = 0;
	for (size_t s = 0; s + 1 < E.ch.size(); s++) {
		hb_chain_t &c = E.ch[s];
		if (!c.n_hits) continue;
		for (uint32_t h = 0; h < c.n_hits; h++) { chits_out[m] = E.chits[c.first_hit + h]; chits_out[m].id_strand = (chits_out[m].id_strand & 0x80000000u) | ord; m++; }
		ord++;
	}
	*n_chits = m;
	for (uint32_t i = 0; i < E.n_ol; i++) { out[i] = E.ch[E.idx[i]]; out[i].first_hit = out[i].pad; out[i].pad = 0; }
	*n_out = E.n_ol;
	memcpy(fc_out, E.fc.data(), E.fc.size() * 8); *n_fc = E.fc.size();
	return 0;
}

// the whole final pass of one read through the product's device functions
int emu_final_read(void *reads, void *ft, void *pt, int w, int k, int is_hpc, int sample_dist, int rewin, int hom_cov, int max_n_chain, uint32_t rid,
                   hb_ma_hit_t *in0, uint32_t n0, const hb_ma_hit_t *in1, uint32_t n1, hb_ma_hit_t *out0, uint32_t *m0, hb_ma_hit_t *out1, uint32_t *m1)
{
	EmuReads *r = (EmuReads *)reads;
	uint32_t high_occ = (uint32_t)(hom_cov * (2.0 - 0.333)), low_occ = (uint32_t)(hom_cov * 0.333); // ecovlp.cpp:3952-3953
	uint32_t cap = r->d.len[rid] / 4 + 64, n_mz;
	std::vector<hb_mz_t> mz(cap);
	if (emu_sketch(reads, ft, w, k, is_hpc, sample_dist, rewin, rid, 0, mz.data(), cap, &n_mz)) return -1;
	uint64_t na = emu_anchors(reads, pt, mz.data(), n_mz, high_occ, low_occ, 0, 0);
	std::vector<hb_hit_t> hits(na + 1);
	emu_anchors(reads, pt, mz.data(), n_mz, high_occ, low_occ, hits.data(), na);
	ChainPar P = chain_par(0.001, k, max_n_chain); EmuChains E;
	run_chains(r, rid, hits.data(), na, P, E, false);
	std::vector<uint8_t> exact(E.ch.size() + 1, 0);
	for (uint32_t i = 0; i < E.n_ol; i++) {
		const hb_chain_t &c = E.ch[E.idx[i]];
		exact[E.idx[i]] = (uint8_t)hb_exact_seq(r->d, rid, c.x_pos_s, (uint64_t)c.x_pos_e + 1, c.y_id, c.y_pos_s, (uint64_t)c.y_pos_e + 1, (int)c.y_pos_strand);
	}
	std::vector<FinOv> ov(E.n_ol + n0 + 1); std::vector<uint64_t> srt(n0 + n1 + 1);
	unsigned long long stat[8];
	std::vector<int32_t> bb(256), be(256); std::vector<RsFrame> fr(HB_RS_STACK); RsScratch W = { bb.data(), be.data(), fr.data() };
	hb_final_merge(r->d, rid, E.ch.data(), E.idx.data(), E.n_ol, exact.data(), in0, n0, in1, n1, ov.data(), srt.data(), out0, m0, out1, m1, stat, W);
	return 0;
}

// step A of the EC alignment stage for every chain of one read (body of k_ec_overlap).  win[] = the window records of
// the window pass (k_windows layout; rec.chain = chain ordinal, windows of a chain contiguous and in order).
int emu_ec_align_A(void *reads, uint32_t rid, const hb_chain_t *ch, uint32_t n_ch, const uint64_t *fc, const hb_win_t *win, uint32_t n_win, double e_rate, int32_t w_l,
                   hb_aln_t *out, hb_wl_t *wl, uint16_t *pool, uint64_t pool_cap, uint64_t *pool_used)
{
	EmuReads *r = (EmuReads *)reads; int err = 0; unsigned long long used = 0;
	std::vector<uint64_t> path((size_t)w_l * 5); std::vector<uint16_t> ctmp(HB_EC_CIG_TMP);
	EcCtx C; C.R = r->d; C.e_rate = e_rate; C.w_l = w_l; C.pool = pool; C.pool_used = &used; C.pool_cap = pool_cap; C.err = &err;
	C.ez.path = path.data(); C.ez.cig = ctmp.data(); C.ez.cn = 0;
	uint32_t w0 = 0;
	for (uint32_t j = 0; j < n_ch; j++) {
		uint32_t w1 = w0; while (w1 < n_win && win[w1].chain == (int32_t)j) w1++;
		C.q = hb_rd_view(r->d, rid, 0); C.t = hb_rd_view(r->d, ch[j].y_id, ch[j].y_pos_strand);
		hb_aln_t res; res.w_off = w0; res.pad = 0;
		hb_ec_overlap_A(C, ch[j], fc + ch[j].fc_off, win + w0, (int32_t)(w1 - w0), wl + w0, &res);
		out[j] = res; w0 = w1;
	}
	*pool_used = used;
	return err | (used > pool_cap ? 128 : 0);
}

// scratch of the re-seeding rescue (hb_ecrechain.cuh), what k_ecb_rechain gives each of its threads
struct RcScratch {
	std::vector<hb_wl_t> zw; std::vector<uint16_t> zc, cig1; std::vector<uint64_t> path1; std::vector<hb_hit_t> rh; std::vector<int64_t> rt, rp; std::vector<int32_t> rf, bb, be; std::vector<RsFrame> fr;
	unsigned long long zc_used; int rc_err; EcRc S;
	RcScratch(const DevReads &R, const uint16_t *poolA, int32_t w_l, int32_t hcap) : zw(1024), zc((size_t)1 << 16), cig1(HB_EC_CIG_TMP), path1((size_t)w_l * 5), rh(hcap), rt(hcap), rp(hcap), rf(hcap), bb(256), be(256), fr(HB_RS_STACK), zc_used(0), rc_err(0)
	{
		S.R = R; S.zw = zw.data(); S.zcap = (int32_t)zw.size(); S.poolA = poolA; S.zc = zc.data(); S.zc_used = &zc_used; S.zc_cap = zc.size(); S.path1 = path1.data(); S.cig1 = cig1.data();
		S.h = rh.data(); S.hcap = hcap; S.hn = 0; S.t = rt.data(); S.p = rp.data(); S.f = rf.data(); S.rs.bb = bb.data(); S.rs.be = be.data(); S.rs.st = fr.data(); S.err = &rc_err; S.ovf = 0;
		const double tmp = expf((float)(-0.01 * (double)HB_E_KHIT)); S.pen_gap = 0.5f; S.pen_skip = 0.0005f; S.pen_gap *= tmp; S.pen_skip *= tmp; S.h_khit = HB_E_KHIT; // set_lchain_dp_op, anchor.cpp:2272
	}
};

// step B (base-level CIGAR) for the overlaps step A accepted: body of k_ec_cigar.  hits = compacted chain anchors of the read
// (modified in place like return_t_chain does), aln / wlA = step A's output.  path_words / cig_words = per-thread scratch sizes.
// rechain != 0: the re-seeding rescue (hb_ecb_rechain, body of k_ecb_rechain) runs on the overlaps that ask for it; poolA = step A's cigar pool;
// rechain = (hit capacity << 8) | 1 (0 in the upper bits: 8192 hits)
int emu_ec_align_B(void *reads, uint32_t rid, const hb_chain_t *ch, uint32_t n_ch, const uint64_t *fc, hb_hit_t *hits, uint64_t n_hits,
                   const hb_aln_t *aln, const hb_wl_t *wlA, double e_rate, int32_t w_l, int32_t gaps, uint64_t path_words, int32_t cig_words,
                   hb_alnb_t *out, hb_wl_t *wl, uint64_t wl_cap, uint16_t *pool, uint64_t pool_cap, uint64_t *pool_used, uint64_t *n_wl,
                   const uint16_t *poolA, int32_t rechain)
{
	EmuReads *r = (EmuReads *)reads; unsigned long long used = 0; uint64_t nw = 0; int rc = 0;
	RcScratch RS(r->d, poolA, w_l, (rechain >> 8) ? (rechain >> 8) : 8192); EcRc &S = RS.S; int &rc_err = RS.rc_err;
	std::vector<uint64_t> path(path_words), vec(11 * HB_MW_MAXW); std::vector<uint16_t> ecig(cig_words), wc(cig_words);
	std::vector<uint16_t> gout(cig_words);
	EcBCtx C; C.gout = gout.data(); C.gcap = cig_words; C.e_rate = e_rate; C.w_l = w_l; C.pool = pool; C.pool_used = &used; C.pool_cap = pool_cap; C.do_gaps = gaps; C.no_myers = 0;
	C.ez.path = path.data(); C.ez.pcap = path_words; C.ez.vec = vec.data(); C.ez.vstride = HB_MW_MAXW; C.ez.warp = 0; C.ez.cig = ecig.data(); C.ez.ccap = cig_words; C.wc = wc.data(); C.wccap = cig_words;
	for (uint32_t j = 0; j < n_ch; j++) {
		hb_alnb_t res; memset(&res, 0, sizeof(res)); res.st = aln[j].st; res.w_off = nw;
		if (aln[j].st == 2) {
			const hb_chain_t &c = ch[j];
			uint64_t i = c.first_hit; const uint32_t pid = HB_HIT_ID(hits[i]);
			for (; i < n_hits && HB_HIT_ID(hits[i]) == pid && pid != 0x7fffffffu; i++);
			const int64_t scn = (int64_t)(i - c.first_hit);
			std::vector<int64_t> t(scn + 1), p(scn + 1); std::vector<int32_t> f(scn + 1);
			EcZ z; z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_pos_s = c.y_pos_s; z.y_id = c.y_id; z.rev = c.y_pos_strand; z.fc = fc + c.fc_off; z.fc_n = c.fc_n;
			z.align_length = aln[j].align_length; z.w = (hb_wl_t *)(wlA + aln[j].w_off); z.wn = (int32_t)aln[j].w_n;
			C.q = hb_rd_view(r->d, rid, 0); C.t = hb_rd_view(r->d, c.y_id, c.y_pos_strand); C.ql = r->d.len[rid]; C.tl = r->d.len[c.y_id];
			C.aw = wl + nw; C.awcap = (int32_t)std::min<uint64_t>(wl_cap - nw, (uint64_t)scn + 2);
			hb_ec_overlap_B(C, z, aln[j].re, hits + c.first_hit, scn, t.data(), p.data(), f.data(), &res);
			if ((rechain & 1) && res.st == 2 && res.need_rechain) {
				C.awcap = (int32_t)std::min<uint64_t>(wl_cap - nw, (uint64_t)scn + 2 + HB_RC_SPARE_WIN);
				hb_ecb_rechain(C, z, aln[j].re, S, &res);
				if (rc_err) rc |= 2;
			}
			if (res.st < 0) rc |= res.st == -1 ? 1 : 2;
			nw += res.w_n;
		}
		out[j] = res;
	}
	*pool_used = used; *n_wl = nw;
	return rc | (used > pool_cap ? 128 : 0);
}

// the same step as the GPU runs it: prep per overlap, every segment on its own with the tier-0 scratch (640 trace words, 4-word band,
// 72 cigar runs) and, when that is too small, with the large one; then the merge over the stored results with cig_words-sized buffers.
int emu_ec_align_B_par(void *reads, uint32_t rid, const hb_chain_t *ch, uint32_t n_ch, const uint64_t *fc, hb_hit_t *hits, uint64_t n_hits,
                       const hb_aln_t *aln, const hb_wl_t *wlA, double e_rate, int32_t w_l, int32_t gaps, int32_t cig_words,
                       hb_alnb_t *out, hb_wl_t *wl, uint64_t wl_cap, uint16_t *pool, uint64_t pool_cap, uint64_t *pool_used, uint64_t *n_wl, uint64_t *n_tier,
                       const uint16_t *poolA, int32_t rechain)
{
	EmuReads *r = (EmuReads *)reads; unsigned long long used = 0, sused = 0; uint64_t nw = 0; int rc = 0;
	RcScratch RS(r->d, poolA, w_l, 8192); std::vector<uint16_t> rbuf((size_t)3 * 65535);
	std::vector<uint64_t> path0(640), vec0(11 * 4), path1((size_t)1 << 22), vec1(11 * HB_MW_MAXW); std::vector<uint16_t> cig0(72), cig1(65535), spool((size_t)1 << 22);
	std::vector<uint16_t> mbuf((size_t)3 * cig_words);
	n_tier[0] = n_tier[1] = 0;
	for (uint32_t j = 0; j < n_ch; j++) {
		hb_alnb_t res; memset(&res, 0, sizeof(res)); res.st = aln[j].st; res.w_off = nw;
		if (aln[j].st == 2) {
			const hb_chain_t &c = ch[j];
			uint64_t i = c.first_hit; const uint32_t pid = HB_HIT_ID(hits[i]);
			for (; i < n_hits && HB_HIT_ID(hits[i]) == pid && pid != 0x7fffffffu; i++);
			const int64_t scn = (int64_t)(i - c.first_hit);
			std::vector<int64_t> t(scn + 1), p(scn + 1); std::vector<int32_t> f(scn + 1);
			EcZ z; z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_pos_s = c.y_pos_s; z.y_id = c.y_id; z.rev = c.y_pos_strand; z.fc = fc + c.fc_off; z.fc_n = c.fc_n;
			z.align_length = aln[j].align_length; z.w = (hb_wl_t *)(wlA + aln[j].w_off); z.wn = (int32_t)aln[j].w_n;
			EcBCtx C; C.gout = 0; C.gcap = 0; C.e_rate = e_rate; C.w_l = w_l; C.do_gaps = 0; C.no_myers = 0; C.pool = 0; C.pool_used = 0; C.pool_cap = 0; C.aw = 0; C.awcap = 0; C.wc = 0; C.wccap = 0;
			C.q = hb_rd_view(r->d, rid, 0); C.t = hb_rd_view(r->d, c.y_id, c.y_pos_strand); C.ql = r->d.len[rid]; C.tl = r->d.len[c.y_id];
			EcPrep pr; hb_ecb_prep(z, aln[j].re, C.ql, C.tl, hits + c.first_hit, scn, t.data(), p.data(), f.data(), &pr);
			const int32_t ns = (pr.shortcut || pr.ch_n <= 0) ? 0 : pr.ch_n + 1;
			std::vector<EcSeg> segs(ns + 1);
			for (int32_t k = 0; k < ns; k++) {
				int64_t uq[2], ut[2], um;
				{ // the pre-pass: no scratch at all, stops where an alignment would start
					uint16_t one[2]; C.ez.path = 0; C.ez.pcap = 0; C.ez.vec = 0; C.ez.vstride = 0; C.ez.warp = 0; C.ez.cig = one; C.ez.ccap = 2; C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0; C.no_myers = 1;
					const int st = hb_ecb_segment(C, z, hits + c.first_hit, pr.ch_n, k, uq, ut, &um);
					C.no_myers = 0;
					if (st != 5 && !C.ez.ovf) { hb_seg_store(C, st, uq, ut, um, &segs[k], spool.data(), &sused, spool.size()); n_tier[0]++; continue; }
				}
				for (int tier = getenv("HB_EMU_ALN_ALLWARP") ? 1 : 0; tier < 2; tier++) {
					if (tier == 0) { C.ez.path = path0.data(); C.ez.pcap = 640; C.ez.vec = vec0.data(); C.ez.vstride = 4; C.ez.warp = 0; C.ez.cig = cig0.data(); C.ez.ccap = 72; }
					else { C.ez.path = path1.data(); C.ez.pcap = path1.size(); C.ez.vec = vec1.data(); C.ez.vstride = HB_MW_MAXW; C.ez.warp = getenv("HB_EMU_ALN_THREAD") ? 0 : 1; /* the queue tiers run hb_mw_align_w (virtual lanes) */ C.ez.cig = cig1.data(); C.ez.ccap = 65535; }
					C.ez.ovf = 0; C.ez.cn = 0; C.bad = 0;
					const int st = hb_ecb_segment(C, z, hits + c.first_hit, pr.ch_n, k, uq, ut, &um);
					hb_seg_store(C, st, uq, ut, um, &segs[k], spool.data(), &sused, spool.size());
					n_tier[tier]++;
					if (segs[k].status != 4) break;
				}
				if (sused > spool.size()) return 256;
			}
			EcBCtx M; M.gout = 0; M.gcap = 0; M.no_myers = 0; M.e_rate = e_rate; M.w_l = w_l; M.pool = pool; M.pool_used = &used; M.pool_cap = pool_cap; M.do_gaps = gaps;
			M.wc = mbuf.data(); M.wccap = cig_words; M.ez.cig = mbuf.data() + cig_words; M.ez.ccap = cig_words; M.ez.path = (uint64_t *)(mbuf.data() + 2 * (size_t)cig_words); M.ez.pcap = (uint64_t)cig_words / 4;
			M.ez.vec = 0; M.ez.vstride = 0; M.ez.warp = 0; M.q = C.q; M.t = C.t; M.ql = C.ql; M.tl = C.tl;
			M.aw = wl + nw; M.awcap = (int32_t)std::min<uint64_t>(wl_cap - nw, (uint64_t)scn + 2 + HB_RC_SPARE_WIN);
			hb_ecb_merge(M, z, aln[j].re, pr, segs.data(), spool.data(), &res);
			if (rechain && res.st == 2 && res.need_rechain) { // what k_ecb_rechain does: a fresh context with the large aligner scratch over the merged list
				EcBCtx Rc; Rc.gout = rbuf.data() + 2 * 65535; Rc.gcap = 65535; Rc.e_rate = e_rate; Rc.w_l = w_l; Rc.pool = pool; Rc.pool_used = &used; Rc.pool_cap = pool_cap; Rc.do_gaps = gaps; Rc.no_myers = 0;
				Rc.q = C.q; Rc.t = C.t; Rc.ql = C.ql; Rc.tl = C.tl; Rc.aw = M.aw; Rc.awcap = M.awcap;
				Rc.ez.path = path1.data(); Rc.ez.pcap = path1.size(); Rc.ez.vec = vec1.data(); Rc.ez.vstride = HB_MW_MAXW; Rc.ez.warp = 0; Rc.ez.cig = rbuf.data(); Rc.ez.ccap = 65535;
				Rc.wc = rbuf.data() + 65535; Rc.wccap = 65535;
				hb_ecb_rechain(Rc, z, aln[j].re, RS.S, &res);
				if (RS.rc_err) rc |= 2;
			}
			if (res.st < 0) rc |= res.st == -1 ? 1 : 2;
			nw += res.w_n;
		}
		out[j] = res;
	}
	*pool_used = used; *n_wl = nw;
	return rc | (used > pool_cap ? 128 : 0);
}

// phasing (row a13) of the accepted overlaps of one read: bodies of k_ph_count / k_ph_decide.  alnb / wl / pool = output of steps B + C,
// aln = step A (align_length).  is_match / strong are written per ACCEPTED overlap, in list order.
int emu_ec_phase(void *reads, uint32_t rid, const hb_chain_t *ch, uint32_t n_ch, const hb_aln_t *aln, const hb_alnb_t *alnb, const hb_wl_t *wl, const uint16_t *pool,
                 uint8_t *is_match, int8_t *strong, uint32_t *n_acc)
{
	EmuReads *r = (EmuReads *)reads; std::vector<PhOv> ov; int ovf = 0;
	for (uint32_t j = 0; j < n_ch; j++) {
		if (alnb[j].st != 2) continue;
		PhOv o; o.w = wl + alnb[j].w_off; o.wn = alnb[j].w_n; o.pool = pool; o.y_id = ch[j].y_id; o.rev = ch[j].y_pos_strand; o.align_length = aln[j].align_length; o.is_match = 1; o.strong = 0;
		ov.push_back(o);
	}
	const int64_t ql = r->d.len[rid]; std::vector<uint8_t> cnt(hb_ph_bits_bytes((uint64_t)ql), 0); uint32_t ns = 0, ne = 0;
	hb_ph_count_w(ov.data(), (uint32_t)ov.size(), cnt.data(), ql, &ns, &ne);
	std::vector<uint32_t> site_pos(ns + 1), site_off(ns + 2), ov_off(ov.size() + 2); std::vector<PhEv> ev(ne + 1), ev2(ne + 1); std::vector<PhSnp> snp(4 * (size_t)ns + 1); std::vector<uint64_t> ord(ov.size() + 1);
	std::vector<int32_t> bb(256), be(256); std::vector<RsFrame> fr(HB_RS_STACK); RsScratch W = { bb.data(), be.data(), fr.data() };
	hb_ph_decide_w(r->d, rid, ov.data(), (uint32_t)ov.size(), cnt.data(), ql, ns, ne, site_pos.data(), site_off.data(), ev.data(), ev2.data(), snp.data(), ord.data(), ov_off.data(), W, 3, 3, 0.04, &ovf);
	for (size_t k = 0; k < ov.size(); k++) { is_match[k] = ov[k].is_match; strong[k] = ov[k].strong; }
	*n_acc = (uint32_t)ov.size();
	return ovf;
}

// the round's reverse_paf[i] from the phased overlaps (body of k_ec_rpaf)
int emu_ec_reverse(void *reads, uint32_t rid, const hb_phase_t *ph, uint32_t n, hb_ma_hit_t *out, uint32_t *n_out)
{
	EmuReads *r = (EmuReads *)reads; int ovf = 0; std::vector<PhPair> ord(n + 1);
	std::vector<int32_t> bb(256), be(256); std::vector<RsFrame> fr(HB_RS_STACK); RsScratch W = { bb.data(), be.data(), fr.data() };
	*n_out = hb_ec_reverse_list(r->d, rid, ph, n, ord.data(), W, out, &ovf);
	return ovf;
}

// ---- closing steps of an EC round (hb_ecround.cuh) ----
// a16: a new read store = every read with its edit script applied (bodies of k_sl_len / k_sl_apply)
void *emu_ec_apply(void *reads, const uint16_t *sc, const uint64_t *sc_off)
{
	EmuReads *r = (EmuReads *)reads, *o = new EmuReads(); const uint64_t n = r->d.n;
	o->off.resize(n + 1); o->len.resize(n); o->noff.assign(n + 1, 0);
	std::vector<uint32_t> changed(n);
	uint64_t off = 0;
	for (uint64_t i = 0; i < n; i++) {
		hb_sl_len(sc + sc_off[i], (uint32_t)(sc_off[i + 1] - sc_off[i]), r->d.len[i], &o->len[i], &changed[i]);
		o->off[i] = off; off += ((o->len[i] / 4 + 1) + 31) & ~31ULL;
	}
	o->off[n] = off; o->packed.assign(off + 64, 0);
	std::vector<uint32_t> tmp(r->d.noff[n] + 1);
	for (uint64_t i = 0; i < n; i++) {
		uint32_t nn;
		if (changed[i]) nn = hb_sl_apply(hb_rd_view(r->d, i, 0), sc + sc_off[i], (uint32_t)(sc_off[i + 1] - sc_off[i]), &o->packed[o->off[i]], tmp.data() + r->d.noff[i]);
		else { memcpy(&o->packed[o->off[i]], r->d.packed + r->d.off[i], r->d.len[i] / 4 + 1); nn = (uint32_t)(r->d.noff[i + 1] - r->d.noff[i]); for (uint32_t k = 0; k < nn; k++) tmp[r->d.noff[i] + k] = r->d.npos[r->d.noff[i] + k]; }
		o->noff[i + 1] = o->noff[i] + nn;
	}
	o->npos.resize(o->noff[n] + 1);
	for (uint64_t i = 0; i < n; i++) for (uint64_t k = 0; k < o->noff[i + 1] - o->noff[i]; k++) o->npos[o->noff[i] + k] = tmp[r->d.noff[i] + k];
	o->d.n = n; o->d.packed = o->packed.data(); o->d.off = o->off.data(); o->d.len = o->len.data(); o->d.noff = o->noff.data(); o->d.npos = o->npos.data();
	return o;
}
// a18: reverse complement of every read (body of k_rc_reads)
void *emu_ec_rc(void *reads)
{
	EmuReads *r = (EmuReads *)reads, *o = new EmuReads(); const uint64_t n = r->d.n;
	o->off = r->off; o->len = r->len; o->noff = r->noff; o->npos.resize(r->npos.size()); o->packed.assign(r->packed.size(), 0);
	for (uint64_t i = 0; i < n; i++) {
		const RdView v = hb_rd_view(r->d, i, 0);
		for (uint32_t j = 0; j < v.len / 4 + 1; j++) o->packed[o->off[i] + j] = hb_rc_byte(v, j);
		for (uint32_t k = 0; k < v.nn; k++) o->npos[r->d.noff[i] + k] = v.len - 1 - v.npos[v.nn - 1 - k];
	}
	o->d.n = n; o->d.packed = o->packed.data(); o->d.off = o->off.data(); o->d.len = o->len.data(); o->d.noff = o->noff.data(); o->d.npos = o->npos.data();
	return o;
}
// copy of a read store in the All_reads layout (len/4+1 bytes per read)
void emu_reads_export(void *reads, uint64_t *len, uint8_t *packed, uint64_t *n_off, uint64_t *n_pos)
{
	EmuReads *r = (EmuReads *)reads; uint64_t o = 0;
	for (uint64_t i = 0; i < r->d.n; i++) { if (len) len[i] = r->d.len[i]; if (packed) memcpy(packed + o, r->d.packed + r->d.off[i], r->d.len[i] / 4 + 1); o += r->d.len[i] / 4 + 1; }
	if (n_off) for (uint64_t i = 0; i <= r->d.n; i++) n_off[i] = r->d.noff[i];
	if (n_pos) for (uint64_t i = 0; i < r->d.noff[r->d.n]; i++) n_pos[i] = r->d.npos[i];
}
uint64_t emu_reads_n_npos(void *reads) { EmuReads *r = (EmuReads *)reads; return r->d.noff[r->d.n]; }
// a17: every record of every paf[i] against the corrected read store (body of k_update_dc); returns the number of exact records
uint64_t emu_ec_update(void *reads, hb_ma_hit_t *paf, const uint64_t *off, const uint16_t *sc, const uint64_t *sc_off)
{
	EmuReads *r = (EmuReads *)reads; uint64_t ne = 0;
	for (uint64_t i = 0; i < r->d.n; i++) for (uint64_t k = off[i]; k < off[i + 1]; k++) ne += hb_update_dc(r->d, i, &paf[k], sc, sc_off);
	return ne;
}
// a18: flip_paf_rc of one list; n_out[i] = new length of list i (records stay at off[i])
void emu_ec_flip(void *reads, hb_ma_hit_t *paf, const uint64_t *off, uint32_t *n_out)
{
	EmuReads *r = (EmuReads *)reads;
	for (uint64_t i = 0; i < r->d.n; i++) n_out[i] = hb_flip_paf(r->d.len, i, paf + off[i], (uint32_t)(off[i + 1] - off[i]));
}

// a15: the round's paf[i] + is_fully_corrected / is_abnormal from the phased overlaps, the step-C window lists and the read's edit script
// (bodies of k_ec_spaf).  ph / alnb: one entry per overlap of the read, parallel arrays.
int emu_ec_source(void *reads, uint32_t rid, const hb_phase_t *ph, const hb_alnb_t *alnb, uint32_t n, const hb_wl_t *wl, const uint16_t *pool, const uint16_t *ec, uint64_t ecn,
                  hb_ma_hit_t *out, uint32_t *n_out, uint8_t *flags)
{
	EmuReads *r = (EmuReads *)reads; int ovf = 0; std::vector<PhPair> ord(n + 1);
	std::vector<int32_t> bb(256), be(256); std::vector<RsFrame> fr(HB_RS_STACK); RsScratch W = { bb.data(), be.data(), fr.data() };
	const uint32_t keep = hb_ec_dedup(ph, n, ord.data(), W, &ovf);
	if (ovf) return ovf;
	*n_out = hb_ec_source_list(r->d, rid, ph, alnb, ord.data(), keep, wl, pool, ec, (int64_t)ecn, out);
	std::vector<uint64_t> srt(2 * (size_t)*n_out + 2);
	hb_check_well_cal(ec, (uint32_t)ecn, srt.data(), out, *n_out, r->d.len[rid], 6 /* MIN_COVERAGE_THRESHOLD * 2, ecovlp.cpp:3324 */, &flags[0], &flags[1]);
	return 0;
}

// a14: the read's edit script by window consensus (body of k_ec_cns) from the phased overlaps and the step-C window lists.
// Returns 0 = script written, 1 = the read needs the graph consensus and no arena was given (g_nodes == 0), 2 = output capacity, 3 = graph arena too small, < 0 = dedup overflow
int emu_ec_cns(void *reads, uint32_t rid, const hb_phase_t *ph, const hb_alnb_t *alnb, uint32_t n, const hb_wl_t *wl, const uint16_t *pool, uint16_t *out, uint32_t out_cap, uint32_t *n_out, uint64_t *nec,
               uint32_t g_nodes, uint32_t g_arcs)
{
	EmuReads *r = (EmuReads *)reads; int ovf = 0; std::vector<PhPair> ord(n + 1);
	std::vector<int32_t> bb(256), be(256); std::vector<RsFrame> fr(HB_RS_STACK); RsScratch W = { bb.data(), be.data(), fr.data() };
	const uint32_t keep = hb_ec_dedup(ph, n, ord.data(), W, &ovf);
	if (ovf) return -1;
	std::vector<CnsOv> ov; uint32_t n_ent = 0;
	for (uint32_t k = 0; k < keep; k++) {
		const hb_phase_t &z = ph[ord[k].idx]; const hb_alnb_t &b = alnb[ord[k].idx];
		if (z.is_match != 1 || !b.w_n) continue;
		CnsOv o; o.w = wl + b.w_off; o.wn = b.w_n; o.y_id = z.y_id; o.rev = z.rev; ov.push_back(o); n_ent += b.w_n;
	}
	std::vector<CnsEnt> ent(n_ent + 1); std::vector<uint32_t> srt(n_ent + 1), aa(n_ent + 1), ab(n_ent + 1), b32(n_ent + 1); std::vector<uint64_t> key(n_ent + 1), ct(2 * HB_CNS_WL);
	CnsCtx C; C.R = r->d; C.q = hb_rd_view(r->d, rid, 0); C.ql = r->d.len[rid]; C.ov = ov.data(); C.pool = pool; C.ent = ent.data(); C.ct = ct.data(); C.b32 = b32.data();
	C.out = out; C.out_cap = out_cap; C.g = 0;
	// with_graph: the arena of the graph consensus (what the second launch of the GPU path owns per thread); sizes are parameters so that the overflow report can be tested
	std::vector<CnsNode> nd(g_nodes + 1); std::vector<CnsArc> arcs(g_arcs + 1); std::vector<uint32_t> gq(g_nodes + 1), gb32(g_arcs + 1), gnp(4100); std::vector<uint8_t> gns(4100);
	std::vector<uint64_t> path(32768), vec(11 * 2 + 4); std::vector<uint16_t> gcig(4096);
	CnsG G; memset(&G, 0, sizeof(G));
	if (g_nodes) {
		G.nd = nd.data(); G.ncap = g_nodes; G.arc = arcs.data(); G.arc_cap = g_arcs; G.q = gq.data(); G.q_cap = g_nodes; G.b32 = gb32.data(); G.b32_cap = g_arcs;
		G.nseq = gns.data(); G.nseq_np = gnp.data(); G.nseq_cap = 4096; G.ez.path = path.data(); G.ez.pcap = path.size(); G.ez.vec = vec.data(); G.ez.vstride = 2; G.ez.warp = 0; G.ez.cig = gcig.data(); G.ez.ccap = (int32_t)gcig.size();
		C.g = &G;
	}
	std::vector<uint64_t> S(HB_CNS_SMEM_WORDS); // the warp's shared-memory words (a warp of one lane here)
	if (getenv("HB_EMU_CNS_THREAD")) *nec = g_nodes ? hb_cns_read<true>(C, (uint32_t)ov.size(), srt.data(), aa.data(), ab.data(), key.data()) : hb_cns_read<false>(C, (uint32_t)ov.size(), srt.data(), aa.data(), ab.data(), key.data());
	else *nec = g_nodes ? hb_cns_read_w<true>(C, S.data(), (uint32_t)ov.size(), srt.data(), aa.data(), ab.data(), key.data()) : hb_cns_read_w<false>(C, S.data(), (uint32_t)ov.size(), srt.data(), aa.data(), ab.data(), key.data());
	*n_out = C.out_n;
	return C.need_full ? (C.need_full == 2 ? 3 : 1) : (C.ovf ? 2 : 0);
}


// one call of the step-B aligner on two reads of the store: which = 0 hb_mw_align (thread), 1 hb_mw_align_w (band over virtual lanes).
// pattern = target[ps0, ps0 + pn) on strand rev, text = query[qs0, qs0 + tn).  -> res[6] = err, ps, pe, ts, te, cn; returns ez.ovf
int emu_mw_align(void *reads, uint32_t qid, uint32_t tid, uint32_t rev, int mode, int64_t ps0, int32_t pn, int64_t qs0, int32_t tn, int32_t thre, int32_t abs_diag, int which,
                 int32_t *res, uint16_t *cig, int32_t ccap)
{
	EmuReads *r = (EmuReads *)reads;
	const RdView Q = hb_rd_view(r->d, qid, 0), T = hb_rd_view(r->d, tid, rev);
	std::vector<uint64_t> path((size_t)HB_MW_MAXW * 5 * (tn + 2) + 256), vec(11 * HB_MW_MAXW);
	MwEz ez; memset(&ez, 0, sizeof(ez)); ez.path = path.data(); ez.pcap = path.size(); ez.vec = vec.data(); ez.vstride = HB_MW_MAXW; ez.cig = cig; ez.ccap = ccap; ez.warp = which;
	if (which == 0) hb_mw_align(mode, T, ps0, pn, Q, qs0, tn, thre, abs_diag, ez);
	else if ((((thre << 1) + 1 + 63) >> 6) <= 32) hb_mw_align_w<1>(mode, T, ps0, pn, Q, qs0, tn, thre, abs_diag, ez);
	else hb_mw_align_w<2>(mode, T, ps0, pn, Q, qs0, tn, thre, abs_diag, ez);
	res[0] = ez.err; res[1] = ez.ps; res[2] = ez.pe; res[3] = ez.ts; res[4] = ez.te; res[5] = ez.err <= ez.thre ? ez.cn : 0;
	return ez.ovf;
}

// counting behind the reference's Bloom filter the way index.cu evaluates it (hb_bloom.cuh): distinct k-mers with the ordinal of their first
// occurrence, grouped by filter block, walked in first-occurrence order.  hashes = every k-mer of the store in (read, position) order.
// -> (key, count) of the k-mers that entered the table, by key.
uint64_t emu_bf_counts(const uint64_t *hashes, uint64_t n, int bf_shift, uint64_t *out_key, uint32_t *out_cnt, uint64_t cap)
{
	const int bf_local = bf_shift - 12; const bool use_bf = hb_bf_active(bf_shift);
	std::vector<std::pair<uint64_t, uint64_t>> a(n); // (hash, ordinal): the stable pair sort
	for (uint64_t i = 0; i < n; i++) a[i] = std::make_pair(hashes[i], i);
	std::sort(a.begin(), a.end());
	struct Run { uint64_t key, first, len; uint8_t fp; };
	std::vector<Run> runs;
	for (uint64_t i = 0, j; i < n; i = j) { for (j = i + 1; j < n && a[j].first == a[i].first; j++) {} Run r = { a[i].first, a[i].second, j - i, 0 }; runs.push_back(r); }
	if (use_bf) {
		std::vector<std::pair<std::pair<uint64_t, uint64_t>, uint64_t>> o(runs.size()); // ((block, first ordinal), run)
		for (uint64_t j = 0; j < runs.size(); j++) o[j] = std::make_pair(std::make_pair(hb_bf_block(runs[j].key, bf_local), runs[j].first), j);
		std::sort(o.begin(), o.end());
		for (uint64_t i = 0, j; i < o.size(); i = j) {
			uint64_t st[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
			for (j = i; j < o.size() && o[j].first.first == o[i].first.first; j++) runs[o[j].second].fp = hb_bf_insert(st, runs[o[j].second].key, bf_local) == 4;
		}
	}
	uint64_t m = 0;
	for (auto &r : runs) {
		const uint32_t c = use_bf ? hb_bf_count(r.len, r.fp) : (uint32_t)(r.len > 4095 ? 4095 : r.len);
		if (!c) continue;
		if (m < cap) { out_key[m] = r.key; out_cnt[m] = c; }
		m++;
	}
	return m;
}

} // extern "C"
