// seq_ref.h — one-thread restatements kept ONLY for tests/hostemu: the straight-line forms of the chaining DP (Hash_Table.cpp:2124-2176) and of the
// one-pass sketch (sketch.cpp:454-579) and the one-thread window consensus (wcns_vote, ecovlp.cpp:2185) that the product's warp-parallel chain kernel,
// two-stage sketch and warp-per-read consensus are checked against on the golden vectors.
// Not part of the product: nothing under hifiasm_b200/ includes this file.
#pragma once
#include "../../hifiasm_b200/csrc/hb_chain.cuh"
#include "../../hifiasm_b200/csrc/hb_sketch.cuh"
#include "../../hifiasm_b200/csrc/hb_eccns.cuh"
#include "../../hifiasm_b200/csrc/hb_eccns_full.cuh"

// the chaining DP over [si,ei), Hash_Table.cpp:2124-2176 (one thread)
HB_HD void hb_chain_dp(const hb_hit_t *a, int32_t a_n, int32_t *f, int32_t *p, int64_t *t, int32_t *ii, const ChainPar &P, int64_t xl, int64_t yl, ChainState &S)
{
	int64_t max_f, n_skip, st, max_j, end_j, sc, msc = S.msc, msc_i = S.msc_i, max_ii, ovl, movl = S.movl, plus = S.plus, si = S.si, ei = S.ei, i, j;
	int32_t max, tmp;
	(void)a_n;
	for (i = st = si, max_ii = -1; i < ei; ++i) {
		max_f = a[i].cnt & 0xffu;
		n_skip = 0; max_j = end_j = -1;
		if (i - st > P.max_iter) st = i - P.max_iter;
		while (HB_HIT_ST(a[i]) != HB_HIT_ST(a[st])) ++st;
		for (j = i - 1; j >= st; --j) {
			int32_t s = hb_link_sc(a[i], a[j], P, xl, yl, 0);
			if (s == HB_LINK_FAIL) continue;
			sc = (int64_t)s + f[j];
			if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
			else if (t[j] == (int32_t)i) { if (++n_skip > P.max_skip) break; }
			if (p[j] >= 0) t[p[j]] = i;
		}
		end_j = j;
		if (max_ii < 0 || (int64_t)a[i].self_offset > (int64_t)a[max_ii].self_offset + P.max_dis || HB_HIT_ST(a[i]) != HB_HIT_ST(a[max_ii])) {
			max = INT32_MIN; max_ii = -1;
			for (j = i - 1; j >= st && (int64_t)a[i].self_offset <= (int64_t)P.max_dis + (int64_t)a[j].self_offset && HB_HIT_ST(a[i]) == HB_HIT_ST(a[j]); --j)
				if (max < f[j]) { max = f[j]; max_ii = j; }
		}
		if (max_ii >= 0 && max_ii < end_j && HB_HIT_ST(a[i]) == HB_HIT_ST(a[max_ii])) {
			tmp = hb_link_sc(a[i], a[max_ii], P, xl, yl, 0);
			if (tmp != HB_LINK_FAIL && max_f < (int64_t)tmp + f[max_ii]) { max_f = (int64_t)tmp + f[max_ii]; max_j = max_ii; }
		}
		f[i] = (int32_t)max_f; p[i] = (int32_t)max_j;
		if (max_ii < 0 || ((int64_t)a[i].self_offset <= (int64_t)P.max_dis + (int64_t)a[max_ii].self_offset && HB_HIT_ST(a[i]) == HB_HIT_ST(a[max_ii]) && f[max_ii] < f[i])) max_ii = i;
		if (f[i] >= msc) {
			ovl = hb_chain_len(a[i].self_offset, a[i].self_offset, xl, a[i].offset, a[i].offset, yl);
			if (f[i] > msc || ovl < movl) { msc = f[i]; msc_i = i; movl = ovl; }
		}
		if (f[i] < plus) plus = f[i];
		ii[i] = 0;
	}
	S.plus = plus; S.msc = msc; S.msc_i = msc_i; S.movl = movl;
}

// Chain one target group (one thread): order + quick check + DP + finish.
// a[0..a_n): the group's anchors (same target id, both strands), ordered here by
// (strand, self_offset, offset) — the order minimizers_qgen0's sort produces
// (anchor.cpp:1046-1049).  des: a_n-sized output slice for the chain anchors.
// f,p,ii (int32) and t (int64): a_n-sized scratch.  out: the group's chain slots
// (n_slots = a_n>=mcopy_khit_cutoff ? mcopy_num : 1).  Returns the number of
// chain anchors written to des.
HB_HD int32_t hb_chain_group(hb_hit_t *a, int32_t a_n, hb_hit_t *des, uint64_t des_abs, int32_t *f, int32_t *p, int64_t *t, int32_t *ii,
                             const ChainPar &P, int64_t xl, int64_t yl, hb_chain_t *out, int32_t n_slots, FcOut &fc)
{
	ChainState S;
	for (int32_t i = 0; i < n_slots; i++) out[i].n_hits = 0;
	if (a_n <= 0) return 0;
	hb_order_group(a, a_n);
	hb_chain_quick(a, a_n, f, p, t, ii, P, xl, yl, S);
	hb_chain_dp(a, a_n, f, p, t, ii, P, xl, yl, S);
	return hb_chain_finish(a, a_n, des, des_abs, f, p, t, ii, P, xl, yl, out, n_slots, fc, S);
}

// Sketch read `rid`; rx/rm/rl = this thread's candidate ring (w slots).
// rid_out is written into ha_mz1_t.rid of the results (sketch.cpp:577).
HB_HD void hb_sketch_read(const DevReads &R, const DevFt &ft, const SketchPar &P, uint64_t rid, uint32_t rid_out,
                          RingRef<uint64_t> rx, RingRef<uint64_t> rm, RingRef<uint32_t> rl, SketchOut &o)
{
	const int32_t w = P.w, k = P.k, len = (int32_t)R.len[rid];
	const uint64_t shift1 = k - 1, mask = (1ULL << k) - 1;
	const uint8_t *seq = R.packed + R.off[rid];
	const uint64_t *seq64 = (const uint64_t *)seq;
	uint64_t pl0 = 0, pl1 = 0, pl2 = 0, pl3 = 0, minx = ~0ULL, minm = SK_DUMMY_META, word = 0;
	int32_t i, j, l = 0, tl = 0, bp = 0, min_bp = 0, span = 0, qf = 0, qc = 0;
	uint32_t min_l = 0xffffffffu;
	uint64_t ni = R.noff ? R.noff[rid] : 0, ne = R.noff ? R.noff[rid + 1] : 0;
	int32_t next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX;
	uint8_t q[64]; // run lengths of the last k HPC symbols (tiny_queue_t, htab.h:39-57); capped at 255: spans >= 256 are never candidates
	o.n = 0; o.ovf = 0;
	for (j = 0; j < w; j++) { rx[j] = ~0ULL; rm[j] = ~0ULL; rl[j] = 0; } // memset 0xff, sketch.cpp:470
	int32_t wbase = -1; // index of the 64-bit word (32 bases) held in `word`
#define SK_BASE(ii) ((int)(((((ii) >> 5) != wbase ? (wbase = (ii) >> 5, word = seq64[wbase]) : word) >> ((((ii) & 31) >> 2 << 3) + ((3 - ((ii) & 3)) << 1))) & 3))
	for (i = 0; i < len; ++i) {
		int c = SK_BASE(i);
		uint64_t ix = ~0ULL, im = SK_DUMMY_META;
		if (i == next_n) { c = 4; ++ni; next_n = ni < ne ? (int32_t)R.npos[ni] : INT32_MAX; }
		if (c < 4) {
			int z;
			if (P.is_hpc) { // sketch.cpp:480-492
				int32_t run = 1;
				while (i + run < len && i + run != next_n && SK_BASE(i + run) == c) ++run;
				i += run - 1;
				q[(qc++ + qf) & 0x3f] = (uint8_t)(run > 255 ? 255 : run);
				span += run > 255 ? 255 : run;
				if (qc > k) { span -= q[qf++]; qf &= 0x3f; --qc; }
			} else span = l + 1 < k ? l + 1 : k;
			pl0 = (pl0 << 1 | (uint64_t)(c & 1)) & mask;
			pl1 = (pl1 << 1 | (uint64_t)(c >> 1)) & mask;
			pl2 = pl2 >> 1 | (uint64_t)(1 - (c & 1)) << shift1;
			pl3 = pl3 >> 1 | (uint64_t)(1 - (c >> 1)) << shift1;
			if (pl1 == pl3) continue; // sketch.cpp:502
			z = pl1 < pl3 ? 0 : 1;
			++l; ++tl;
			if (l >= k && span < 256) {
				uint64_t y = z ? hb_hash64(pl2) + hb_hash64(pl3) : hb_hash64(pl0) + hb_hash64(pl1);
				int32_t cnt = hb_ft_lookup(ft, y);
				if (!(cnt >= 1 << 28)) { ix = y; im = (uint64_t)(uint32_t)cnt | (uint64_t)i << 28 | (uint64_t)z << 55 | (uint64_t)span << 56; }
			}
		} else { l = 0; qc = qf = 0; span = 0; }
		rx[bp] = ix; rm[bp] = im; rl[bp] = (uint32_t)l;
#define SK_POS(m) ((uint32_t)(((m) >> 28) & 0x7ffffffULL))
		if (l == w + k - 1 && minx != ~0ULL) { // sketch.cpp:523-534
			for (j = bp + 1; j < w; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) == 0 && SK_POS(bm) != SK_POS(minm)) o.push(bx, bm, rl[j]); }
			for (j = 0; j < bp; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) == 0 && SK_POS(bm) != SK_POS(minm)) o.push(bx, bm, rl[j]); }
		}
		if (sk_cmp(minx, minm, ix, im) >= 0) { // sketch.cpp:543-547
			if (l >= w + k && minx != ~0ULL) o.push(minx, minm, min_l);
			minx = ix; minm = im; min_bp = bp; min_l = (uint32_t)l;
		} else if (bp == min_bp) { // sketch.cpp:548-568
			if (l >= w + k - 1 && minx != ~0ULL) o.push(minx, minm, min_l);
			minx = ~0ULL; minm = SK_DUMMY_META;
			for (j = bp + 1; j < w; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) >= 0) { minx = bx; minm = bm; min_bp = j; min_l = rl[j]; } }
			for (j = 0; j <= bp; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) >= 0) { minx = bx; minm = bm; min_bp = j; min_l = rl[j]; } }
			if (l >= w + k - 1 && minx != ~0ULL) {
				for (j = bp + 1; j < w; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) == 0 && SK_POS(minm) != SK_POS(bm)) o.push(bx, bm, rl[j]); }
				for (j = 0; j <= bp; ++j) { uint64_t bx = rx[j], bm = rm[j]; if (sk_cmp(minx, minm, bx, bm) == 0 && SK_POS(minm) != SK_POS(bm)) o.push(bx, bm, rl[j]); }
			}
		}
		if (++bp == w) bp = 0;
	}
	if (minx != ~0ULL) o.push(minx, minm, min_l);
	if (o.ovf) return;
	if (P.sample_dist > w && ft.mask != 0) sk_select_mz_h(o, len, P.sample_dist, P.rewin, k, tl); // sketch.cpp:575 (a no-op without filter-table hits)
	for (i = 0; i < (int32_t)o.n; ++i) o.mz[i].info = (o.mz[i].info & ~0xfffffffULL) | (rid_out & 0xfffffff);
#undef SK_BASE
#undef SK_POS
}

// ---- the window consensus of one read by one thread (per-column counts in a plain array), the form round 1 shipped; the product runs hb_cns_read_w ----
// extract_sub_cigar_mm, ecovlp.cpp:283-360: votes of one window alignment over [s, e); ct[2 k] = column k, ct[2 k + 1] = the gap in front of it,
// each word = matches << 32 | voters.  (k counts from the clamped s: the caller offsets ct — see hb_cns_vote.)
HB_HD void hb_cns_sub_mm(CnsCtx &C, CnsEnt &p, int64_t s, int64_t e, uint64_t *ct)
{
	const hb_wl_t &w = C.ov[p.ov].w[p.wid];
	int64_t xk = p.xoff, yk = p.yoff, ck = p.coff, os, oe, t;
	const int64_t s0 = w.x_start, e0 = (int64_t)w.x_end + 1;
	if (s < s0) s = s0; if (e > e0) e = e0;
	if (s >= e) return;
	const uint16_t *cg = C.pool + w.cidx; const int64_t cn = w.clen;
	if (!cn) return;
	int64_t op, ws, we, ovlp;
	if (ck < 0 || ck > cn) { ck = 0; xk = w.x_start; yk = w.y_start; }
	while (ck > 0 && xk >= s) { --ck; op = cg[ck] >> 14; if (op != 2) xk -= cg[ck] & 0x3fff; if (op != 3) yk -= cg[ck] & 0x3fff; }
	while (ck < cn && xk < e) {
		ws = xk; op = cg[ck] >> 14;
		if (op != 2) xk += cg[ck] & 0x3fff; if (op != 3) yk += cg[ck] & 0x3fff;
		ck++; we = xk;
		os = s > ws ? s : ws; oe = e < we ? e : we; ovlp = oe > os ? oe - os : 0;
		if (op != 2) { if (!ovlp) continue; } else { if (ws < s || ws >= e) continue; }
		if (op == 0) {
			for (t = os + 1; t < oe; t++) { ct[(t - s) << 1] += 0x100000001ULL; ct[((t - s) << 1) + 1] += 0x100000001ULL; }
			t = os;
			if (t < oe) { ct[(t - s) << 1] += 0x100000001ULL; if (os > ws) ct[((t - s) << 1) + 1] += 0x100000001ULL; }
		} else if (op != 2) {
			for (t = os + 1; t < oe; t++) { ct[(t - s) << 1]++; ct[((t - s) << 1) + 1]++; }
			t = os;
			if (t < oe) { ct[(t - s) << 1]++; if (os > ws) ct[((t - s) << 1) + 1]++; }
		} else ct[((ws - s) << 1) + 1]++;
	}
	p.xoff = (uint32_t)xk; p.yoff = (uint32_t)yk; p.coff = (int32_t)ck;
}

// wcns_vote, ecovlp.cpp:2185-2272: one block [s, e) of the sweep; id_a = the entries covering it (iterator A), the stretches are voted through iterator B
template <bool GRAPH> HB_HD int64_t hb_cns_vote(CnsCtx &C, uint32_t id_n, uint64_t s, uint64_t e, uint64_t *nec)
{
	uint64_t k, rr = 0, os, oe, wl, oc0, oc1, fI; CnsIt &occ = C.B;
	for (k = 0; k < id_n; k++) {
		CnsEnt &p = C.ent[C.A.act[k]]; const hb_wl_t &w = C.ov[p.ov].w[p.wid];
		const uint64_t q0 = (uint64_t)(int64_t)w.x_start, q1 = (uint64_t)((int64_t)w.x_end + 1);
		if (q1 <= e) rr = 1;
		os = q0 > s ? q0 : s; oe = q1 < e ? q1 : e;
		if (oe > os) hb_cns_sub_mm(C, p, (int64_t)os, (int64_t)oe, C.ct + (os - s)); // the reference offsets the count array by os - s WORDS (not 2 (os - s)): kept
	}
	wl = e - s; os = occ.mms; oe = occ.mme;
	for (k = 0; k < wl; k++) {
		oc0 = (C.ct[k << 1] >> 32) + 1; oc1 = (uint32_t)C.ct[k << 1] + 1;
		if (hb_cns_pass(oc0, oc1, 3, 0.500001) || oc1 < 3) {
			fI = 1;
			oc0 = (C.ct[(k << 1) + 1] >> 32) + 1; oc1 = (uint32_t)C.ct[(k << 1) + 1] + 1;
			if (hb_cns_pass(oc0, oc1, 3, 0.500001) || oc1 < 3) fI = 0;
			if (fI) {
				if (oe > os && os != (uint64_t)-1) { *nec += hb_cns_anchor<GRAPH>(C, os, oe, 0); if (C.need_full) return 0; }
				os = oe = (uint64_t)-1;
			}
			if (s + k == oe) oe++;
			else {
				if (oe > os && os != (uint64_t)-1) { *nec += hb_cns_anchor<GRAPH>(C, os, oe, 0); if (C.need_full) return 0; }
				os = s + k; oe = s + k + 1;
			}
		} else {
			if (oe > os && os != (uint64_t)-1) { *nec += hb_cns_anchor<GRAPH>(C, os, oe, 0); if (C.need_full) return 0; }
			os = oe = (uint64_t)-1;
		}
		C.ct[k << 1] = C.ct[(k << 1) + 1] = 0;
	}
	occ.mms = occ.mme = (uint64_t)-1;
	if (oe > os && os != (uint64_t)-1) { occ.mms = os; occ.mme = oe; }
	return (int64_t)rr;
}

// wcns_gen, ecovlp.cpp:2293-2424.  ov[0..n_ov) = the read's same-haplotype overlaps in the order of the de-duplicated list; ent / srt / act_a / act_b:
// room for one entry per aligned window of those overlaps (b32 too); key: one word per entry; ct: 2 * HB_CNS_WL words (zeroed here).
// Returns the number of corrected bases; C.out / C.out_n = the edit script; C.need_full / C.ovf tell when there is none.
template <bool GRAPH> HB_HD uint64_t hb_cns_read(CnsCtx &C, uint32_t n_ov, uint32_t *srt, uint32_t *act_a, uint32_t *act_b, uint64_t *key)
{
	uint32_t n_ent = 0; uint64_t nec = 0;
	for (uint32_t k = 0; k < n_ov; k++) {
		const CnsOv &z = C.ov[k];
		for (uint32_t i = 0; i < z.wn; i++) {
			if (hb_ualn_w(z.w[i])) continue;
			if (z.w[i].x_end >= z.w[i].x_start) {
				key[n_ent] = ((uint64_t)(uint32_t)z.w[i].x_start << 32) + n_ent;
				CnsEnt &p = C.ent[n_ent]; p.ov = k; p.wid = i; p.xoff = (uint32_t)z.w[i].x_start; p.yoff = (uint32_t)z.w[i].y_start; p.coff = 0;
				n_ent++;
			}
		}
	}
	hb_heapsort64(key, n_ent); // radix_sort_ec64 on distinct keys
	{ // entries that start together are ordered by their end — except the LAST such group, which the reference's loop never reaches (2362-2378): kept
		int64_t k, i, t;
		for (k = 1, i = 0; k < (int64_t)n_ent; k++) {
			if ((key[k] >> 32) != (key[i] >> 32)) {
				if (k - i > 1) {
					for (t = i; t < k; t++) { const CnsEnt &cp = C.ent[(uint32_t)key[t]]; uint64_t m = (uint64_t)((int64_t)C.ov[cp.ov].w[cp.wid].x_end + 1); m <<= 32; m += (uint32_t)key[t]; key[t] = m; }
					hb_heapsort64(key + i, (uint32_t)(k - i));
				}
				i = k;
			}
		}
	}
	for (uint32_t t = 0; t < n_ent; t++) srt[t] = (uint32_t)key[t];
	for (uint32_t t = 0; t < 2 * HB_CNS_WL; t++) C.ct[t] = 0;
	C.A.srt = srt; C.A.act = act_a; C.A.i = 0; C.A.srt_n = n_ent; C.A.act_n = 0; C.A.rr = C.A.ru = 0; C.A.mms = C.A.mme = (uint64_t)-1;
	C.B.srt = srt; C.B.act = act_b; C.B.i = 0; C.B.srt_n = n_ent; C.B.act_n = 0; C.B.rr = C.B.ru = 0; C.B.mms = C.B.mme = (uint64_t)-1;
	C.out_n = 0; C.has_win = 0; C.ax_start = C.ax_end = -1; C.ovf = 0; C.need_full = 0; C.b32_n = 0;
	int64_t s = 0, e = HB_CNS_WL, rr = 0; if (e > C.ql) e = C.ql;
	for (; s < C.ql;) {
		const uint32_t rn = hb_cns_iter(C, C.A, s, e, rr, 0);
		rr = hb_cns_vote<GRAPH>(C, rn, (uint64_t)s, (uint64_t)e, &nec);
		if (C.need_full) return nec;
		s += HB_CNS_WL; e += HB_CNS_WL; if (e > C.ql) e = C.ql;
	}
	if (C.B.mme > C.B.mms && C.B.mms != (uint64_t)-1) { nec += hb_cns_anchor<GRAPH>(C, C.B.mms, C.B.mme, 0); if (C.need_full) return nec; }
	nec += hb_cns_anchor<GRAPH>(C, (uint64_t)C.ql, (uint64_t)C.ql, 1);
	return nec;
}
