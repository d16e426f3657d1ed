// seq_ref.h — one-thread restatements kept ONLY for tests/hostemu: the straight-line forms of the chaining DP (Hash_Table.cpp:2124-2176) and of the
// one-pass sketch (sketch.cpp:454-579) that the product's warp-parallel chain kernel and two-stage sketch are checked against on the golden vectors.
// Not part of the product: nothing under hifiasm_b200/ includes this file.
#pragma once
#include "../../hifiasm_b200/csrc/hb_chain.cuh"
#include "../../hifiasm_b200/csrc/hb_sketch.cuh"

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
