// hb_final.cuh — per-read merge of the final overlap pass.
//
// Behaviour of worker_hap_dc_ec_gen_new_idx after h_ec_lchain (ecovlp.cpp:
// 3959-3978): overlap_region_sort_y_id, h_ec_lchain_fast_new (5047-5196) and
// push_ff_ovlp (2641-2694).  The exact-match test (exact_ec_check on decoded
// strings, ecovlp.cpp:2803) is done beforehand by a warp-per-chain kernel on
// the 2-bit packed reads; this routine consumes its flags.
#pragma once
#include "hb_chain.cuh"

// working record of one overlap while it is merged
struct FinOv {
	uint32_t x_pos_s, x_pos_e, y_id, y_pos_s, y_pos_e;
	uint32_t nhe;      // overlap_region.non_homopolymer_errors
	uint32_t slot;     // chain slot (for the exact flag); 0xffffffff for re-added records
	uint8_t strand, is_match, el, strong, wli, pad[3];
};
struct KeyYid { HB_HD uint32_t operator()(const FinOv &o) const { return o.y_id; } };

// exact_ec_check (ecovlp.cpp:2803-2808) of query[qs,qe) against the strand-
// oriented target[ts,te) on packed reads (recover_UC_Read_sub_region,
// Process_Read.cpp:524-616).  One thread, byte-wise; the product uses the
// warp version in final.cu, this one documents the semantics and serves hostemu.
HB_HD int hb_exact_seq(const DevReads &R, uint64_t qid, uint64_t qs, uint64_t qe, uint64_t tid, uint64_t ts, uint64_t te, int strand)
{
	if (qe - qs != te - ts) return 0;
	const uint8_t *q = R.packed + R.off[qid], *t = R.packed + R.off[tid];
	uint64_t n = qe - qs, i, tl = R.len[tid];
	for (i = 0; i < n; i++) {
		int a = hb_base(q, qs + i), b = strand ? 3 - hb_base(t, tl - 1 - (ts + i)) : hb_base(t, ts + i);
		if (a != b) return 0;
	}
	// N bases are stored as A plus a side list: equal strings need equal N sets
	uint64_t qi = R.noff[qid], qn = R.noff[qid + 1], ti = R.noff[tid], tn = R.noff[tid + 1];
	if (qi == qn && ti == tn) return 1;
	uint64_t cq = 0, ct = 0;
	for (i = qi; i < qn; i++) { uint64_t p = R.npos[i]; if (p >= qs && p < qe) {
		uint64_t rel = p - qs, tp = strand ? tl - 1 - (ts + rel) : ts + rel, j; int hit = 0;
		for (j = ti; j < tn; j++) if (R.npos[j] == tp) { hit = 1; break; }
		if (!hit) return 0;
		cq++; } }
	for (i = ti; i < tn; i++) { uint64_t p = R.npos[i]; uint64_t f = strand ? tl - 1 - p : p; if (f >= ts && f < te) ct++; }
	return cq == ct;
}

// in0/in1: previous paf / reverse_paf of the read (in0 is modified: el reset).
// ch/idx/n_ol: chains kept by hb_chain_post; exact[slot]: flag of the exact kernel.
// ov: scratch of n_ol + n0 records; srt: scratch of n0 + n1 words.
// out0/out1 receive the new paf / reverse_paf (capacity n_ol + n0 each).
HB_HD void hb_final_merge(const DevReads &R, uint32_t rid, const hb_chain_t *ch, const uint32_t *idx, uint32_t n_ol, const uint8_t *exact,
                          hb_ma_hit_t *in0, uint32_t n0, const hb_ma_hit_t *in1, uint32_t n1, FinOv *ov, uint64_t *srt,
                          hb_ma_hit_t *out0, uint32_t *m0, hb_ma_hit_t *out1, uint32_t *m1, unsigned long long *stat, const RsScratch &W)
{
	const double sh = 0.866666; // ecovlp.cpp:3970
	uint64_t k, i, l, m, ns = 0; uint32_t n = n_ol, is_usrt = 0; FinOv t; KeyYid key;
	for (k = 0; k < n_ol; k++) {
		const hb_chain_t &c = ch[idx[k]]; FinOv &z = ov[k];
		z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_id = c.y_id; z.y_pos_s = c.y_pos_s; z.y_pos_e = c.y_pos_e;
		z.strand = (uint8_t)c.y_pos_strand; z.nhe = 0; z.slot = idx[k]; z.is_match = 0; z.el = 0; z.strong = 0; z.wli = 0;
	}
	int rs_ovf = hb_rs_sort32(ov, ov + n, key, W); // overlap_region_sort_y_id, ecovlp.cpp:3959
	for (k = 0; k < n0; k++) srt[ns++] = ((uint64_t)in0[k].tn << 1 | (uint64_t)in0[k].rev) << 32 | (k << 1) | 0; // ecovlp.cpp:5057-5070
	for (k = 0; k < n1; k++) srt[ns++] = ((uint64_t)in1[k].tn << 1 | (uint64_t)in1[k].rev) << 32 | (k << 1) | 1;
	for (i = 1; i < ns; i++) { uint64_t v = srt[i]; for (l = i; l > 0 && srt[l - 1] > v; --l) srt[l] = srt[l - 1]; srt[l] = v; } // keys distinct: any sort
	for (k = m = 0, i = 0; k < n; k++) { // ecovlp.cpp:5073-5135
		FinOv &z = ov[k]; uint64_t tid = z.y_id, trev = z.strand, bq[2], bt[2]; int om, ex = exact[z.slot];
		z.nhe = 0;
		bq[0] = z.x_pos_s; bq[1] = (uint64_t)z.x_pos_e + 1; bt[0] = z.y_pos_s; bt[1] = (uint64_t)z.y_pos_e + 1;
		for (; i < ns && (srt[i] >> 32) < (tid << 1 | trev); i++) {}
		if (i < ns && (srt[i] >> 32) == (tid << 1 | trev)) {
			const hb_ma_hit_t *p; uint32_t is_match; uint64_t aq[2], at[2], os, oe, ovlp; int from0 = !(srt[i] & 1);
			om = 1; z.el = 0; z.is_match = 0;
			if (!from0) { p = &in1[((uint32_t)srt[i]) >> 1]; is_match = 2; }
			else { p = &in0[((uint32_t)srt[i]) >> 1]; is_match = 1; }
			z.strong = (uint8_t)p->ml; z.wli = p->no_l_indel;
			aq[0] = (uint32_t)p->qns; aq[1] = p->qe; at[0] = p->ts; at[1] = p->te;
			os = aq[0] > bq[0] ? aq[0] : bq[0]; oe = aq[1] < bq[1] ? aq[1] : bq[1];
			ovlp = oe > os ? oe - os : 0;
			if (!(ovlp && (double)ovlp >= (double)(aq[1] - aq[0]) * sh && (double)ovlp >= (double)(bq[1] - bq[0]) * sh)) om = 0;
			z.nhe += (uint32_t)((aq[1] - aq[0]) - ovlp);
			os = at[0] > bt[0] ? at[0] : bt[0]; oe = at[1] < bt[1] ? at[1] : bt[1];
			ovlp = oe > os ? oe - os : 0;
			if (!(ovlp && (double)ovlp >= (double)(at[1] - at[0]) * sh && (double)ovlp >= (double)(bt[1] - bt[0]) * sh)) om = 0;
			z.nhe += (uint32_t)((at[1] - at[0]) - ovlp);
			if (om) {
				if (from0 && p->el == 1) in0[((uint32_t)srt[i]) >> 1].el = 0;
				if (ex) {
					if (is_match == 2) { z.strong = 0; z.wli = 1; }
					is_match = 1; z.el = 1;
				}
				z.is_match = (uint8_t)is_match;
			}
		} else {
			om = 0;
			if (ex) { z.strong = 0; z.wli = 1; z.el = 1; z.is_match = 1; om = 1; }
		}
		if (om) { if (m != k) { t = ov[m]; ov[m] = ov[k]; ov[k] = t; } m++; }
	}
	n = (uint32_t)m;
	for (k = 0; k < n0; k++) { // ecovlp.cpp:5138-5165
		if (in0[k].el) {
			FinOv &z = ov[n++];
			z.y_id = in0[k].tn; z.strand = (uint8_t)in0[k].rev;
			z.x_pos_s = (uint32_t)in0[k].qns; z.x_pos_e = in0[k].qe - 1;
			z.y_pos_s = in0[k].ts; z.y_pos_e = in0[k].te - 1;
			z.nhe = 0; z.is_match = 1; z.el = 1; z.slot = 0xffffffffu;
			z.strong = (uint8_t)in0[k].ml; z.wli = in0[k].no_l_indel;
			is_usrt = 1;
		}
	}
	if (is_usrt) rs_ovf |= hb_rs_sort32(ov, ov + n, key, W);
	if (n > 1) { // ecovlp.cpp:5169-5195
		uint64_t mm_k, s; int64_t mm_sc, sc;
		for (k = 1, l = m = 0; k <= n; k++) {
			if (k == n || ov[k].y_id != ov[l].y_id) {
				mm_k = l;
				if (k - l > 1) {
					for (s = l, mm_sc = INT32_MIN, mm_k = (uint64_t)-1; s < k; s++) {
						sc = ov[s].nhe; sc = -sc;
						if (sc > mm_sc || (sc == mm_sc && (ov[mm_k].x_pos_e + 1 - ov[mm_k].x_pos_s) < (ov[s].x_pos_e + 1 - ov[s].x_pos_s))) { mm_sc = sc; mm_k = s; }
					}
				}
				if (mm_k != (uint64_t)-1) {
					if (mm_k != m) { t = ov[mm_k]; ov[mm_k] = ov[m]; ov[m] = t; }
					m++;
				}
				l = k;
			}
		}
		n = (uint32_t)m;
	}
	// push_ff_ovlp x2, ecovlp.cpp:2641-2694
	uint32_t c0 = 0, c1 = 0; unsigned long long st[7] = { 0, 0, 0, 0, 0, 0, 0 };
	st[6] = (unsigned long long)rs_ovf; // sort stack overflow: the host fails the pass
	for (k = 0; k < n; k++) {
		const FinOv &o = ov[k]; hb_ma_hit_t *z;
		if (o.is_match == 1) z = &out0[c0++]; else if (o.is_match == 2) z = &out1[c1++]; else continue;
		z->qns = (uint64_t)rid << 32 | o.x_pos_s; z->tn = o.y_id;
		z->qe = o.x_pos_e + 1; z->ts = o.y_pos_s; z->te = o.y_pos_e + 1;
		z->rev = o.strand; z->bl = R.len[o.y_id] & 0x7fffffffu;
		z->ml = o.strong & 1; z->no_l_indel = o.wli; z->el = o.el;
		if (z->rev) { z->ts = z->bl - o.y_pos_e - 1; z->te = z->bl - o.y_pos_s; }
		z->del = 0;
		for (int b = 0; b < 6; b++) z->pad[b] = 0;
		if (o.is_match == 1) { if (z->ml == 1) st[2]++; if (z->ml == 0) st[3]++; if (z->el == 1) st[4]++; if (z->no_l_indel) st[5]++; }
	}
	st[0] = c0; st[1] = c1;
	*m0 = c0; *m1 = c1;
	if (stat) for (int b = 0; b < 7; b++) stat[b] = st[b];
}
