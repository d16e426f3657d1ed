// hb_ecphase.cuh — phasing of the accepted overlaps of one read: rphase_hc (Correct.cpp:20191-20320),
// HiFi path (bd = 0, occ_thres = 1, no HPC mask, no base qualities, no large-indel phasing).
//
// The reference sweeps the query in 428-column windows with per-window cursors into every cigar; the quantities it derives are
// position-local, so this version works on whole-read arrays instead:
//   1. cnt[x]  = number of aligned windows whose cigar has a mismatch column at query position x (saturating at 127)   [extract_sub_cigar_hc, set_f]
//   2. SNP candidate sites = positions with cnt > 1
//   3. evidence = for every candidate site, every window that covers it with a match / mismatch column                   [extract_sub_cigar_hc, !set_f]
//   4. per site: allele statistics (push_info, Correct.cpp:10511) -> per overlap: the greedy haplotype call
//      (generate_haplotypes_naive_HiFi, Correct.cpp:8845) -> is_match 1 / 2 and `strong`
// One thread per read: step 4 is a greedy sequential pass over the read's overlaps, and steps 1-3 only walk cigar RUNS
// (a few thousand per read), so there is nothing to gain from splitting a read further.  Two launches: count (evidence
// and site numbers per read), then fill + decide with exactly-sized scratch.
#pragma once
#include "hb_ecaln.cuh"
#include "hb_chain.cuh"
#include "hb_warp.cuh"

struct alignas(8) PhEv { uint32_t site, ov, osite, cov; uint8_t type, base, pad[6]; };           // haplotype_evdience (Correct.h:130-143); base = 0..3, 4 = N; 24 bytes: the array doubles as a u64 work list
struct PhSnp { uint32_t id, overlap_num, occ_0, occ_1, occ_2, site; int32_t score; uint32_t pad; }; // SnpStats (Correct.h:152-167)
struct PhOv { const hb_wl_t *w; uint32_t wn; const uint16_t *pool; uint32_t y_id, rev, align_length; uint8_t is_match; int8_t strong; };

HB_HD bool hb_ph_ualn(const hb_wl_t &u) { return u.error == INT16_MAX && u.clen == 0 && u.extra_end < 0; } // is_ualn_win, Correct.h:1362

// ---- one WARP per read (hb_warp.cuh; a warp of one lane in tests/hostemu) ---------------------------------------------------------------------
// Scratch of a read (zeroed by the caller): PhBits over nwd = ceil(ql / 32) words — s1 / s2: a mismatch column seen at this query position by
// at least one / at least two alignments (the reference's saturating per-position counter is only ever tested for > 1), pre: candidate sites in front
// of each word, so that "sites inside [a, b)" is two popcounts.  Lanes own alignment windows and walk cigar RUNS; nothing is per column.
struct PhBits { uint32_t *s1, *s2, *pre; uint32_t nwd; };
HB_HD PhBits hb_ph_bits(uint8_t *scratch, int64_t ql) { PhBits b; b.nwd = (uint32_t)((ql + 31) >> 5); b.s1 = (uint32_t *)scratch; b.s2 = b.s1 + b.nwd; b.pre = b.s2 + b.nwd; return b; }
HB_HD uint64_t hb_ph_bits_bytes(uint64_t ql) { return ((((ql + 31) >> 5) * 3 + 1) * 4 + 15) & ~15ull; }
HB_HD uint32_t hb_ph_rank(const PhBits &b, int64_t x) { const uint32_t w = (uint32_t)(x >> 5), r = (uint32_t)(x & 31); return w < b.nwd ? b.pre[w] + (r ? (uint32_t)hb_popc64(b.s2[w] & ((1u << r) - 1)) : 0u) : b.pre[b.nwd]; }
#if defined(__CUDA_ARCH__)
HB_D uint32_t hb_atom_or32(uint32_t *p, uint32_t v) { return atomicOr(p, v); }
#else
inline uint32_t hb_atom_or32(uint32_t *p, uint32_t v) { const uint32_t o = *p; *p |= v; return o; }
#endif
// f(j, window) for every aligned window with a cigar of every overlap.
template <typename F> HB_HD void hb_ph_for_windows(const PhOv *ov, uint32_t n_ov, F f)
{
	// lanes take OVERLAPS (after step B an overlap's aligned stretch is mostly ONE fused window with a long cigar: windows give no parallelism)
	const int lane = hb_lane();
	for (uint32_t j = lane; j < n_ov; j += HB_WS) for (uint32_t w = 0; w < ov[j].wn; w++) {
		const hb_wl_t &u = ov[j].w[w];
		if (hb_ph_ualn(u) || u.x_end < u.x_start || !u.clen) continue;
		f(j, u);
	}
}
// pass 1 (extract_sub_cigar_hc with set_f, then the scan for counts > 1): candidate sites and the number of evidence records
HB_HD void hb_ph_count_w(const PhOv *ov, uint32_t n_ov, uint8_t *scratch, int64_t ql, uint32_t *n_site, uint32_t *n_ev)
{
	const int lane = hb_lane(); PhBits B = hb_ph_bits(scratch, ql);
	hb_ph_for_windows(ov, n_ov, [&](uint32_t j, const hb_wl_t &u) {
		int64_t xk = u.x_start; const int64_t e0 = (int64_t)u.x_end + 1; const uint16_t *cg = ov[j].pool + u.cidx;
		for (uint32_t ci = 0; ci < u.clen && xk < e0; ci++) {
			const uint32_t op = cg[ci] >> 14, cl = cg[ci] & 0x3fff; const int64_t ws = xk;
			if (op != 2) xk += cl;
			if (op != 1) continue;
			const int64_t oe = xk < e0 ? xk : e0;
			for (int64_t t = ws; t < oe; t++) { const uint32_t bit = 1u << (t & 31); if (hb_atom_or32(B.s1 + (t >> 5), bit) & bit) hb_atom_or32(B.s2 + (t >> 5), bit); }
		}
	});
	hb_wsync();
	uint32_t run = 0; // sites in front of each word: chunks of HB_WS words, one warp scan per chunk
	for (uint32_t w0 = 0; w0 < B.nwd; w0 += HB_WS) {
		const uint32_t w = w0 + (uint32_t)lane, c = w < B.nwd ? (uint32_t)hb_popc64(B.s2[w]) : 0u; uint32_t tot;
		const uint32_t ex = hb_wscan(c, &tot);
		if (w < B.nwd) B.pre[w] = run + ex;
		run += tot;
	}
	if (lane == 0) B.pre[B.nwd] = run;
	hb_wsync();
	int32_t ne = 0;
	hb_ph_for_windows(ov, n_ov, [&](uint32_t j, const hb_wl_t &u) {
		int64_t xk = u.x_start; const int64_t e0 = (int64_t)u.x_end + 1; const uint16_t *cg = ov[j].pool + u.cidx;
		for (uint32_t ci = 0; ci < u.clen && xk < e0; ci++) {
			const uint32_t op = cg[ci] >> 14, cl = cg[ci] & 0x3fff; const int64_t ws = xk;
			if (op != 2) xk += cl;
			if (op > 1) continue;
			const int64_t oe = xk < e0 ? xk : e0;
			if (oe > ws) ne += (int32_t)(hb_ph_rank(B, oe) - hb_ph_rank(B, ws));
		}
	});
	*n_site = run; *n_ev = (uint32_t)hb_wsum(ne);
}

// pass 2: evidence in (site, overlap) order, allele statistics, haplotype call.  Every lane calls it; lanes share the evidence scatter, lane 0 runs the
// reference's sequential statistics and greedy call.  scratch = pass 1's bit arrays of the read.
// site_pos[n_site], site_off[n_site + 1], ev[n_ev], ev2[n_ev], snp[4 * n_site], ord[n_ov] (uint64)
HB_HD void hb_ph_decide_w(const DevReads &R, uint64_t qid, PhOv *ov, uint32_t n_ov, uint8_t *scratch, int64_t ql, uint32_t n_site, uint32_t n_ev,
                          uint32_t *site_pos, uint32_t *site_off, PhEv *ev, PhEv *ev2, PhSnp *snp, uint64_t *ord, uint32_t *ov_off, const RsScratch &W, int s_hap_cov, int infor_cov, double up, int *ovf)
{
	if (!n_site || !n_ev) return;
	const int lane = hb_lane(); const PhBits B = hb_ph_bits(scratch, ql);
	const RdView Q = hb_rd_view(R, qid, 0);
	const uint32_t ns = n_site;
	for (uint32_t w = lane; w < B.nwd; w += HB_WS) { uint32_t m = B.s2[w], k = B.pre[w]; while (m) { const int b = hb_ctz64(m); m &= m - 1; site_pos[k++] = (w << 5) + (uint32_t)b; } }
	// evidence in (site, overlap) order without an ordered pass: a bit mask per site over the overlaps that cover it with a match / mismatch column
	// (pass A, lanes over overlaps, atomicOr) gives every record its place — start of the site + number of covering overlaps below it (popcounts) —
	// and pass B writes it there.  The masks borrow ev2 (unused until the statistics): ns * MW words, MW = ceil(n_ov / 32).
	const uint32_t MW = (n_ov + 31) >> 5; uint32_t *mask = (uint32_t *)ev2;
	if ((uint64_t)ns * MW * 4 > (uint64_t)n_ev * sizeof(PhEv)) { *ovf = 1; return; } // cannot happen while every site has >= 2 records and n_ov <= 384
	for (uint64_t k = lane; k < (uint64_t)ns * MW; k += HB_WS) mask[k] = 0;
	hb_wsync();
	hb_ph_for_windows(ov, n_ov, [&](uint32_t j, const hb_wl_t &u) {
		int64_t xk = u.x_start; const int64_t e0 = (int64_t)u.x_end + 1; const uint16_t *cg = ov[j].pool + u.cidx;
		for (uint32_t ci = 0; ci < u.clen && xk < e0; ci++) {
			const uint32_t op = cg[ci] >> 14, cl = cg[ci] & 0x3fff; const int64_t ws = xk;
			if (op != 2) xk += cl;
			if (op > 1) continue;
			const int64_t oe = xk < e0 ? xk : e0;
			if (oe > ws) for (uint32_t si = hb_ph_rank(B, ws), se = hb_ph_rank(B, oe); si < se; si++) hb_atom_or32(mask + (uint64_t)si * MW + (j >> 5), 1u << (j & 31));
		}
	});
	hb_wsync();
	{ uint32_t run = 0; // site_off[k] = start of site k
		for (uint32_t k0 = 0; k0 < ns; k0 += HB_WS) {
			const uint32_t k = k0 + (uint32_t)lane; uint32_t c = 0, tot;
			if (k < ns) for (uint32_t q = 0; q < MW; q++) c += (uint32_t)hb_popc64(mask[(uint64_t)k * MW + q]);
			const uint32_t ex = hb_wscan(c, &tot);
			if (k < ns) site_off[k] = run + ex;
			run += tot;
		}
		if (lane == 0) site_off[ns] = run;
	}
	hb_wsync();
	for (uint32_t j = lane; j < n_ov; j += HB_WS) {
		const RdView T = hb_rd_view(R, ov[j].y_id, ov[j].rev);
		for (uint32_t w = 0; w < ov[j].wn; w++) {
			const hb_wl_t &u = ov[j].w[w];
			if (hb_ph_ualn(u) || u.x_end < u.x_start || !u.clen) continue;
			int64_t xk = u.x_start, yk = u.y_start; const int64_t e0 = (int64_t)u.x_end + 1; const uint16_t *cg = ov[j].pool + u.cidx;
			for (uint32_t ci = 0; ci < u.clen && xk < e0; ci++) {
				const uint32_t op = cg[ci] >> 14, cl = cg[ci] & 0x3fff; const int64_t ws = xk;
				if (op != 2) xk += cl;
				if (op != 3) yk += cl;
				if (op > 1) continue;
				const int64_t oe = xk < e0 ? xk : e0;
				if (oe > ws) for (uint32_t si = hb_ph_rank(B, ws), se = hb_ph_rank(B, oe); si < se; si++) {
					const uint32_t *m = mask + (uint64_t)si * MW; uint32_t rk = (uint32_t)hb_popc64(m[j >> 5] & ((1u << (j & 31)) - 1));
					for (uint32_t q = 0; q < (j >> 5); q++) rk += (uint32_t)hb_popc64(m[q]);
					const int64_t t = site_pos[si]; PhEv e;
					e.site = (uint32_t)t; e.ov = j; e.osite = (uint32_t)(t - xk + yk); e.cov = 1; e.type = (uint8_t)op; for (int z = 0; z < 6; z++) e.pad[z] = 0;
					e.base = (uint8_t)(op == 0 ? Q.at(t) : T.at(t - xk + yk));
					ev[site_off[si] + rk] = e;
				}
			}
		}
	}
	hb_wsync();
	for (uint32_t k0 = 0; k0 < ns; k0 += HB_WS) { // the statistics below read site_off[k] as the END of site k: shift by one, chunk by chunk in ascending order
		const uint32_t k = k0 + (uint32_t)lane, v = k < ns ? site_off[k + 1] : 0u;
		hb_wsync();
		if (k < ns) site_off[k] = v;
		hb_wsync();
	}
	// push_info per site (Correct.cpp:10511-10600, oa = NULL, v8 = NULL): allele statistics, evidence kept only for real alleles.  A lane takes a site
	// (its records are contiguous); the places of its statistics and kept records come from two warp scans over the chunk of 32 sites.
	uint32_t n_snp = 0, m_ev = 0;
	for (uint32_t k0 = 0; k0 < ns; k0 += HB_WS) {
		const uint32_t k = k0 + (uint32_t)lane; uint32_t beg = 0, end = 0, m = 0, nrec = 0;
		uint64_t occ_0 = 0, occ_1[5] = { 0, 0, 0, 0, 0 }, occ_2 = 0, diff = 0;
		if (k < ns) {
			beg = k ? site_off[k - 1] : 0; end = site_off[k];
			for (uint32_t i = beg; i < end; i++) { if (ev[i].type == 0) occ_0++; else { occ_1[ev[i].base]++; diff++; } occ_2++; }
			if (!(occ_0 == 0 || diff <= 1)) {
				for (int b = 0; b < 4; b++) if (occ_1[b] >= 2) m++;
				if (m) for (uint32_t i = beg; i < end; i++) { const PhEv &e = ev[i]; if (e.type == 0 || occ_1[e.base < 4 ? e.base : 4] >= 2 && e.base < 4) nrec++; }
			}
		}
		uint32_t tm, tr; const uint32_t em = hb_wscan(m, &tm), er = hb_wscan(nrec, &tr);
		if (m) {
			int64_t sid[5] = { -1, -1, -1, -1, -1 }; uint32_t q = n_snp + em;
			for (int b = 0; b < 4; b++) if (occ_1[b] >= 2) {
				PhSnp &p = snp[q]; p.id = q; p.occ_0 = (uint32_t)(1 + occ_0); p.occ_1 = (uint32_t)occ_1[b]; p.occ_2 = (uint32_t)(occ_2 - p.occ_0 - p.occ_1);
				p.site = site_pos[k]; p.score = -1; p.overlap_num = (uint32_t)occ_2; p.pad = 0; sid[b] = q++;
			}
			uint32_t w = m_ev + er;
			for (uint32_t i = beg; i < end; i++) {
				PhEv e = ev[i];
				if (e.type == 0) e.osite = q - 1;
				else if (e.base < 4 && sid[e.base] >= 0) { e.cov = e.osite; e.osite = (uint32_t)sid[e.base]; }
				else continue;
				ev2[w++] = e;
			}
		}
		n_snp += tm; m_ev += tr;
	}
	hb_wsync();
	if (lane != 0) return;
	if (!m_ev) return;
	// generate_haplotypes_naive_HiFi, Correct.cpp:8845-9110 (multi_check = 1, st_max = -1 => is_st_bs is never true)
	// (a) drop the SNP sites that touch another SNP site; evidence follows its statistics
	uint32_t m_snp = 0, m_list = 0;
	{
		uint32_t i = 0;
		for (uint32_t k = 1, l = 0; k <= n_snp; ++k) {
			if (k == n_snp || snp[k].site != snp[l].site) {
				if (l > 0 && snp[l].site == snp[l - 1].site + 1) { l = k; continue; }
				if (k < n_snp && snp[l].site + 1 == snp[k].site) { l = k; continue; }
				for (; i < m_ev && ev2[i].site != snp[l].site; i++);
				const uint32_t m_off = l - m_snp;
				for (; i < m_ev && ev2[i].site == snp[l].site; i++) { ev2[m_list] = ev2[i]; ev2[m_list++].osite -= m_off; }
				for (; l < k; l++) snp[m_snp++] = snp[l];
			}
		}
	}
	n_snp = m_snp; m_ev = m_list;
	if (!n_snp || !m_ev) return;
	// (b) evidence grouped by overlap with klib's unstable radix sort restated move for move (hb_rs_sort32): the order of an overlap's
	//     sites decides which statistics the reference's stale `s` pointer refers to in the multi_check block below
	{ auto key = [](const PhEv &e) -> uint32_t { return e.ov; }; if (hb_rs_sort32(ev2, ev2 + m_ev, key, W, n_ov <= 256 ? 0 : n_ov <= 65536 ? 8 : 24)) { *ovf = 1; return; } } // (keys < n_ov: the upper digit levels would find one bucket each and move nothing)
	PhEv *L = ev2;
	for (uint32_t j = 0; j <= n_ov; j++) ov_off[j] = 0;
	for (uint32_t i = 0; i < m_ev; i++) ov_off[L[i].ov + 1]++;
	for (uint32_t j = 0; j < n_ov; j++) ov_off[j + 1] += ov_off[j];
#define PH_REAL(s_) (!((s_).occ_0 < 2 || (s_).occ_1 < 2))
#define PH_ALLELE(s_) ((int)(s_).occ_0 >= s_hap_cov && (int)(s_).occ_1 >= infor_cov)
	int64_t sp = -1, tp = -1; // the reference's function-level `SnpStats *s, *t` (indices; -1 = NULL): they are read stale further down
	uint32_t n_ord = 0;
	for (uint32_t j = 0; j < n_ov; j++) {
		uint64_t o = 0;
		for (uint32_t i = ov_off[j]; i < ov_off[j + 1]; i++) { if (L[i].type != 1) continue; sp = L[i].osite; const PhSnp &q = snp[sp]; if (PH_REAL(q) && PH_ALLELE(q)) o++; }
		if (o > 0) ord[n_ord++] = ((uint64_t)(0xffffffffu - o) << 32) | ov_off[j]; // more alleles first, then list position
	}
	if (n_ord) {
		hb_heapsort64(ord, n_ord); // keys are unique (list positions): any sort gives the reference's order; an insertion sort is quadratic on repeat-rich reads
		for (uint32_t k = 0; k < n_ord; k++) {
			const uint32_t l = (uint32_t)ord[k], j = L[l].ov; uint64_t o = 0;
			for (uint32_t i = l; i < ov_off[j + 1]; i++) { if (L[i].type != 1) continue; sp = L[i].osite; const PhSnp &q = snp[sp]; if (PH_REAL(q) && PH_ALLELE(q)) o++; }
			if (!o) continue;
			if (ov[j].is_match == 1) ov[j].is_match = 2;
			for (uint32_t i = l; i < ov_off[j + 1]; i++) {
				sp = L[i].osite;
				if (L[i].type == 1) snp[sp].score = 1;
				else for (int64_t z = L[i].osite; z >= 0; z--) { tp = z; if (snp[sp].site != snp[z].site) break; snp[z].occ_0 -= L[i].cov; }
			}
		}
		for (uint32_t k = 0; k < n_ord; k++) {
			const uint32_t l = (uint32_t)ord[k], j = L[l].ov; uint64_t o = 0;
			for (uint32_t i = l; i < ov_off[j + 1]; i++) { if (L[i].type != 1) continue; sp = L[i].osite; const PhSnp &q = snp[sp]; if (PH_REAL(q) && q.score == 1) o++; }
			if (ov[j].is_match == 1 && o > 0) ov[j].is_match = 2;
		}
		for (uint32_t j = 0; j < n_ov; j++) if (ov[j].is_match == 1) for (uint32_t i = ov_off[j]; i < ov_off[j + 1]; i++) if (L[i].type == 1) snp[L[i].osite].score = -1;
	}
	// multi_check: isolated minor alleles shared by >= 2 same-haplotype overlaps (Correct.cpp:9036-9083); stat ids collected in `ev` (free now)
	{
		uint64_t *srt = (uint64_t *)ev; uint32_t sn = 0;
		for (uint32_t j = 0; j < n_ov; j++) {
			if (ov_off[j] == ov_off[j + 1] || ov[j].is_match == 2) continue;
			uint32_t o = 0;
			for (uint32_t i = ov_off[j]; i < ov_off[j + 1]; i++) {
				if (L[i].type != 1) continue;
				sp = L[i].osite; const PhSnp &q = snp[sp];
				if (!PH_REAL(q) || PH_ALLELE(q) || q.score == 1) continue;
				srt[sn + o++] = L[i].osite;
			}
			if ((double)o >= (double)ov[j].align_length * up) {
				uint64_t *a = srt + sn; uint32_t z = 0;
				hb_heapsort64(a, o); // (bare keys: any sort gives the reference's array; an insertion sort is quadratic on repeat-rich reads)
				for (uint32_t i = 0; i < o; i++) { // sp / tp are NOT reset here: for i = 0 and i = o-1 the reference reads what earlier code left in s / t
					if (i > 0) sp = (int64_t)a[i - 1];
					if (i + 1 < o) tp = (int64_t)a[i + 1];
					if (sp >= 0 && snp[sp].site + 32 > snp[a[i]].site) continue;
					if (tp >= 0 && snp[a[i]].site + 32 > snp[tp].site) continue;
					a[z++] = a[i];
				}
				if (z >= 2) sn += z;
			}
		}
		if (sn) {
			hb_heapsort64(srt, sn);
			for (uint32_t k = 1, l = 0; k <= sn; ++k) { if (k == sn || srt[k] != srt[l]) { if (k - l >= 2) snp[srt[l]].score = 1; } l = k; }
		}
	}
	for (uint32_t j = 0; j < n_ov; j++) { // Correct.cpp:9085-9108
		if (ov_off[j] == ov_off[j + 1]) continue;
		if (ov[j].is_match == 2) ov[j].strong = 1;
		else if (ov[j].is_match == 1) for (uint32_t i = ov_off[j]; i < ov_off[j + 1]; i++) {
			const PhSnp &q = snp[L[i].osite];
			if (q.score == 1 && PH_REAL(q)) { ov[j].strong = 1; if (L[i].type == 1) { ov[j].is_match = 2; break; } }
		}
	}
#undef PH_REAL
#undef PH_ALLELE
}

// ---- dedup_chains (ecovlp.cpp:2984-3030) + push_ne_ovlp(flag = 2, ec = NULL) (ecovlp.cpp:2585-2638): the round's reverse_paf[i] -------------
// ph[0..n) = the read's overlaps after phasing (hb_phase_t, st == 2 = accepted, list order); ord = scratch of n u64 (y_id | index pairs for
// overlap_region_sort_y_id, whose unstable radix sort is restated move for move); out = up to n records.  Returns the number of records.
struct PhPair { uint32_t y_id, idx; };
// dedup_chains alone: ord[0..keep) = the surviving overlaps (indices into ph[]) in the order the reference leaves them
HB_HD uint32_t hb_ec_dedup(const hb_phase_t *ph, uint32_t n, PhPair *ord, const RsScratch &W, int *ovf)
{
	uint32_t m = 0;
	for (uint32_t j = 0; j < n; j++) if (ph[j].st == 2) { ord[m].y_id = ph[j].y_id; ord[m].idx = j; m++; }
	uint32_t keep = m;
	if (m > 1) {
		auto key = [](const PhPair &p) -> uint32_t { return p.y_id; };
		if (hb_rs_sort32(ord, ord + m, key, W)) { *ovf = 1; return 0; }
		uint32_t k, l; keep = 0;
		for (k = 1, l = 0; k <= m; k++) {
			if (k == m || ord[k].y_id != ord[l].y_id) {
				uint32_t mm_k = l;
				if (k - l > 1) {
					int64_t mm_sc = INT32_MIN; uint32_t mm_m = 3; mm_k = 0xffffffffu;
					for (uint32_t s = l; s < k; s++) {
						const hb_phase_t &z = ph[ord[s].idx]; const int64_t len = (int64_t)z.x_pos_e + 1 - z.x_pos_s, sc = len - (int64_t)z.nh_err * 12; bool sf = false;
						if (z.is_match < mm_m) sf = true;
						else if (z.is_match == mm_m) {
							if (sc > mm_sc) sf = true;
							else if (sc == mm_sc) { const hb_phase_t &b = ph[ord[mm_k].idx]; if (len > (int64_t)b.x_pos_e + 1 - b.x_pos_s) sf = true; }
						}
						if (sf) { mm_sc = sc; mm_k = s; mm_m = z.is_match; }
					}
				}
				if (mm_k != 0xffffffffu) { if (mm_k != keep) { const PhPair t = ord[mm_k]; ord[mm_k] = ord[keep]; ord[keep] = t; } keep++; }
				l = k;
			}
		}
	}
	return keep;
}
HB_HD uint32_t hb_ec_reverse_list(const DevReads &R, uint64_t qid, const hb_phase_t *ph, uint32_t n, PhPair *ord, const RsScratch &W, hb_ma_hit_t *out, int *ovf)
{
	const uint32_t keep = hb_ec_dedup(ph, n, ord, W, ovf);
	if (*ovf) return 0;
	uint32_t no = 0;
	for (uint32_t k = 0; k < keep; k++) {
		const hb_phase_t &z = ph[ord[k].idx];
		if (z.is_match != 2) continue;
		hb_ma_hit_t h; h.qns = (qid << 32) | z.x_pos_s; h.qe = z.x_pos_e + 1; h.tn = z.y_id; h.ts = z.y_pos_s; h.te = z.y_pos_e + 1; h.rev = z.rev;
		h.bl = R.len[z.y_id]; h.ml = (uint32_t)z.strong; h.no_l_indel = 0; h.del = 0; h.el = 0; for (int b = 0; b < 6; b++) h.pad[b] = 0;
		out[no++] = h;
	}
	return no;
}
