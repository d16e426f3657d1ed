// hb_mwalign_w.cuh — the multi-word banded Myers aligner of step B (hb_mw_align, hb_ecaln.cuh) with the band spread over the lanes of a warp.
//
// ed_band_cal_{global,extension_0,extension_1,semi}_infi_w_trace (Levenshtein_distance.h:2516/2694/2823/3020) keep a band of 2*thre+1 bits in
// up to 64 words and update them word by word, column by column.  Here a lane owns a word (two words for bands of 33-64 words): a column is
//   * the match mask of the lane's 64 band positions straight from two bit planes of the pattern (low / high bit of the base + a "valid" plane
//     for positions outside the pattern or holding an N) — four logic ops instead of a shifted Peq table per base;
//   * the ripple add VP + (X & VP) across words as ONE carry-lookahead over two warp votes: with g = "the word's add overflows" and p = "the
//     word's sum is all ones", the carry into every word is ((G | P) + G) ^ P on the vote masks;
//   * the one-bit shifts of D0 and of the planes across words as one shuffle each;
//   * a trace row of 3 words per band word (D0, VP, VN; HP / HN follow from the row before) written side by side by the lanes.
// The traceback, the end-point scans and every scalar decision are the reference's, executed identically by all lanes (uniform loads).
// "Virtual lanes": the body is written over l = 0 .. 32*WPL-1; in the product lane L runs l = L and L + 32 (registers), ONLY in tests/hostemu
// a plain loop runs all l (arrays), so the logic is checked against the golden vectors without a GPU.
#pragma once
#include "hb_warp.cuh"
#include "hb_ecaln.cuh"

#if defined(__CUDA_ARCH__)
#define HB_VL_DECL(T, name) T name[WPL]
#define HB_VL_FOR(l) _Pragma("unroll") for (int s_ = 0; s_ < WPL; s_++) { const int l = hb_lane() + 32 * s_;
#define HB_VL(name) name[s_]
template <int WPL> HB_D uint64_t hb_vl_ballot(const uint32_t *p) { uint64_t m = 0; _Pragma("unroll") for (int s = 0; s < WPL; s++) m |= (uint64_t)__ballot_sync(0xffffffffu, p[s] != 0) << (32 * s); return m; }
// dst[l] = src[l + 1] (the last virtual lane receives 0)
template <int WPL> HB_D void hb_vl_down1(const uint32_t *src, uint32_t *dst)
{
	const int lane = hb_lane();
	_Pragma("unroll") for (int s = 0; s < WPL; s++) {
		const uint32_t a = __shfl_down_sync(0xffffffffu, src[s], 1);
		const uint32_t b = s + 1 < WPL ? __shfl_sync(0xffffffffu, src[s + 1 < WPL ? s + 1 : s], 0) : 0u;
		dst[s] = lane < 31 ? a : b;
	}
}
template <int WPL, typename T> HB_D T hb_vl_bcast(const T *src, int k) { T v = src[0]; _Pragma("unroll") for (int s = 1; s < WPL; s++) if ((k >> 5) == s) v = src[s]; return __shfl_sync(0xffffffffu, v, k & 31); }
template <int WPL> HB_D int32_t hb_vl_sum(const int32_t *src) { int32_t v = 0; _Pragma("unroll") for (int s = 0; s < WPL; s++) v += src[s]; return hb_wsum(v); }
#else
#define HB_VL_DECL(T, name) T name[64]
#define HB_VL_FOR(l) for (int l = 0; l < 32 * WPL; l++) {
#define HB_VL(name) name[l]
template <int WPL> inline uint64_t hb_vl_ballot(const uint32_t *p) { uint64_t m = 0; for (int l = 0; l < 32 * WPL; l++) if (p[l]) m |= 1ULL << l; return m; }
template <int WPL> inline void hb_vl_down1(const uint32_t *src, uint32_t *dst) { for (int l = 0; l < 32 * WPL; l++) dst[l] = l + 1 < 32 * WPL ? src[l + 1] : 0u; }
template <int WPL, typename T> inline T hb_vl_bcast(const T *src, int k) { return src[k]; }
template <int WPL> inline int32_t hb_vl_sum(const int32_t *src) { int32_t v = 0; for (int l = 0; l < 32 * WPL; l++) v += src[l]; return v; }
#endif
#define HB_VL_END }

HB_HD uint64_t hb_lsub_word(int32_t len, int l) { const int32_t r = len - 64 * l; return r >= 64 ? ~0ULL : (r > 0 ? (1ULL << r) - 1 : 0ULL); } // word l of "the low len bits set"

// same contract as hb_mw_align; trace layout ez.compact = 2: 2 * nword header words (initial VP | VN) + 3 * nword per column (D0 | VP | VN)
template <int WPL> HB_HD_NI void hb_mw_align_w(int mode, const RdView &T, int64_t ps0, int32_t pn, const RdView &Q, int64_t qs0, int32_t tn, int32_t thre, int32_t abs_diag, MwEz &ez)
{
	const int32_t bd0 = (thre << 1) + 1, nword = (bd0 >> 6) + ((bd0 & 63) ? 1 : 0), cut = thre + (thre << 1);
	int32_t i, err, i_bd, pidx = 0, tidx = 0, tmp_e = INT32_MAX, k, poff;
	ez.cn = 0; ez.thre = thre; ez.err = INT32_MAX; ez.pl = pn; ez.tl = tn;
	if (mode == 0) { ez.ps = ez.ts = 0; if (pn > tn + thre || tn > pn + thre) return; }
	else if (mode == 1) { ez.ps = ez.ts = 0; ez.pe = ez.te = -1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; }
	else if (mode == 2) { ez.ps = ez.ts = INT32_MAX; ez.pe = pn - 1; ez.te = tn - 1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; pidx = ez.pe; tidx = ez.te; }
	else { ez.ps = ez.pe = -1; ez.ts = 0; ez.te = tn - 1; if (pn > tn + cut || tn > pn + cut) return; }
	const int32_t tn0 = tn - 1, pe = pn - 1;
	ez.nword = nword;
	if ((uint64_t)nword * (2 + 3 * (uint64_t)tn) > ez.pcap || nword > ez.vstride || nword > 32 * WPL) { ez.ovf = 1; return; }
	ez.compact = 2;
	auto pch = [&](int32_t j) -> int { return T.at(ps0 + (mode == 2 ? pidx - j : j)); };
	auto tch = [&](int32_t j) -> int { return Q.at(qs0 + (mode == 2 ? tidx - j : j)); };
	const int32_t off = mode == 3 ? abs_diag : thre; // band bit b of column j looks at pattern position b - off + j
	HB_VL_DECL(uint64_t, VP); HB_VL_DECL(uint64_t, VN); HB_VL_DECL(uint64_t, lo); HB_VL_DECL(uint64_t, hi); HB_VL_DECL(uint64_t, va);
	HB_VL_DECL(uint64_t, X); HB_VL_DECL(uint64_t, S); HB_VL_DECL(uint64_t, D0); HB_VL_DECL(uint64_t, HP); HB_VL_DECL(uint64_t, HN);
	HB_VL_DECL(uint32_t, fg); HB_VL_DECL(uint32_t, fp); HB_VL_DECL(uint32_t, b0); HB_VL_DECL(uint32_t, nb); HB_VL_DECL(int32_t, ps);
	HB_VL_FOR(l)
		uint64_t a = 0, b = 0, v = 0;
		if (l < nword) for (int bit = 0; bit < 64; bit++) {
			const int32_t B = 64 * l + bit, p = B - off;
			if (B <= (thre << 1) && p >= 0 && p < pn) { const int c = pch(p); if (c < 4) { v |= 1ULL << bit; a |= (uint64_t)(c & 1) << bit; b |= (uint64_t)(c >> 1) << bit; } }
		}
		HB_VL(lo) = a; HB_VL(hi) = b; HB_VL(va) = v;
		if (mode == 3) { HB_VL(VP) = 0; HB_VL(VN) = l < nword ? hb_lsub_word(abs_diag, l) : 0; }
		else { HB_VL(VN) = l < nword ? hb_lsub_word(thre, l) : 0; HB_VL(VP) = l < nword ? (hb_lsub_word((thre << 1) + 1, l) ^ HB_VL(VN)) : 0; }
		if (l < nword) { ez.path[l] = HB_VL(VP); ez.path[nword + l] = HB_VL(VN); }
		HB_VL(D0) = HB_VL(HP) = HB_VL(HN) = 0;
	HB_VL_END
	if (mode == 3) { i_bd = (thre << 1) - abs_diag; err = abs_diag; } else { i_bd = thre; err = thre; }
	ez.pn = 2 * (uint64_t)nword;
	const int32_t Peq_i = (thre << 1) >> 6; const uint64_t Peq_m = 1ULL << ((thre << 1) & 63);
	for (i = 0; i <= tn0; i++) {
		const int tc = tch(i); const uint64_t Lm = (tc & 1) ? ~0ULL : 0ULL, Hm = (tc & 2) ? ~0ULL : 0ULL;
		HB_VL_FOR(l)
			const uint64_t m = tc < 4 ? (HB_VL(va) & ~(HB_VL(lo) ^ Lm) & ~(HB_VL(hi) ^ Hm)) : 0ULL, x = m | HB_VL(VN), vp = HB_VL(VP), s = (x & vp) + vp;
			HB_VL(X) = x; HB_VL(S) = s; HB_VL(fg) = s < vp; HB_VL(fp) = s == ~0ULL; (void)l;
		HB_VL_END
		const uint64_t Gm = hb_vl_ballot<WPL>(fg), Pm = hb_vl_ballot<WPL>(fp), Cm = ((Gm | Pm) + Gm) ^ Pm; // carry into every word
		HB_VL_FOR(l)
			const uint64_t vp = HB_VL(VP), s = HB_VL(S) + ((Cm >> l) & 1ULL), d0 = (s ^ vp) | HB_VL(X);
			HB_VL(D0) = d0; HB_VL(HN) = vp & d0; HB_VL(HP) = ~(vp | d0) | HB_VL(VN); HB_VL(b0) = l < nword ? (uint32_t)(d0 & 1ULL) : 0u;
		HB_VL_END
		hb_vl_down1<WPL>(b0, nb);
		HB_VL_FOR(l)
			if (l < nword) { const uint64_t x = (HB_VL(D0) >> 1) | ((uint64_t)HB_VL(nb) << 63); HB_VL(VN) = x & HB_VL(HP); HB_VL(VP) = ~(x | HB_VL(HP)) | HB_VL(HN); }
		HB_VL_END
		if (!hb_vl_bcast<WPL, uint32_t>(b0, 0)) { ++err; if (err > cut) return; }
		if (i < tn0) {
			if (mode == 1 || mode == 2) { // running best end point of an extension (Levenshtein_distance.h:2758-2779 / 2889-2910)
				poff = i - thre; k = i + thre - pe;
				if (k >= 0) {
					if (tmp_e == INT32_MAX) {
						const int32_t cnt = pe - poff; // bits 0 .. cnt-1 of VP / VN
						HB_VL_FOR(l) const uint64_t mk = l < nword ? hb_lsub_word(cnt, l) : 0ULL; HB_VL(ps) = hb_popc64(HB_VL(VP) & mk) - hb_popc64(HB_VL(VN) & mk); HB_VL_END
						tmp_e = err + hb_vl_sum<WPL>(ps);
					} else {
						k = (thre << 1) - k;
						if (k >= 0) { HB_VL_FOR(l) HB_VL(ps) = (int32_t)((HB_VL(HP) >> (k & 63)) & 1ULL) - (int32_t)((HB_VL(HN) >> (k & 63)) & 1ULL); (void)l; HB_VL_END tmp_e += hb_vl_bcast<WPL, int32_t>(ps, k >> 6); }
					}
					if (tmp_e <= ez.thre && tmp_e < ez.err) { ez.err = tmp_e; if (mode == 1) { ez.pe = pe; ez.te = i; } else { ez.ps = pidx - pe; ez.ts = tidx - i; } }
				}
			}
			HB_VL_FOR(l) HB_VL(b0) = l < nword ? (uint32_t)((HB_VL(lo) & 1ULL) | ((HB_VL(hi) & 1ULL) << 1) | ((HB_VL(va) & 1ULL) << 2)) : 0u; HB_VL_END
			hb_vl_down1<WPL>(b0, nb);
			++i_bd; const int c = i_bd < pn ? pch(i_bd) : 4;
			HB_VL_FOR(l)
				if (l < nword) {
					const uint64_t q = HB_VL(nb);
					HB_VL(lo) = (HB_VL(lo) >> 1) | ((q & 1ULL) << 63); HB_VL(hi) = (HB_VL(hi) >> 1) | (((q >> 1) & 1ULL) << 63); HB_VL(va) = (HB_VL(va) >> 1) | (((q >> 2) & 1ULL) << 63);
					if (l == Peq_i && c < 4) { HB_VL(va) |= Peq_m; if (c & 1) HB_VL(lo) |= Peq_m; if (c & 2) HB_VL(hi) |= Peq_m; }
				}
			HB_VL_END
		}
		uint64_t *o = ez.path + ez.pn;
		HB_VL_FOR(l) if (l < nword) { o[l] = HB_VL(D0); o[nword + l] = HB_VL(VP); o[2 * nword + l] = HB_VL(VN); } HB_VL_END
		ez.pn += 3 * (uint64_t)nword;
	}
	uint64_t *fVP = ez.vec, *fVN = ez.vec + ez.vstride; // the last column's VP / VN for the end-point scans below
	HB_VL_FOR(l) if (l < nword) { fVP[l] = HB_VL(VP); fVN[l] = HB_VL(VN); } HB_VL_END
	hb_wsync();
	if (mode == 0) {
		int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
		for (i = 0; site < ct; site++, i++) { err += hb_mw_bit(fVP, i); err -= hb_mw_bit(fVN, i); }
		if (site == ct && err <= thre) { ez.err = err; ez.pe = pn - 1; ez.te = tn - 1; }
		hb_mw_gen_trace(ez, thre, 1);
	} else if (mode == 1 || mode == 2) {
		int32_t site = tn - 1 - thre; const int32_t ct = pn - 1;
		for (i = 0; site < ct; i++) {
			err += hb_mw_bit(fVP, i); err -= hb_mw_bit(fVN, i); site++;
			if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
		}
		if (err <= thre && err < ez.err) { ez.err = err; if (mode == 1) { ez.pe = site; ez.te = tn - 1; } else { ez.ps = pidx - site; ez.ts = tidx + 1 - tn; } }
		if (mode == 1) hb_mw_gen_trace(ez, thre, 1);
		else {
			poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff;
			hb_mw_gen_trace(ez, thre, 0);
			poff = ez.ps; ez.ps = pidx - ez.pe; ez.pe = pidx - poff;
		}
	} else {
		int32_t site = tn - 1 - abs_diag, uge = INT32_MAX; const int32_t ai = pn - tn + abs_diag;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += hb_mw_bit(fVP, i); err -= hb_mw_bit(fVN, i); }
		if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site; }
		site -= i;
		while (i < ai) {
			err += hb_mw_bit(fVP, i); err -= hb_mw_bit(fVN, i); ++i;
			if (err <= thre && err <= ez.err) { ez.err = err; ez.pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == ez.err) ez.pe = site + thre;
		hb_mw_gen_trace(ez, abs_diag, 1);
	}
	hb_wsync();
}
