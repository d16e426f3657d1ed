// hb_kernels_ea.cuh — row a12, gen_hc_r_alin_ea (ecovlp.cpp:2810-2866), one WARP per overlap (included by engine.cu only).
// A chain whose target / strand has an exact (el) record in the read's overlap list of the previous round — the first such record in list order —
// with the same coordinates, and whose two substrings are still identical, is accepted without alignment.  The list scan is uniform over the
// warp (a few dozen records, broadcast loads); the compare of the two substrings (up to a whole read) is spread over the lanes, 16 bases per
// 32-bit word straight from the packed reads (reverse strand: reverse-complement of the mirrored word), as in k_exact.  Round 1 gave a thread
// to an overlap and compared serially: 0.8 s per 3 Gbp stage.
#pragma once
__global__ void __launch_bounds__(128) k_ec_ea_w(DevReads R, uint64_t r0, uint64_t n_ov, const OvDesc *__restrict__ desc, const hb_chain_t *__restrict__ ch,
                                                  const hb_ma_hit_t *__restrict__ prev, const uint64_t *__restrict__ prev_off, uint8_t *__restrict__ ea)
{
	const uint64_t o = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; const int lane = hb_lane(); if (o >= n_ov) return;
	const OvDesc d = desc[o]; const hb_chain_t c = ch[d.slot]; const uint64_t g = r0 + d.read;
	int cand = 0; uint32_t qs = 0, qe = 0, ts = 0, te = 0;
	for (uint64_t k0 = prev_off[g], k1 = prev_off[g + 1]; k0 < k1; k0 += 32) { // the first exact record of the (target, strand) pair: lanes test 32 records at a time
		const uint64_t k = k0 + lane; bool hit = false; hb_ma_hit_t p;
		if (k < k1) { p = prev[k]; hit = p.el && p.tn == c.y_id && (p.rev & 1) == c.y_pos_strand; }
		const unsigned m = __ballot_sync(HB_FULL, hit);
		if (m) {
			const int l0 = __ffs(m) - 1;
			int same = 0;
			if (lane == l0) { same = c.x_pos_s == (uint32_t)p.qns && c.x_pos_e + 1 == p.qe && c.y_pos_s == p.ts && c.y_pos_e + 1 == p.te; qs = (uint32_t)p.qns; qe = p.qe; ts = p.ts; te = p.te; }
			cand = __shfl_sync(HB_FULL, same, l0); qs = __shfl_sync(HB_FULL, qs, l0); qe = __shfl_sync(HB_FULL, qe, l0); ts = __shfl_sync(HB_FULL, ts, l0); te = __shfl_sync(HB_FULL, te, l0);
			break; // only the first exact record of the pair is looked at
		}
	}
	int ok = 0;
	if (cand) {
		const uint64_t qid = g, tid = c.y_id;
		if (R.noff[qid] != R.noff[qid + 1] || R.noff[tid] != R.noff[tid + 1]) { // N bases (rare): the per-base compare on one lane
			if (lane == 0) ok = hb_exact_seq(R, qid, qs, qe, tid, ts, te, (int)c.y_pos_strand);
			ok = __shfl_sync(HB_FULL, ok, 0);
		} else if (qe - qs == te - ts) {
			const uint8_t *q = R.packed + R.off[qid], *t = R.packed + R.off[tid];
			const uint64_t n = qe - qs, tl = R.len[tid], nw = n >> 4; ok = 1;
			for (uint64_t w = lane; w < nw && ok; w += 32) {
				const uint32_t a = ex_fwd16(q, qs + (w << 4)), b = c.y_pos_strand ? ex_rc16(ex_fwd16(t, tl - (ts + (w << 4)) - 16)) : ex_fwd16(t, ts + (w << 4));
				if (a != b) ok = 0;
			}
			for (uint64_t i = (nw << 4) + lane; i < n && ok; i += 32) {
				const int a = hb_base(q, qs + i), b = c.y_pos_strand ? 3 - hb_base(t, tl - 1 - (ts + i)) : hb_base(t, ts + i);
				if (a != b) ok = 0;
			}
			ok = __all_sync(HB_FULL, ok);
		}
	}
	if (lane == 0) ea[o] = (uint8_t)ok;
}
