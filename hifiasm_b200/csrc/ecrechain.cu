// ecrechain.cu — k_ecb_rechain: the re-seeding rescue of step B (row a10, rechain_aln_hc Correct.cpp:17669) for the overlaps the merge
// kernel queued: one thread per overlap, grid-stride over the queue, with the largest aligner scratch (a window can be 10 kb wide with
// a 64-word band).  A rare path — a structural difference of >= 512 bp inside an accepted overlap — kept in its own translation unit.
#include "hb_rechain_launch.h"
#include "hb_ecrechain.cuh"

__global__ void __launch_bounds__(32) k_ecb_rechain(RcLaunch L)
{
	const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (uint64_t)gridDim.x * blockDim.x;
	const uint64_t n = *L.rc_n; if (tid >= n) return;
	const uint64_t cw = (uint64_t)L.cig_words;
	EcBCtx C; C.e_rate = L.e_rate; C.w_l = L.w_l; C.pool = L.pool; C.pool_used = L.pool_used; C.pool_cap = L.pool_cap; C.do_gaps = L.gaps; C.no_myers = 0;
	C.ez.path = L.path + tid * L.path_words; C.ez.pcap = L.path_words; C.ez.vec = L.vec + tid * 11 * (uint64_t)HB_MW_MAXW; C.ez.vstride = HB_MW_MAXW; C.ez.warp = 0;
	C.ez.cig = L.cig3 + tid * 3 * cw; C.ez.ccap = L.cig_words; C.wc = C.ez.cig + cw; C.wccap = L.cig_words; C.gout = C.ez.cig + 2 * cw; C.gcap = L.cig_words;
	EcRc S; S.R = L.R; S.zw = L.zw + tid * (uint64_t)L.zcap; S.zcap = L.zcap; S.poolA = L.poolA; S.zc = L.zc + tid * L.zc_cap; S.zc_used = L.zc_used + tid; S.zc_cap = L.zc_cap;
	S.path1 = L.path1 + tid * (uint64_t)L.w_l * 5; S.cig1 = L.cig1 + tid * HB_EC_CIG_TMP;
	S.h = L.h + tid * (uint64_t)L.hcap; S.hcap = L.hcap; S.hn = 0; S.t = L.t + tid * (uint64_t)L.hcap; S.p = L.p + tid * (uint64_t)L.hcap; S.f = L.f + tid * (uint64_t)L.hcap;
	S.rs.bb = L.rs_b + tid * 512; S.rs.be = S.rs.bb + 256; S.rs.st = (RsFrame *)L.rs_f + tid * HB_RS_STACK;
	S.pen_gap = L.pen_gap; S.pen_skip = L.pen_skip; S.h_khit = HB_E_KHIT; S.err = L.err; S.ovf = 0;
	for (uint64_t wk = tid; wk < n; wk += nthr) {
		const uint64_t o = L.rc_q[wk]; const hb_aln_t a = L.aln[o]; const OvDesc d = L.desc[o]; const hb_chain_t c = L.ch[d.slot];
		EcZ z; z.x_pos_s = c.x_pos_s; z.x_pos_e = c.x_pos_e; z.y_pos_s = c.y_pos_s; z.y_id = c.y_id; z.rev = c.y_pos_strand;
		z.fc = L.fc + L.fc_grp_base[d.slot] + c.fc_off; z.fc_n = c.fc_n; z.align_length = a.align_length; z.w = (hb_wl_t *)(L.wlA + a.w_off); z.wn = (int32_t)a.w_n;
		C.q = hb_rd_view(L.R, L.r0 + d.read, 0); C.t = hb_rd_view(L.R, c.y_id, c.y_pos_strand); C.ql = C.q.len; C.tl = C.t.len;
		C.aw = L.wl + L.wl_off[o]; C.awcap = (int32_t)(L.wl_off[o + 1] - L.wl_off[o]);
		hb_alnb_t r = L.out[o];
		hb_ecb_rechain(C, z, a.re, S, &r);
		if (r.st == -2) atomicOr(L.err, 32);
		L.out[o] = r;
	}
}

size_t hb_rechain_scratch_bytes(int32_t w_l, int32_t cig_words, int32_t zcap, uint64_t zc_cap, int32_t hcap)
{
	return (size_t)HB_MW_MAXW * HB_MAX_SIN_L * 5 * 8 + 11 * (size_t)HB_MW_MAXW * 8 + 3 * (size_t)cig_words * 2 + (size_t)zcap * sizeof(hb_wl_t) + zc_cap * 2 + 8 +
	       (size_t)w_l * 5 * 8 + HB_EC_CIG_TMP * 2 + (size_t)hcap * (sizeof(hb_hit_t) + 8 + 8 + 4) + 512 * 4 + HB_RS_STACK * sizeof(RsFrame) + 256;
}
cudaError_t hb_launch_ecb_rechain(const RcLaunch &L, cudaStream_t stream)
{
	k_ecb_rechain<<<L.blocks, 32, 0, stream>>>(L);
	return cudaGetLastError();
}
