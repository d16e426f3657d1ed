// oracle/refdump.cpp — TEST INFRASTRUCTURE ONLY.
//
// Our own driver, linked against the UNMODIFIED reference objects
// (oracle/_ref/libhifiasm_ref.so, built by oracle/Makefile from the sources
// under /root/reference).  It replays the reference's overlap/EC stage the way
// ha_assemble() does (Assembly.cpp:2055-2113) and dumps
//   * per-read intermediate arrays of the hot path (sketch, index probe,
//     anchors, chains) taken by calling the reference's own non-static
//     functions (mz1_ha_sketch htab.h:121, ha_pt_get htab.h:85,
//     minimizers_qgen0 anchor.cpp:987, h_ec_lchain anchor.cpp:2302), and
//   * the read store and overlap lists before / after the final overlap pass
//     (ha_ec_ff -> cal_ov_r, Assembly.cpp:1942 / ecovlp.cpp:6385) through the
//     reference's own writers (write_All_reads Process_Read.cpp:69,
//     write_ma_hit_ts Overlaps.cpp:23442).
// The dumps are the golden vectors the oracle port and the CUDA path are pinned
// against (tests/golden/, made by tests/golden/make_golden.py).
//
// usage: refdump <raw|final|roundK> <out_prefix> <hifiasm args...>
//   roundK: EC rounds 0..K; of round K the states before / inside / after cal_ec_r (see the wrappers below)
//   raw   : filter table + round-0 index on the raw reads, stage dumps, exit
//   final : 3 EC rounds, "<p>.pre.*" bins, final index, stage dumps,
//           cal_ov_r, "<p>.fin.*" bins
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <vector>
#include <float.h>
#include "CommandLines.h"
#include "Process_Read.h"
#include "Overlaps.h"
#include "Assembly.h"
#include "htab.h"
#include "Hash_Table.h"
#include "ecovlp.h"
#include "Levenshtein_distance.h"
#include "Correct.h"
#include <dlfcn.h>

// non-static reference functions that no header declares
void ha_ec(int64_t round, int num_pround, int des_idx, uint64_t *tot_b, uint64_t *tot_e);
void ha_ec_ff(int renew_idx); // Assembly.cpp:1942
void ha_opt_update_cov(hifiasm_opt_t *opt, int hom_cov);
void ha_opt_reset_to_round(hifiasm_opt_t *asm_opt, int round);
void write_ma_hit_ts(ma_hit_t_alloc *x, long long n_read, char *read_file_name);
void minimizers_qgen0(ha_abuf_t *ab, char *rs, int64_t rl, uint64_t mz_w, uint64_t mz_k, Candidates_list *cl, kvec_t_u8_warp *k_flag,
                      void *ha_flt_tab, ha_pt_t *ha_idx, All_reads *rdb, kvec_t_u64_warp *dbg_ct, st_mt_t *sp, uint32_t *high_occ, uint32_t *low_occ);
void h_ec_lchain(ha_abuf_t *ab, uint32_t rid, char *rs, uint64_t rl, uint64_t mz_w, uint64_t mz_k, All_reads *rref, overlap_region_alloc *overlap_list, Candidates_list *cl, double bw_thres,
                 int max_n_chain, int apend_be, kvec_t_u8_warp *k_flag, kvec_t_u64_warp *dbg_ct, st_mt_t *sp, uint32_t *high_occ, uint32_t *low_occ, uint32_t is_accurate, uint32_t gen_off, int64_t mcopy_num, double mcopy_rate, uint32_t chain_cutoff, uint32_t mcopy_khit_cut, uint64_t ocv_w);

int32_t init_waln(int64_t err, int64_t s, int64_t l, int64_t w_l, int64_t *aux_beg, int64_t *aux_end, int64_t *r_s, int64_t *r_l); // Correct.cpp:764
int64_t get_num_wins(int64_t s, int64_t e, int64_t block_s); // Correct.cpp:783
uint32_t align_hc_ed_post_extz(overlap_region *z, All_reads *rref, char *qstr, char *tstr, bit_extz_t *exz, double e_rate, int64_t w_l, double ovlp_cut, int64_t force_aln, void *km); // Correct.cpp:12951
double gen_extend_err_exz(overlap_region *z, const ul_idx_t *uref, hpc_t *hpc_g, All_reads *rref, char *qstr, char *tstr, bit_extz_t *exz, uint64_t *v_idx, int64_t block_s, double ovlp_cut, double e_rate, double e_max, int64_t max_err, int64_t sec_check, int64_t *r_e); // Correct.cpp:13400
uint64_t gen_hc_fast_cigar(overlap_region *z, Candidates_list *cl, All_reads *rref, int64_t wl, char *qstr, UC_Read *tu, bit_extz_t *exz, overlap_region *aux_o, double e_rate, int64_t ql, int64_t rid, int64_t khit, int64_t *re); // Correct.cpp:25137
void reassign_gaps(overlap_region *z, overlap_region *aux_o, char *qstr, int64_t ql, char *tstr, int64_t tl, All_reads *rref, UC_Read *tu, asg16_v *buf); // Correct.cpp:25409
overlap_region *fetch_aux_ovlp(overlap_region_alloc *ol); // ecovlp.cpp:257
void dedup_chains(overlap_region_alloc *ol); // ecovlp.cpp:2984
void push_ne_ovlp(ma_hit_t_alloc *paf, overlap_region_alloc *ov, uint32_t flag, All_reads *R_INF, asg16_v *ec); // ecovlp.cpp:2585
void gen_hc_r_alin_ea(overlap_region_alloc *ol, Candidates_list *cl, All_reads *rref, UC_Read *qu, UC_Read *tu, bit_extz_t *exz, overlap_region *aux_o, double e_rate, int64_t wl, int64_t rid, int64_t khit, int64_t move_gap,
                      asg16_v *buf, asg64_v *srt, ma_hit_t_alloc *in, uint8_t chem_drop, double align_gap_rate, int64_t align_gap_max); // ecovlp.cpp:2810

#define HA_KMER_GOOD_RATIO 0.333 /* ecovlp.cpp:9 */
#define COV_W 3072               /* ecovlp.cpp:17 */

static FILE *xopen(const char *pfx, const char *suf)
{
	char *fn = (char *)malloc(strlen(pfx) + strlen(suf) + 2);
	sprintf(fn, "%s%s", pfx, suf);
	FILE *fp = fopen(fn, "wb");
	if (!fp) { fprintf(stderr, "refdump: cannot write %s\n", fn); exit(1); }
	free(fn);
	return fp;
}

// one record per chain: the overlap_region fields the hot path defines
typedef struct {
	uint32_t x_pos_s, x_pos_e, y_id, y_pos_s, y_pos_e, y_pos_strand;
	int32_t shared_seed;
	uint32_t first_hit; // overlap_region.non_homopolymer_errors (Hash_Table.cpp:2280)
	uint32_t n_fc;      // f_cigar.length
} chain_rec_t;

// one overlap's window list (Hash_Table.h:54-69): n, c.n, window_list[n] (32 B each), cigar pool u16[c.n]
static void dump_wl(FILE *fp, const overlap_region *z)
{
	uint32_t h[2] = { (uint32_t)z->w_list.n, (uint32_t)z->w_list.c.n };
	fwrite(h, 4, 2, fp);
	fwrite(z->w_list.a, sizeof(window_list), z->w_list.n, fp);
	fwrite(z->w_list.c.a, 2, z->w_list.c.n, fp);
}

static void dump_stages(const char *pfx, double bw_thres)
{
	uint64_t i, j, n_reads = R_INF.total_reads;
	uint32_t high_occ = asm_opt.hom_cov * (2.0 - HA_KMER_GOOD_RATIO); // ecovlp.cpp:3952
	uint32_t low_occ = asm_opt.hom_cov * HA_KMER_GOOD_RATIO;
	FILE *fmz = xopen(pfx, ".mz.bin"), *fidx = xopen(pfx, ".idx.bin"), *fan = xopen(pfx, ".anchors.bin");
	FILE *fch = xopen(pfx, ".chains.bin"), *fpa = xopen(pfx, ".params.txt"), *fwn = xopen(pfx, ".windows.bin");
	FILE *fal = xopen(pfx, ".aln.bin"), *fph = xopen(pfx, ".phase.bin"), *fea = xopen(pfx, ".ea.bin"); asg16_v v16; memset(&v16, 0, sizeof(v16));
	haplotype_evdience_alloc hap; InitHaplotypeEvdience(&hap); kv_ul_ov_t pidx; memset(&pidx, 0, sizeof(pidx)); asg64_v v64, buf0; memset(&v64, 0, sizeof(v64)); memset(&buf0, 0, sizeof(buf0));
	UC_Read tr; init_UC_Read(&tr); bit_extz_t exz; init_bit_extz_t(&exz, 31);
	const double e_rate = asm_opt.max_ov_diff_ec; const int64_t w_l = asm_opt.is_ont ? WINDOW_OHC : WINDOW_HC; // ecovlp.cpp:3288
	UC_Read ur; init_UC_Read(&ur);
	ha_abuf_t *ab = ha_abuf_init();
	Candidates_list cl; init_Candidates_list(&cl);
	overlap_region_alloc ol; init_overlap_region_alloc(&ol);
	st_mt_t sp; memset(&sp, 0, sizeof(sp));
	ha_mz1_v mz; memset(&mz, 0, sizeof(mz));

	fprintf(fpa, "n_reads %lu\nk %d\nw %d\nhom_cov %d\nhet_cov %d\nmax_n_chain %d\nhigh_occ %u\nlow_occ %u\nbw_thres %.6f\nmz_sample_dist %d\nmz_rewin %d\nis_hpc %d\nhave_flt %d\n",
	        (unsigned long)n_reads, asm_opt.k_mer_length, asm_opt.mz_win, asm_opt.hom_cov, asm_opt.het_cov, asm_opt.max_n_chain,
	        high_occ, low_occ, bw_thres, asm_opt.mz_sample_dist, asm_opt.mz_rewin, !(asm_opt.flag & HA_F_NO_HPC), ha_flt_tab ? 1 : 0);
	fclose(fpa);

	for (i = 0; i < n_reads; i++) {
		recover_UC_Read(&ur, &R_INF, i);
		// (1) sketch, with the arguments minimizers_qgen0 passes (anchor.cpp:1003)
		mz.n = 0;
		mz1_ha_sketch(ur.seq, ur.length, asm_opt.mz_win, asm_opt.k_mer_length, 0, !(asm_opt.flag & HA_F_NO_HPC), &mz, ha_flt_tab,
		              asm_opt.mz_sample_dist, NULL, NULL, NULL, -1, asm_opt.dp_min_len, -1, &sp, asm_opt.mz_rewin, 0, NULL);
		uint32_t n = mz.n;
		fwrite(&n, 4, 1, fmz);
		fwrite(mz.a, sizeof(ha_mz1_t), n, fmz);
		// (2) index probe of every query minimizer
		for (j = 0; j < n; j++) {
			int c; const ha_idxpos_t *p = ha_pt_get(ha_idx, mz.a[j].x, &c);
			uint32_t cc = c;
			fwrite(&mz.a[j].x, 8, 1, fidx); fwrite(&cc, 4, 1, fidx);
			if (c) fwrite(p, sizeof(ha_idxpos_t), c, fidx);
		}
		// (3) anchors after sort + weighting (cl->list of minimizers_qgen0)
		minimizers_qgen0(ab, ur.seq, ur.length, asm_opt.mz_win, asm_opt.k_mer_length, &cl, NULL, ha_flt_tab, ha_idx, &R_INF, NULL, &sp, &high_occ, &low_occ);
		uint64_t na = cl.length;
		fwrite(&na, 8, 1, fan);
		fwrite(cl.list, sizeof(k_mer_hit), na, fan);
		// (4) chains: the call worker_hap_dc_ec_gen_new_idx makes (ecovlp.cpp:3957) but with the mode's bw
		h_ec_lchain(ab, i, ur.seq, ur.length, asm_opt.mz_win, asm_opt.k_mer_length, &R_INF, &ol, &cl, bw_thres, asm_opt.max_n_chain, 1, NULL, NULL, &sp, &high_occ, &low_occ, 1, 1, 3, 0.7, 2, 32, COV_W);
		uint32_t nc = ol.length; uint64_t nh = cl.length;
		fwrite(&nc, 4, 1, fch); fwrite(&nh, 8, 1, fch);
		for (j = 0; j < nc; j++) {
			overlap_region *o = &ol.list[j]; chain_rec_t r;
			r.x_pos_s = o->x_pos_s; r.x_pos_e = o->x_pos_e; r.y_id = o->y_id; r.y_pos_s = o->y_pos_s; r.y_pos_e = o->y_pos_e;
			r.y_pos_strand = o->y_pos_strand; r.shared_seed = o->shared_seed; r.first_hit = o->non_homopolymer_errors; r.n_fc = o->f_cigar.length;
			fwrite(&r, sizeof(r), 1, fch);
			fwrite(o->f_cigar.buffer, 8, o->f_cigar.length, fch);
		}
		fwrite(cl.list, sizeof(k_mer_hit), nh, fch); // compacted chain hits (des region)
		// (5) window pass: every 775-bp query window of every chain, exactly the per-window work of
		// align_hc_ed_post_extz (Correct.cpp:12951-13011) without its early exit / gap filling:
		// {chain, q_s, q_e, t_s, t_pri_l, thre, aux_beg, aux_end, err, pe}, err = INT32_MAX when unaligned,
		// t_pri_l = -1 when init_waln rejects the window
		{
			std::vector<int32_t> wr;
			for (j = 0; j < nc; j++) {
				overlap_region *z = &ol.list[j];
				int64_t q_s, q_e, nw, k, q_l, t_tot_l, aux_beg, aux_end, t_s, thre, aln_l, t_pri_l;
				nw = get_num_wins(z->x_pos_s, z->x_pos_e + 1, w_l);
				get_win_se_by_normalize_xs(z, (z->x_pos_s / w_l) * w_l, w_l, &q_s, &q_e);
				for (k = 0; k < nw; k++) {
					aux_beg = aux_end = 0; q_l = 1 + q_e - q_s;
					thre = q_l * e_rate; thre = Adjust_Threshold(thre, q_l);
					if (thre > THRESHOLD_MAX_SIZE) thre = THRESHOLD_MAX_SIZE;
					t_s = (q_s - z->x_pos_s) + z->y_pos_s;
					t_s += y_start_offset(q_s, &(z->f_cigar));
					aln_l = q_l + (thre << 1); t_tot_l = Get_READ_LENGTH(R_INF, z->y_id);
					int32_t rec[10] = { (int32_t)j, (int32_t)q_s, (int32_t)q_e, 0, -1, (int32_t)thre, 0, 0, INT32_MAX, -1 };
					if (init_waln(thre, t_s, t_tot_l, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) {
						resize_UC_Read(&tr, t_pri_l + 8);
						recover_UC_Read_sub_region(tr.seq, t_s, t_pri_l, z->y_pos_strand, &R_INF, z->y_id);
						ed_band_cal_semi_64_w_absent_diag(tr.seq, t_pri_l, ur.seq + q_s, q_l, thre, aux_beg, &exz);
						rec[3] = (int32_t)t_s; rec[4] = (int32_t)t_pri_l; rec[6] = (int32_t)aux_beg; rec[7] = (int32_t)aux_end; rec[8] = exz.err; rec[9] = exz.pe;
					} else { rec[3] = (int32_t)t_s; }
					wr.insert(wr.end(), rec, rec + 10);
					q_s = q_e + 1; q_e = q_s + w_l - 1;
					if (q_e >= (int64_t)z->x_pos_e) q_e = z->x_pos_e;
				}
			}
			uint32_t nwr = wr.size() / 10;
			fwrite(&nwr, 4, 1, fwn); fwrite(wr.data(), 4, wr.size(), fwn);
		}
		// (6) the alignment stage of an EC round, overlap by overlap, in the order and with the arguments of
		// gen_hc_r_alin (Correct.cpp:25617-25675), the state dumped after each step:
		//   A  align_hc_ed_post_extz (12951) + gen_extend_err_exz (13400): ok, align_length, rr, re, window list
		//   B  gen_hc_fast_cigar (25137): re, window list (base-level cigars)
		//   C  reassign_gaps (25409): window list
		{
			overlap_region *aux_o = fetch_aux_ovlp(&ol); // ecovlp.cpp:3279; may move ol.list
			const double err = e_rate, e_max = err * 1.5; const int64_t ql = ur.length;
			resize_UC_Read(&tr, ((w_l + (THRESHOLD_MAX_SIZE << 1) + 1) << 1) + 8);
			fwrite(&nc, 4, 1, fal);
			uint64_t kacc = 0; // gen_hc_r_alin keeps the accepted overlaps at the front of the list, in order (Correct.cpp:25665-25671)
			for (j = 0; j < nc; j++) {
				overlap_region *z = &ol.list[j]; int64_t re = INT64_MAX; double rr = DBL_MAX;
				z->shared_seed = z->non_homopolymer_errors;
				int32_t ok = align_hc_ed_post_extz(z, &R_INF, ur.seq, tr.seq, &exz, err, w_l, OVERLAP_THRESHOLD_HIFI_FILTER, 0, NULL);
				if (ok) rr = gen_extend_err_exz(z, NULL, NULL, &R_INF, ur.seq, tr.seq, &exz, NULL, w_l, -1, err, (e_max + 0.000001), THRESHOLD_MAX_SIZE, 0, &re);
				int32_t st = !ok ? 0 : (rr > err ? 1 : 2); uint32_t al = z->align_length;
				fwrite(&st, 4, 1, fal); fwrite(&al, 4, 1, fal); fwrite(&rr, 8, 1, fal); fwrite(&re, 8, 1, fal);
				dump_wl(fal, z);
				z->is_match = 0;
				if (st != 2) continue;
				z->non_homopolymer_errors = re;
				gen_hc_fast_cigar(z, &cl, &R_INF, w_l, ur.seq, &tr, &exz, aux_o, e_rate, ql, i, 31 /* E_KHIT, ecovlp.cpp */, &re);
				fwrite(&re, 8, 1, fal); dump_wl(fal, z);
				reassign_gaps(z, aux_o, ur.seq, ql, NULL, -1, &R_INF, &tr, &v16);
				dump_wl(fal, z);
				{ int64_t nhe = z->non_homopolymer_errors; uint32_t xy[4] = { z->x_pos_s, z->x_pos_e, z->y_pos_s, z->y_pos_e }; fwrite(&nhe, 8, 1, fal); fwrite(xy, 4, 4, fal); }
				if (kacc != j) { overlap_region t = ol.list[kacc]; ol.list[kacc] = ol.list[j]; ol.list[j] = t; }
				z = &ol.list[kacc++]; z->is_match = 1; z->strong = z->without_large_indel = 0;
			}
			// (7) phasing and de-duplication of the accepted overlaps, as worker_hap_ec does next (ecovlp.cpp:3299-3306):
			// rphase_hc (Correct.cpp:20191) marks overlaps that carry informative minor alleles is_match = 2; dedup_chains
			// (ecovlp.cpp:2984) keeps the best chain per target.  Dumped per overlap, in list order:
			// {y_id, strand, x_pos_s, x_pos_e, y_pos_s, y_pos_e, non_homopolymer_errors, is_match, strong}
			ol.length = kacc; ol.mapped_overlaps_length = 0;
			rphase_hc(&ol, &R_INF, &hap, &ur, &tr, &pidx, &v64, &buf0, 0, WINDOW_MAX_SIZE, ur.length, 1, i, 0, 0, NULL, NULL, NULL, 0);
			for (int rep = 0; rep < 2; rep++) {
				uint32_t n = ol.length; fwrite(&n, 4, 1, fph);
				for (j = 0; j < ol.length; j++) {
					overlap_region *z = &ol.list[j];
					uint32_t r9[9] = { z->y_id, z->y_pos_strand, z->x_pos_s, z->x_pos_e, z->y_pos_s, z->y_pos_e, z->non_homopolymer_errors, z->is_match, (uint32_t)(int32_t)z->strong };
					fwrite(r9, 4, 9, fph);
				}
				if (rep == 0) dedup_chains(&ol);
			}
			// the round's reverse_paf[i] as worker_hap_ec emits it (ecovlp.cpp:3321): push_ne_ovlp(flag = 2, ec = NULL) over the de-duplicated
			// list; without_large_indel of these overlaps is 0 (wcns_gen only sets it for is_match == 1, ecovlp.cpp:2302-2303).
			// Dumped fields: qns, qe, tn, ts, te, rev, bl, ml, no_l_indel (el / del are never written on this path)
			{
				static ma_hit_t_alloc tmp; uint32_t k;
				for (j = 0; j < ol.length; j++) ol.list[j].without_large_indel = 0;
				push_ne_ovlp(&tmp, &ol, 2, &R_INF, NULL);
				uint32_t n = tmp.length; fwrite(&n, 4, 1, fph);
				for (k = 0; k < n; k++) {
					ma_hit_t *h = &tmp.buffer[k]; uint64_t q = h->qns; uint32_t r8[8] = { h->qe, h->tn, h->ts, h->te, (uint32_t)h->rev, h->bl, (uint32_t)h->ml, (uint32_t)h->no_l_indel };
					fwrite(&q, 8, 1, fph); fwrite(r8, 4, 8, fph);
				}
			}
		}
		// (8) the whole alignment stage as worker_hap_ec calls it (ecovlp.cpp:3288): gen_hc_r_alin_ea with the read's overlap list of the
		// previous round — overlaps whose previous record was exact (el) and still is are accepted without alignment (row a12).
		// Chains are rebuilt first (steps 6-7 consumed them).  Per accepted overlap, in list order:
		// {y_id, strand, x_pos_s, x_pos_e, y_pos_s, y_pos_e, non_homopolymer_errors, is_match} + window list
		{
			h_ec_lchain(ab, i, ur.seq, ur.length, asm_opt.mz_win, asm_opt.k_mer_length, &R_INF, &ol, &cl, bw_thres, asm_opt.max_n_chain, 1, NULL, NULL, &sp, &high_occ, &low_occ, 1, 1, 3, 0.7, 2, 32, COV_W);
			overlap_region *aux_o = fetch_aux_ovlp(&ol);
			gen_hc_r_alin_ea(&ol, &cl, &R_INF, &ur, &tr, &exz, aux_o, e_rate, w_l, i, 31, 1, &v16, &v64, &(R_INF.paf[i]), 0, -1, -1);
			uint32_t n = ol.length; fwrite(&n, 4, 1, fea);
			for (j = 0; j < ol.length; j++) {
				overlap_region *z = &ol.list[j];
				uint32_t r8[8] = { z->y_id, z->y_pos_strand, z->x_pos_s, z->x_pos_e, z->y_pos_s, z->y_pos_e, z->non_homopolymer_errors, z->is_match };
				fwrite(r8, 4, 8, fea); dump_wl(fea, z);
			}
		}
	}
	fclose(fmz); fclose(fidx); fclose(fan); fclose(fch); fclose(fwn); fclose(fal); fclose(fph); fclose(fea);
	destory_UC_Read(&ur);
}


// ---------------------------------------------------------------------------------------------------------------------
// `round` mode: the states an error-correction round (cal_ec_r, ecovlp.cpp:6268) passes through, taken from the real run.
// libhifiasm_ref.so is the unmodified reference compiled -fPIC, so its calls to its own global functions bind through the
// PLT; defining sl_ec_r / cal_update_ec_multiple / cns_gen_full here (same signatures) puts this file between cal_ec_r and
// them.  Each wrapper dumps R_INF with the reference's own writers and forwards to the reference's function (RTLD_NEXT).
//   <p>.hap.*  after cal_ec_multiple = kt_for(worker_hap_ec) (ecovlp.cpp:6063): paf / reverse_paf as push_ne_ovlp left them
//              (3320-3321), is_fully_corrected / is_abnormal (check_well_cal 2750), the edit scripts scc.a[i] (push_nec_re 2426)
//   <p>.sl.ec  after sl_ec_r (6402, worker_sl_ec 5965): the corrected reads
//   <p>.upd.*  after cal_update_ec_multiple (6095, worker_update_dc_ec 3808): paf with the exact intervals remapped
//   <p>.post.* after cal_ec_r: reads reverse-complemented and lists flipped when the round does that (worker_hap_post_rev 3866)
typedef struct { size_t n, m; asg16_v *a; uint8_t *f; } ref_cc_v; // layout of cc_v, ecovlp.cpp:100
extern ref_cc_v scc;                                              // ecovlp.cpp:101
struct ec_ovec_buf_t; struct cc_idx_t; struct cns_gfa;
static const char *g_hook_pfx = NULL;
static uint64_t g_full_calls = 0, g_full_bases = 0;
static void write_tag(const char *tag, int reads, int lists)
{
	char *fn = (char *)malloc(strlen(g_hook_pfx) + strlen(tag) + 32);
	if (reads) { sprintf(fn, "%s.%s.ec", g_hook_pfx, tag); write_All_reads(&R_INF, fn); }
	if (lists) {
		sprintf(fn, "%s.%s.ovlp.source", g_hook_pfx, tag); write_ma_hit_ts(R_INF.paf, R_INF.total_reads, fn);
		sprintf(fn, "%s.%s.ovlp.reverse", g_hook_pfx, tag); write_ma_hit_ts(R_INF.reverse_paf, R_INF.total_reads, fn);
	}
	free(fn);
}
void sl_ec_r(uint64_t n_thre, uint64_t n_a)
{
	static void (*real)(uint64_t, uint64_t) = (void (*)(uint64_t, uint64_t))dlsym(RTLD_NEXT, "_Z7sl_ec_rmm");
	if (g_hook_pfx) {
		write_tag("hap", 0, 1);
		FILE *fp = xopen(g_hook_pfx, ".hap.scc.bin"); uint64_t i;
		for (i = 0; i < n_a; i++) { uint32_t n = scc.a[i].n; fwrite(&n, 4, 1, fp); fwrite(scc.a[i].a, 2, n, fp); }
		fclose(fp);
	}
	real(n_thre, n_a);
	if (g_hook_pfx) write_tag("sl", 1, 0);
}
void cal_update_ec_multiple(ec_ovec_buf_t *b, uint64_t n_thre, uint64_t n_a)
{
	static void (*real)(ec_ovec_buf_t *, uint64_t, uint64_t) = (void (*)(ec_ovec_buf_t *, uint64_t, uint64_t))dlsym(RTLD_NEXT, "_Z22cal_update_ec_multipleP13ec_ovec_buf_tmm");
	real(b, n_thre, n_a);
	if (g_hook_pfx) write_tag("upd", 0, 1);
}
// how often the graph consensus (cns_gen_full, ecovlp.cpp:1919) runs, and over how many query bases
uint64_t cns_gen_full(overlap_region *ol, All_reads *rref, uint64_t s0, uint64_t e0, uint64_t wl, char *qstr, UC_Read *tu, bit_extz_t *exz, cc_idx_t *idx, uint64_t occ_tot, double occ_max,
                      asg32_v *b32, cns_gfa *cns, uint64_t max_trace, window_list *ridx, window_list_alloc *res, uint32_t rid)
{
	typedef uint64_t (*fn_t)(overlap_region *, All_reads *, uint64_t, uint64_t, uint64_t, char *, UC_Read *, bit_extz_t *, cc_idx_t *, uint64_t, double, asg32_v *, cns_gfa *, uint64_t, window_list *, window_list_alloc *, uint32_t);
	static fn_t real = (fn_t)dlsym(RTLD_NEXT, "_Z12cns_gen_fullP14overlap_regionP9All_readsmmmPcP7UC_ReadP10bit_extz_tP8cc_idx_tmdP7asg32_vP7cns_gfamP11window_listP17window_list_allocj");
	__sync_fetch_and_add(&g_full_calls, 1); __sync_fetch_and_add(&g_full_bases, e0 - s0);
	if (getenv("REFDUMP_FULL_TRACE")) fprintf(stderr, "[refdump] cns_gen_full rid %u [%lu, %lu)\n", rid, (unsigned long)s0, (unsigned long)e0);
	return real(ol, rref, s0, e0, wl, qstr, tu, exz, idx, occ_tot, occ_max, b32, cns, max_trace, ridx, res, rid);
}

static void write_bins(const char *pfx, const char *tag)
{
	char *fn = (char *)malloc(strlen(pfx) + strlen(tag) + 32);
	sprintf(fn, "%s.%s.ec", pfx, tag); write_All_reads(&R_INF, fn);
	sprintf(fn, "%s.%s.ovlp.source", pfx, tag); write_ma_hit_ts(R_INF.paf, R_INF.total_reads, fn);
	sprintf(fn, "%s.%s.ovlp.reverse", pfx, tag); write_ma_hit_ts(R_INF.reverse_paf, R_INF.total_reads, fn);
	free(fn);
}

// refdump edsemi <cases.bin> <out.bin>: runs the reference's window aligner
// ed_band_cal_semi_64_w_absent_diag (Levenshtein_distance.h:3727) on a file of
// cases {i32 pn, tn, thre, abs_diag; char p[pn], t[tn]} -> {i32 err, pe}
static int run_edsemi(const char *in, const char *out)
{
	FILE *fi = fopen(in, "rb"), *fo = fopen(out, "wb"); int32_t n, k, h[4];
	if (!fi || !fo) return 1;
	bit_extz_t ez; init_bit_extz_t(&ez, 31);
	if (fread(&n, 4, 1, fi) != 1) return 1;
	for (k = 0; k < n; k++) {
		if (fread(h, 4, 4, fi) != 4) return 1;
		char *p = (char *)malloc(h[0] + 1), *t = (char *)malloc(h[1] + 1);
		if (fread(p, 1, h[0], fi) != (size_t)h[0] || fread(t, 1, h[1], fi) != (size_t)h[1]) return 1;
		ed_band_cal_semi_64_w_absent_diag(p, h[0], t, h[1], h[2], h[3], &ez);
		int32_t r[2] = { ez.err, ez.pe };
		fwrite(r, 4, 2, fo);
		free(p); free(t);
	}
	fclose(fi); fclose(fo);
	return 0;
}

// refdump bench <n_query> <warmup> <steps> <hifiasm args...>: the reference's own final overlap
// pass, timed.  Builds the filter table and the index like ha_assemble() does
// (Assembly.cpp:2083, 1007), then runs cal_ov_r (ecovlp.cpp:6385) over the first
// n_query reads (0 = all) with empty previous overlap lists and prints one JSON
// line: seconds (wall of cal_ov_r alone), query bases, threads.
static int run_bench(int argc, char *argv[])
{
	uint64_t nq = strtoull(argv[2], 0, 10), i, bases = 0; int hom_cov = -1, het_cov = -1, warm = atoi(argv[3]), steps = atoi(argv[4]), s;
	yak_reset_realtime();
	init_opt(&asm_opt);
	argv[4] = argv[0];
	if (!CommandLine_process(argc - 4, argv + 4, &asm_opt)) return 1;
	ha_flt_tab = NULL; ha_idx = NULL;
	double t0 = yak_realtime();
	if (!(asm_opt.flag & HA_F_NO_KMER_FLT)) { ha_flt_tab = ha_ft_gen(&asm_opt, &R_INF, &hom_cov, 0, 0); ha_opt_update_cov(&asm_opt, hom_cov); }
	ha_idx = ha_pt_gen(&asm_opt, ha_flt_tab, 0, 0, &R_INF, &hom_cov, &het_cov);
	asm_opt.hom_cov = hom_cov; asm_opt.het_cov = het_cov;
	if (ha_flt_tab == 0) ha_opt_update_cov(&asm_opt, hom_cov);
	double t_idx = yak_realtime() - t0;
	if (nq == 0 || nq > R_INF.total_reads) nq = R_INF.total_reads;
	for (i = 0; i < nq; i++) bases += R_INF.read_length[i];
	double dt = 0; uint64_t n_src = 0, n_rev = 0;
	for (s = 0; s < warm + steps; s++) { // every step sees the same input: empty previous overlap lists
		for (i = 0; i < R_INF.total_reads; i++) R_INF.paf[i].length = R_INF.reverse_paf[i].length = 0;
		double t1 = yak_realtime();
		cal_ov_r(asm_opt.thread_num, nq, 1);
		if (s >= warm) dt += yak_realtime() - t1;
	}
	dt /= steps > 0 ? steps : 1;
	for (i = 0; i < nq; i++) { n_src += R_INF.paf[i].length; n_rev += R_INF.reverse_paf[i].length; }
	printf("{\"seconds\": %.6f, \"query_reads\": %lu, \"query_bases\": %lu, \"total_reads\": %lu, \"threads\": %d, \"index_seconds\": %.3f, \"n_src\": %lu, \"n_rev\": %lu, \"hom_cov\": %d}\n",
	       dt, (unsigned long)nq, (unsigned long)bases, (unsigned long)R_INF.total_reads, asm_opt.thread_num, t_idx, (unsigned long)n_src, (unsigned long)n_rev, hom_cov);
	return 0;
}

// refdump stage <hifiasm args...>: the reference's WHOLE overlap / error-correction stage, timed the way SURVEY.md §8(d) defines the metric —
// from the first ha_ft_gen to the end of the final overlap pass (Assembly.cpp:2081-2108: filter table, number_of_round x ha_ec with the
// index rebuilt each round, ha_ec_ff = final index + cal_ov_r).  Prints one JSON line.
static int run_stage(int argc, char *argv[])
{
	int r, hom_cov = -1, het_cov = -1; uint64_t tot_b = 0, tot_e = 0, i, bases = 0, n_src = 0, n_rev = 0, corrected = 0;
	yak_reset_realtime();
	init_opt(&asm_opt);
	argv[1] = argv[0];
	if (!CommandLine_process(argc - 1, argv + 1, &asm_opt)) return 1;
	ha_flt_tab = NULL; ha_idx = NULL;
	const double t0 = yak_realtime();
	if (!(asm_opt.flag & HA_F_NO_KMER_FLT)) { ha_flt_tab = ha_ft_gen(&asm_opt, &R_INF, &hom_cov, 0, 0); ha_opt_update_cov(&asm_opt, hom_cov); }
	const double t_ft = yak_realtime() - t0;
	for (r = 0; r < asm_opt.number_of_round; ++r) {
		ha_opt_reset_to_round(&asm_opt, r);
		tot_b = tot_e = 0;
		ha_ec(r, asm_opt.number_of_pround, (r < asm_opt.number_of_round - 1) ? 1 : 0, &tot_b, &tot_e);
		corrected += tot_e; if (r == 0) bases = tot_b;
	}
	const double t_ec = yak_realtime() - t0 - t_ft;
	ha_opt_reset_to_round(&asm_opt, asm_opt.number_of_round);
	ha_ec_ff(1);
	const double t_all = yak_realtime() - t0;
	for (i = 0; i < R_INF.total_reads; i++) { n_src += R_INF.paf[i].length; n_rev += R_INF.reverse_paf[i].length; }
	printf("{\"seconds\": %.6f, \"filter_seconds\": %.3f, \"ec_rounds_seconds\": %.3f, \"final_seconds\": %.3f, \"reads\": %lu, \"bases\": %lu, \"corrected_bases\": %lu, \"threads\": %d, \"n_src\": %lu, \"n_rev\": %lu}\n",
	       t_all, t_ft, t_ec, t_all - t_ft - t_ec, (unsigned long)R_INF.total_reads, (unsigned long)bases, (unsigned long)corrected, asm_opt.thread_num, (unsigned long)n_src, (unsigned long)n_rev);
	return 0;
}

// refdump ftq <hashes.bin> <counts.bin> <hifiasm args...>: the reference's filter table (ha_ft_gen with the given -f, i.e. with its Bloom
// filter when -f >= 21) probed with ha_ft_cnt (htab.cpp:1064) for every 64-bit hash of hashes.bin -> int32 counts; prints hom_cov.
static int run_ftq(int argc, char *argv[])
{
	int hom_cov = -1; const char *fin = argv[2], *fout = argv[3];
	yak_reset_realtime();
	init_opt(&asm_opt);
	argv[3] = argv[0];
	if (!CommandLine_process(argc - 3, argv + 3, &asm_opt)) return 1;
	ha_flt_tab = ha_ft_gen(&asm_opt, &R_INF, &hom_cov, 0, 0);
	FILE *fi = fopen(fin, "rb"), *fo = fopen(fout, "wb"); if (!fi || !fo) return 1;
	uint64_t h; while (fread(&h, 8, 1, fi) == 1) { int32_t c = ha_ft_cnt(ha_flt_tab, h); fwrite(&c, 4, 1, fo); }
	fclose(fi); fclose(fo);
	printf("{\"hom_cov\": %d, \"bf_shift\": %d}\n", hom_cov, asm_opt.bf_shift);
	return 0;
}

int main(int argc, char *argv[])
{
	if (argc == 4 && strcmp(argv[1], "edsemi") == 0) return run_edsemi(argv[2], argv[3]);
	if (argc >= 5 && strcmp(argv[1], "ftq") == 0) return run_ftq(argc, argv);
	if (argc >= 3 && strcmp(argv[1], "stage") == 0) return run_stage(argc, argv);
	if (argc >= 6 && strcmp(argv[1], "bench") == 0) return run_bench(argc, argv);
	if (argc < 4) { fprintf(stderr, "usage: refdump <raw|final> <out_prefix> <hifiasm args...>\n"); return 1; }
	const char *mode = argv[1], *pfx = argv[2];
	int r, hom_cov = -1, het_cov = -1; uint64_t tot_b, tot_e;
	yak_reset_realtime();
	init_opt(&asm_opt);
	argv[2] = argv[0];
	if (!CommandLine_process(argc - 2, argv + 2, &asm_opt)) return 1;

	ha_flt_tab = NULL; ha_idx = NULL;
	if (!(asm_opt.flag & HA_F_NO_KMER_FLT)) { // Assembly.cpp:2081-2085
		ha_flt_tab = ha_ft_gen(&asm_opt, &R_INF, &hom_cov, 0, 0);
		ha_opt_update_cov(&asm_opt, hom_cov);
	}
	if (strcmp(mode, "raw") == 0) {
		ha_idx = ha_pt_gen(&asm_opt, ha_flt_tab, 0, 0, &R_INF, &hom_cov, &het_cov); // Assembly.cpp:1007
		asm_opt.hom_cov = hom_cov; asm_opt.het_cov = het_cov;
		if (ha_flt_tab == 0) ha_opt_update_cov(&asm_opt, hom_cov);
		write_bins(pfx, "raw");
		dump_stages(pfx, asm_opt.is_ont ? 0.05 : 0.02); // bw of worker_hap_ec (ecovlp.cpp:3274)
		return 0;
	}
	int hook_round = -1;
	if (strncmp(mode, "round", 5) == 0) hook_round = atoi(mode + 5); // round<K>: dump the states of EC round K, then stop
	for (r = 0; r < asm_opt.number_of_round; ++r) { // Assembly.cpp:2088-2097
		ha_opt_reset_to_round(&asm_opt, r);
		tot_b = tot_e = 0;
		if (r == hook_round) { g_hook_pfx = pfx; g_full_calls = g_full_bases = 0; if (r > 0) write_tag("pre", 1, 1); } // round 0 parses the input inside ha_pt_gen: its "pre" state is the raw read set, no overlaps
		ha_ec(r, asm_opt.number_of_pround, (r < asm_opt.number_of_round - 1) ? 1 : 0, &tot_b, &tot_e);
		if (r == hook_round) {
			write_tag("post", 1, 1);
			FILE *fpa = xopen(pfx, ".params.txt");
			fprintf(fpa, "n_reads %lu\nround %d\nn_round %d\nhom_cov %d\nhet_cov %d\ntot_b %lu\ntot_e %lu\nfull_calls %lu\nfull_bases %lu\n", (unsigned long)R_INF.total_reads, r, asm_opt.number_of_round,
			        asm_opt.hom_cov, asm_opt.het_cov, (unsigned long)tot_b, (unsigned long)tot_e, (unsigned long)g_full_calls, (unsigned long)g_full_bases);
			fclose(fpa);
			return 0;
		}
		fprintf(stderr, "[refdump] round %d: bases %lu corrected %lu\n", r + 1, (unsigned long)tot_b, (unsigned long)tot_e);
	}
	ha_opt_reset_to_round(&asm_opt, asm_opt.number_of_round);
	write_bins(pfx, "pre");
	// ha_ec_ff(1), Assembly.cpp:1942-1959, opened up so the index can be probed
	if (ha_idx) { ha_pt_destroy(ha_idx); ha_idx = NULL; }
	ha_idx = ha_pt_gen(&asm_opt, ha_flt_tab, 1, 0, &R_INF, &hom_cov, &het_cov);
	asm_opt.hom_cov = hom_cov; asm_opt.het_cov = het_cov;
	dump_stages(pfx, 0.001); // bw of worker_hap_dc_ec_gen_new_idx (ecovlp.cpp:3957)
	double t0 = yak_realtime();
	cal_ov_r(asm_opt.thread_num, R_INF.total_reads, 1);
	fprintf(stderr, "[refdump] cal_ov_r wall %.3f s\n", yak_realtime() - t0);
	ha_pt_destroy(ha_idx); ha_idx = NULL;
	write_bins(pfx, "fin");
	return 0;
}
