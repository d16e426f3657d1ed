/* hifiasm_b200.h — C-ABI of the B200-native overlap engine.
 *
 * Drop-in boundary for hifiasm's overlap / error-correction stage (reference
 * v0.25.0-r726).  hifiasm has no plugin API: the seam is a set of C++ free
 * functions over process globals.  Every entry point below
 * names the reference interface it replaces; INTEGRATION.md shows the shim a
 * hifiasm maintainer links in place of ecovlp.o / htab.o for this path.
 *
 * Conventions: plain pointers and sizes only; all buffers are caller-owned
 * host memory unless stated; every call returns 0 on success or a negative
 * HB_E_* code (hb_last_error() gives the text).  There is NO CPU fallback:
 * without a CUDA device hb_create() fails with HB_E_NO_DEVICE.
 */
#ifndef HIFIASM_B200_H
#define HIFIASM_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK            0
#define HB_E_NO_DEVICE  (-1)
#define HB_E_CUDA       (-2)
#define HB_E_ARG        (-3)
#define HB_E_STATE      (-4)
#define HB_E_OVERFLOW   (-5)
#define HB_E_NOMEM      (-6)
#define HB_E_IO         (-7)

typedef struct hb_ctx hb_ctx_t;

/* The fields of hifiasm_opt_t (CommandLines.h:38-186) the path reads. */
typedef struct {
	int32_t k_mer_length;      /* -k, CommandLines.cpp:259 (51) */
	int32_t mz_win;            /* -w, CommandLines.cpp:263 (51) */
	int32_t is_hpc;            /* !(flag & HA_F_NO_HPC) */
	int32_t mz_sample_dist;    /* CommandLines.cpp:268 (500) */
	int32_t mz_rewin;          /* CommandLines.cpp:266 (1000) */
	int32_t min_hist_kmer_cnt; /* CommandLines.cpp:277 (5) */
	int32_t max_kmer_cnt;      /* CommandLines.cpp:270 (2000) */
	int32_t max_n_chain;       /* -N, CommandLines.cpp:276 (100) */
	double  high_factor;       /* -D, CommandLines.cpp:271 (5.0) */
	int32_t hom_cov, het_cov;  /* asm_opt.hom_cov / het_cov */
	int32_t is_ont;            /* --ont (not supported yet: must be 0) */
	int32_t bf_shift;          /* -f, CommandLines.cpp:269: bits of the Bloom filter in front of ha_ft_gen's counting; 0 (hb_opt_init) = exact
	                              counting = hifiasm -f0; hifiasm's own default is 37.  Active from 21 (htab.cpp:140-158, 78-91) */
} hb_opt_t;

/* ha_mz1_t (htab.h:13-18): x = hash, info = rid:28 | pos:27 | rev:1 | span:8 */
typedef struct { uint64_t x, info; } hb_mz_t;
/* k_mer_hit (Hash_Table.h:116-120): id_strand = readID:31 | strand:1 */
typedef struct { uint32_t id_strand, offset, self_offset, cnt; } hb_hit_t;
/* ma_hit_t (Overlaps.h:116-124) with the bit-fields widened */
typedef struct {
	uint64_t qns; uint32_t qe, tn, ts, te;
	uint32_t ml, rev, bl, del;
	uint8_t el, no_l_indel, pad[6];
} hb_ma_hit_t;
/* the overlap_region fields the path defines (Hash_Table.h:78-106) */
typedef struct {
	uint32_t x_pos_s, x_pos_e, y_id, y_pos_s, y_pos_e, y_pos_strand;
	int32_t shared_seed;
	uint32_t first_hit;  /* overlap_region.non_homopolymer_errors after h_ec_lchain */
	uint32_t n_hits;     /* number of anchors on the chain */
	uint32_t fc_off, fc_n; /* Fake_Cigar entries (u64: x<<32 | |dd|<<1 | sign) in the per-read pool */
	uint32_t pad;
} hb_chain_t;

void hb_opt_init(hb_opt_t *o);                       /* init_opt, CommandLines.cpp:243 */
void hb_opt_update_cov(hb_opt_t *o, int hom_cov);    /* ha_opt_update_cov, CommandLines.cpp:411 */

int hb_device_count(void);
/* one context per GPU (one process per GPU in multi-GPU runs) */
int hb_create(hb_ctx_t **ctx, int device, const hb_opt_t *opt);
void hb_destroy(hb_ctx_t *ctx);
const char *hb_last_error(const hb_ctx_t *ctx);
int hb_set_opt(hb_ctx_t *ctx, const hb_opt_t *opt);
int hb_get_opt(const hb_ctx_t *ctx, hb_opt_t *opt);

/* ---- read store: All_reads R_INF (Process_Read.h:115-148) -------------------
 * Mirrors read_length[], read_sperate[] (2-bit, ha_compress_base layout,
 * Process_Read.cpp:792) and N_site[] into HBM.  packed = concatenated per-read
 * byte arrays, byte_off[n+1]; n_pos/n_off = flattened N_site lists.          */
int hb_reads_upload(hb_ctx_t *ctx, uint64_t n_reads, const uint64_t *read_length,
                    const uint8_t *packed, const uint64_t *byte_off,
                    const uint64_t *n_pos, const uint64_t *n_off);
/* the same from the reference's own arrays of per-read pointers */
int hb_reads_upload_ptrs(hb_ctx_t *ctx, uint64_t n_reads, const uint64_t *read_length,
                         uint8_t *const *read_sperate, uint64_t *const *N_site);

/* ---- index: ha_ft_gen / ha_pt_gen (htab.h:77,83; htab.cpp:1136,1232) -------
 * hb_ft_gen counts all (HPC) k-mers of the resident reads — exactly when
 * opt.bf_shift = 0 (hifiasm -f0, the default of hb_opt_init), behind a Bloom
 * filter of 2^bf_shift bits like hifiasm -f<bf_shift> otherwise (its own default
 * is 37) — and keeps those occurring >= high_factor*peak_hom times.
 * hb_pt_gen sketches every read, counts minimizers, derives hom/het peaks
 * (ha_analyze_count, hist.cpp:74) and builds the position index in HBM.      */
int hb_ft_gen(hb_ctx_t *ctx, int *hom_cov);
int hb_ft_size(const hb_ctx_t *ctx, uint64_t *n);
int hb_ft_cnt(hb_ctx_t *ctx, const uint64_t *hash, uint64_t n, int32_t *cnt); /* ha_ft_cnt, htab.cpp:1064 */
void hb_ft_destroy(hb_ctx_t *ctx);                                           /* ha_ft_destroy */
int hb_pt_gen(hb_ctx_t *ctx, int *hom_cov, int *het_cov);
int hb_pt_stat(const hb_ctx_t *ctx, uint64_t *n_keys, uint64_t *n_pos);
/* ha_pt_get (htab.cpp:518) batched: cnt[i] = occurrences of hash[i]; when pos
 * != NULL it receives the concatenated ha_idxpos_t lists (pos_cap entries max) */
int hb_pt_get(hb_ctx_t *ctx, const uint64_t *hash, uint64_t n, uint32_t *cnt, uint64_t *pos, uint64_t pos_cap);
void hb_pt_destroy(hb_ctx_t *ctx);                                           /* ha_pt_destroy */

/* ---- stages of h_ec_lchain (anchor.cpp:2302), for parity tests --------------
 * Each works on the read-id range [r0, r1) of the resident store and returns
 * flattened per-read arrays: off[r1-r0+1] + records.  Call with rec == NULL to
 * get only off[] (sizes).                                                    */
int hb_sketch(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *off, hb_mz_t *rec, uint64_t rec_cap);       /* mz1_ha_sketch, sketch.cpp:454 */
int hb_anchors(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *off, hb_hit_t *rec, uint64_t rec_cap);     /* minimizers_qgen0, anchor.cpp:987 */
/* chains after lchain_qgen_mcopy_fast (anchor.cpp:1920); hit_off/hits = the
 * compacted chain anchors (cl->list[0..cl->length)); fc = fake-cigar pool     */
int hb_chains(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres,
              uint64_t *off, hb_chain_t *rec, uint64_t rec_cap,
              uint64_t *hit_off, hb_hit_t *hits, uint64_t hit_cap,
              uint64_t *fc_off, uint64_t *fc, uint64_t fc_cap);

/* ---- window pass of an EC round: for every chain h_ec_lchain returns for
 * reads [r0,r1) and every WINDOW_HC-sized query window of it, the work align_hc_ed_post_extz does
 * per window (Correct.cpp:12951-13011): threshold, target start from the fake cigar, init_waln
 * clipping, ed_band_cal_semi_64_w_absent_diag (Levenshtein_distance.h:3727) on the packed reads.
 * chain = index into the read's chain list (hb_chains order); err = INT32_MAX when the window does
 * not align within thre, t_pri_l = -1 when init_waln rejects it.  off[r1-r0+1] + records.        */
typedef struct { int32_t chain, q_s, q_e, t_s, t_pri_l, thre, aux_beg, aux_end, err, pe; } hb_win_t;
int hb_windows(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l,
               uint64_t *off, hb_win_t *rec, uint64_t rec_cap);

/* ---- alignment stage of an EC round, step A (SURVEY.md §8 rows a8 + a9): for every chain h_ec_lchain returns for
 * reads [r0,r1), what gen_hc_r_alin (Correct.cpp:25617-25645) does before the base-level CIGAR:
 * align_hc_ed_post_extz (12951: window pass, gap filling by push_hc_wlst_exz 12776, 0.9 aligned-fraction cut) and
 * gen_extend_err_exz (13400: error estimate of the still unaligned windows by extension from their neighbours).
 * hb_wl_t is window_list (Hash_Table.h:54-62); cidx indexes cig[] (uint16 = op<<14 | len, push_trace
 * Levenshtein_distance.h:522).  st: 0 = rejected by the window pass, 1 = aligned but rr > e_rate, 2 = accepted with
 * re = estimated number of errors.  off[r1-r0+1] counts overlaps per read (hb_chains order); rec[j].w_off/w_n
 * locate the overlap's window list in wl[].  *n_wl / *n_cig receive the used sizes of wl[] / cig[].                */
typedef struct { int32_t x_start, x_end, y_start, y_end; int16_t extra_begin, extra_end, error, error_threshold; uint32_t cidx, clen; } hb_wl_t;
typedef struct { int32_t st; uint32_t align_length; double rr; int64_t re; uint64_t w_off; uint32_t w_n, pad; } hb_aln_t;
int hb_ec_align(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l,
                uint64_t *off, hb_aln_t *rec, uint64_t rec_cap, hb_wl_t *wl, uint64_t wl_cap, uint16_t *cig, uint64_t cig_cap,
                uint64_t *n_wl, uint64_t *n_cig);

/* ---- alignment stage of an EC round, step B (row a10): base-level CIGAR of every overlap step A accepted —
 * gen_hc_fast_cigar (Correct.cpp:25137 -> 17813): return_t_chain, hc_ovlp_base_direct, hc_aln_exz_adv_hc with its
 * threshold escalation, multi-word banded Myers with traceback, push_alnw, update_overlap_region.
 * rec[j] (hb_chains order): st = step A's status (2 = accepted and aligned here); re = error total (Correct.cpp:17857);
 * x/y_pos_* = the overlap's coordinates after update_overlap_region; w_off/w_n locate its window list in wl[];
 * an overlap that keeps an unaligned window of >= 512 bp on both reads goes through the re-seeding rescue (rechain_aln_hc,
 * Correct.cpp:17669: exact runs of step A's window alignments -> hits -> chain with fixed ends -> the pieces between the hits
 * aligned again; k_ecb_rechain); need_rechain = 1 is left only when that rescue ran out of scratch: the result is then not final.
 * gaps != 0 adds step C (row a11), reassign_gaps (Correct.cpp:25409): the indels of every window are left-normalised
 * (move_wins 25274, adjust_gap 25167, ajust_end_cigar 25252); nh_err = overlap_region.non_homopolymer_errors afterwards
 * (step A's estimate minus the mismatches the normalisation removed); re stays step B's total.
 * gaps & 2 adds row a12, the exact shortcut of gen_hc_r_alin_ea (ecovlp.cpp:2810): a chain that matches an exact (el) record of the read's
 * overlap list of the previous round (staged with hb_ec_stage_prev = R_INF.paf[] flattened) in target, strand and coordinates, and whose
 * substrings are still identical, is accepted without alignment: one exact window, nh_err = 0, pad = 1.                                  */
int hb_ec_stage_prev(hb_ctx_t *ctx, const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off);
typedef struct { int32_t st, need_rechain; int64_t re, nh_err; uint32_t x_pos_s, x_pos_e, y_pos_s, y_pos_e; uint64_t w_off; uint32_t w_n, pad; } hb_alnb_t;
int hb_ec_cigar(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t gaps,
                uint64_t *off, hb_alnb_t *rec, uint64_t rec_cap, hb_wl_t *wl, uint64_t wl_cap, uint16_t *cig, uint64_t cig_cap,
                uint64_t *n_wl, uint64_t *n_cig);

/* ---- phasing of an EC round (row a13) on top of steps A-C: rphase_hc (Correct.cpp:20191, HiFi path) for every read of [r0,r1).
 * rec[j] (hb_chains order): st = step A's status; for st == 2 the overlap's coordinates after step C, non_homopolymer_errors,
 * is_match (1 = same haplotype, 2 = carries informative minor alleles: the other haplotype) and strong
 * (overlap_region.is_match / strong, Hash_Table.h:78-106); is_match = 0 for overlaps the alignment stage rejected.          */
typedef struct { int32_t st; uint32_t y_id, rev, x_pos_s, x_pos_e, y_pos_s, y_pos_e, nh_err, is_match; int32_t strong; uint32_t need_rechain, pad; } hb_phase_t;
int hb_ec_phase(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, uint64_t *off, hb_phase_t *rec, uint64_t rec_cap);

/* ---- R_INF.reverse_paf[i] of an EC round (part of row a15): the overlaps phasing assigned to the other haplotype, after dedup_chains
 * (ecovlp.cpp:2984: best chain per target), as push_ne_ovlp(flag = 2) emits them (ecovlp.cpp:2585; el / del are not defined on that
 * path and come back 0).  off[r1-r0+1] + records.  hb_ec_round_lists below returns both lists of the round.                              */
int hb_ec_reverse_paf(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, uint64_t *off, hb_ma_hit_t *rec, uint64_t rec_cap);

/* ---- the rest of an EC round: cal_ec_r (ecovlp.h:13; ecovlp.cpp:6268) after the alignment and phasing of every read ------------------
 * Edit scripts: scc.a[i] (ecovlp.cpp:101), one per read, the reference's own encoding (push_trace_bp_f, Levenshtein_distance.h:640):
 * uint16 runs, op = w >> 14: 0 match (len 14 bits) / 1 mismatch (new base 2 bits, old base 2 bits, len 10 bits) / 2 insertion (base, len 12
 * bits) / 3 deletion (base, len 12 bits).  hb_ec_round writes them on the device (row a14); hb_ec_stage_scc takes them from the caller instead
 * (scc = concatenated scripts, scc_off[n_reads + 1]: scripts gathered from other GPUs' shards, or the reference's own).  Rows a15-a17 read them from HBM. */
int hb_ec_stage_scc(hb_ctx_t *ctx, const uint16_t *scc, const uint64_t *scc_off);
/* row a15 — what cal_ec_multiple / worker_hap_ec (ecovlp.cpp:6063, 3234) leaves in R_INF.paf[i] / R_INF.reverse_paf[i] for reads [r0,r1):
 * alignment stage (use_prev != 0: with the exact shortcut of gen_hc_r_alin_ea against the lists staged by hb_ec_stage_prev), rphase_hc,
 * dedup_chains, then push_ne_ovlp(flag 1, scc.a[i]) (2585: longest exact interval of each same-haplotype overlap mapped through the read's
 * edit script by extract_max_exact 2520, no_l_indel from the large-indel test of wcns_gen 2299-2360), push_ne_ovlp(flag 2) and
 * check_well_cal (2750): flags[2 i] = is_fully_corrected, flags[2 i + 1] = is_abnormal.  src_off / rev_off: r1 - r0 + 1 entries.          */
int hb_ec_round_lists(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t use_prev,
                      uint64_t *src_off, hb_ma_hit_t *src, uint64_t src_cap, uint64_t *rev_off, hb_ma_hit_t *rev, uint64_t rev_cap, uint8_t *flags);
/* rows a14 + a15 — the same with the consensus on the device: wcns_gen (ecovlp.cpp:2293) writes each read's edit script (wcns_vote 2185 per 512
 * columns, anchors by push_cns_anchor 2109, stretches between anchors voted by cns_gen0 1159), which then feeds push_ne_ovlp / check_well_cal as above.
 * scc_off[r1 - r0 + 1] + scc = the scripts of the range (scc may be NULL: sizes only); *n_corrected = corrected bases (cal_ec_multiple's second counter).
 * status[i]: 0 = done; bit 0 = the read's graph consensus (cns_gen_full, ecovlp.cpp:1919) ran out of its arena — it gets an empty script and
 * its lists carry no exact intervals; bit 2 = the re-seeding rescue (rechain_aln_hc) of one of the read's overlaps ran out of scratch.  When the range is
 * the whole store the scripts stay staged in HBM (as by hb_ec_stage_scc) for hb_ec_apply / hb_ec_update_paf.                                       */
int hb_ec_round(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, double bw_thres, double e_rate, int32_t w_l, int32_t use_prev,
                uint64_t *src_off, hb_ma_hit_t *src, uint64_t src_cap, uint64_t *rev_off, hb_ma_hit_t *rev, uint64_t rev_cap, uint8_t *flags,
                uint64_t *scc_off, uint16_t *scc, uint64_t scc_cap, uint8_t *status, uint64_t *n_corrected);
/* row a16 — sl_ec_r / worker_sl_ec (ecovlp.cpp:6402, 5965): every resident read with its staged edit script applied; the read store in
 * HBM is replaced (new lengths, 2-bit bases, N lists).  The filter table and the position index still describe the old reads: rebuild
 * them (hb_ft_gen / hb_pt_gen) before the next pass, as ha_ec does (Assembly.cpp:1007,1026).                                            */
int hb_ec_apply(hb_ctx_t *ctx, uint64_t *n_changed, uint64_t *total_bases);
/* row a17 — cal_update_ec_multiple / worker_update_dc_ec (ecovlp.cpp:6095, 3808) on the corrected store: every record of paf[] (flattened,
 * paf_off[n_reads + 1], updated in place) whose el is set has its interval remapped through the TARGET's edit script
 * (adjust_exact_match 3454), extended to the read ends and compared base by base (quick_exact_match 3521); el is rewritten.             */
int hb_ec_update_paf(hb_ctx_t *ctx, hb_ma_hit_t *paf, const uint64_t *paf_off, uint64_t *n_exact, uint64_t *n_inexact);
/* row a18 — worker_hap_post_rev (ecovlp.cpp:3866): resident reads reverse-complemented in HBM; both lists flipped by flip_paf_rc (3845),
 * compacted in place (records and offsets are rewritten; lists only shrink).  Either list may be NULL.                                   */
int hb_ec_post_rev(hb_ctx_t *ctx, hb_ma_hit_t *paf, uint64_t *paf_off, hb_ma_hit_t *rpaf, uint64_t *rpaf_off);
/* cal_ec_r(n_thre, round, n_round, n_a, is_sv, &tot_b, &tot_e) (ecovlp.h:13; ecovlp.cpp:6268) as one call on the resident store and index: the five
 * steps above in the reference's order.  prev_src = R_INF.paf[] of the previous round (empty offsets in round 0); out_* = this round's paf[] /
 * reverse_paf[] as the round leaves them (updated, flipped when the round reverses: !is_sv || (round & 1)); flags / status as hb_ec_round; tot_b /
 * tot_e = the two counters of the "[M::pec] # bases / # corrected bases" line, n_exact / n_inexact those of "# exact o / # non-exact o".
 * e_rate = asm_opt.max_ov_diff_ec (0.04), w_l = WINDOW_HC (775).  n_round (asm_opt.number_of_pround) must be 0.  The corrected reads stay in HBM
 * (hb_reads_download copies them back); the index describes the old reads and must be rebuilt before the next pass.                           */
int hb_cal_ec_r(hb_ctx_t *ctx, uint64_t round, uint64_t n_round, uint64_t is_sv, double e_rate, int32_t w_l,
                const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off,
                hb_ma_hit_t *out_src, uint64_t *out_src_off, uint64_t out_src_cap, hb_ma_hit_t *out_rev, uint64_t *out_rev_off, uint64_t out_rev_cap,
                uint8_t *flags, uint8_t *status, uint64_t *tot_b, uint64_t *tot_e, uint64_t *n_exact, uint64_t *n_inexact);
/* the resident read store back in the All_reads layout (Process_Read.h:115-146): read_length[n], packed = len/4+1 bytes per read
 * concatenated (pad bits zero), n_off[n + 1] / n_pos = the N_site lists.  Any output may be NULL.                                       */
int hb_reads_download(hb_ctx_t *ctx, uint64_t *read_length, uint8_t *packed, uint64_t packed_cap, uint64_t *n_off, uint64_t *n_pos, uint64_t n_pos_cap);

/* ---- final overlap pass: cal_ov_r(n_thre, n_a, new_idx=1) (ecovlp.h:15;
 * ecovlp.cpp:6385 -> worker_hap_dc_ec_gen_new_idx 3948) -----------------------
 * prev_* = R_INF.paf[] / R_INF.reverse_paf[] of the last EC round, flattened
 * (off[n_reads+1]).  Results replace them: out_*_off[n_reads+1] and records
 * (caps in entries).  stat[7] receives the counters ha_print_ovlp_stat_0
 * prints (forward, reverse, strong, weak, exact, no_l_indel, inexact).
 * Reads [r0,r1) only are processed (r0=0,r1=n_reads for the whole store);
 * offsets are relative to r0.                                                */
int hb_cal_ov_r(hb_ctx_t *ctx, uint64_t r0, uint64_t r1,
                const hb_ma_hit_t *prev_src, const uint64_t *prev_src_off,
                const hb_ma_hit_t *prev_rev, const uint64_t *prev_rev_off,
                hb_ma_hit_t *out_src, uint64_t *out_src_off, uint64_t out_src_cap,
                hb_ma_hit_t *out_rev, uint64_t *out_rev_off, uint64_t out_rev_cap,
                uint64_t *stat);
/* device-resident variant used by bench.py: inputs staged once, result stays in
 * HBM; returns totals only.  Timed region of `value` (no PCIe).               */
int hb_cal_ov_r_resident(hb_ctx_t *ctx, uint64_t r0, uint64_t r1, uint64_t *n_src, uint64_t *n_rev, uint64_t *stat);


/* ---- the whole stage as one call: what ha_assemble() runs between reading the reads and building the string graph (Assembly.cpp:2076-2108) on the
 * resident store — ha_ft_gen, n_round x ha_ec (ha_pt_gen + cal_ec_r, Assembly.cpp:996-1030), ha_ec_ff (ha_pt_gen + cal_ov_r, 1942-1959).  The
 * corrected reads stay in HBM (hb_reads_download); the final R_INF.paf[] / R_INF.reverse_paf[] come back flattened in host memory owned by the
 * context (valid until the next hb_stage_run / hb_destroy), with the last EC round's is_fully_corrected / is_abnormal flags.
 * world > 1 (one process per GPU, every rank holds all reads): the query reads of every pass are sharded contiguously over the ranks, and each EC
 * round ends with one all-gather of the shard's edit scripts + lists through the caller's transport: allgather(user, send, recv, bytes) must deliver
 * `bytes` bytes from every rank into recv[rank * bytes ..] (host memory) on all ranks, and return 0.  Every rank leaves with the same lists.        */
typedef int (*hb_allgather_fn)(void *user, const void *send, void *recv, uint64_t bytes_per_rank);
typedef struct {
	uint64_t n_reads, n_src, n_rev;
	const hb_ma_hit_t *src, *rev; const uint64_t *src_off, *rev_off; /* n_reads + 1 offsets each */
	const uint8_t *is_fully_corrected, *is_abnormal;
	int32_t hom_cov, het_cov;                                        /* of the final index (written to <o>.ec.bin) */
	uint64_t corrected_bases[8], n_unfinished;                       /* per round: the second counter of "[M::pec] # corrected bases"; reads a round reported (status != 0) */
	double ms_ft, ms_pt[9], ms_ec[8], ms_final, ms_exchange, ms_total; /* host wall clock per step */
	double device_ms;                                                /* CUDA events on the context's stream around the whole call */
} hb_stage_result_t;
int hb_stage_run(hb_ctx_t *ctx, int n_round, int rank, int world, hb_allgather_fn allgather, void *user, hb_stage_result_t *res);

/* ---- window alignment: ed_band_cal_semi_64_w_absent_diag
 * (Levenshtein_distance.h:3727) batched.  Case i: pattern = pat[pat_off[i]..
 * pat_off[i+1]) (target slice, ASCII), text = txt[txt_off[i]..txt_off[i+1])
 * (query window), thre[i], abs_diag[i]; err[i] = INT32_MAX when unaligned.   */
int hb_ed_semi_64(hb_ctx_t *ctx, uint64_t n_cases, const char *pat, const uint64_t *pat_off,
                  const char *txt, const uint64_t *txt_off, const int32_t *thre, const int32_t *abs_diag,
                  int32_t *err, int32_t *pe);

/* ---- instrumentation ---------------------------------------------------- */
/* ---- the stage's on-disk products in the reference's own formats, written by the host side from the flat
 * arrays of this interface (no context, no GPU).  names / name_index = All_reads.name / name_index (name of read i =
 * names[name_index[i] .. name_index[i+1])).  Byte-identical to the reference's files on the same state.
 *   hb_write_paf       Output_PAF, Assembly.cpp:1673-1717 (--write-paf: the same-haplotype list R_INF.paf after cal_ov_r)
 *   hb_write_ec_fa     Output_corrected_reads, Assembly.cpp:884-905 (--write-ec: the reads after the EC rounds)
 *   hb_write_ovlp_bin  write_ma_hit_ts, Overlaps.cpp:23442-23465 (<o>.ovlp.source.bin / <o>.ovlp.reverse.bin; flags may be NULL = 0)
 *   hb_write_ec_bin    write_All_reads, Process_Read.cpp:69-125 (<o>.ec.bin; the pad byte of reads with len % 4 == 0 is whatever
 *                      packed[] holds — the reference leaves it uninitialised)                                     */
/* ingest (htab.cpp:761-813 step 0 + ha_insert_read_len Process_Read.cpp:414 + ha_compress_base 792): FASTA / FASTQ files, plain or gzip, ->
 * the flat All_reads arrays hb_reads_upload takes (read i: packed[byte_off[i] .. +len/4+1), N sites n_pos[n_off[i] .. n_off[i+1]), name
 * names[name_index[i] .. name_index[i+1])).  adapter_len = -z (bases trimmed from both ends; reads that become empty are skipped).
 * All arrays are malloc-owned: hb_readset_free releases them.                                                                        */
typedef struct {
	uint64_t n_reads, total_bases, total_name_length, index_size, name_index_size; /* the All_reads header fields write_All_reads stores */
	uint64_t *read_length, *byte_off; uint8_t *packed; uint64_t *n_off, *n_pos; char *names; uint64_t *name_index;
} hb_readset_t;
int hb_readset_load(const char *const *paths, int n_paths, int32_t adapter_len, hb_readset_t *out);
void hb_readset_free(hb_readset_t *rs);
int hb_write_paf(const char *path, uint64_t n_reads, const uint64_t *read_length, const char *names, const uint64_t *name_index,
                 const uint64_t *off, const hb_ma_hit_t *rec);
int hb_write_ec_fa(const char *path, uint64_t n_reads, const uint64_t *read_length, const uint8_t *packed, const uint64_t *byte_off,
                   const uint64_t *n_off, const uint64_t *n_pos, const char *names, const uint64_t *name_index);
int hb_write_ovlp_bin(const char *path, uint64_t n_reads, const uint64_t *off, const hb_ma_hit_t *rec, const uint8_t *is_fully_corrected, const uint8_t *is_abnormal);
int hb_write_ec_bin(const char *path, int32_t adapter_len, uint64_t index_size, uint64_t name_index_size, uint64_t n_reads, uint64_t total_reads_bases,
                    const uint64_t *read_length, const uint8_t *packed, const uint64_t *byte_off, const uint64_t *n_off, const uint64_t *n_pos,
                    const char *names, uint64_t total_name_length, const uint64_t *name_index, const uint8_t *trio_flag, int32_t hom_cov, int32_t het_cov);

/* per-kernel launch counters and device time of the last hb_cal_ov_r* call:
 * names[i] (static strings), launches[i], ms[i]; returns number of entries   */
int hb_profile(const hb_ctx_t *ctx, const char **names, uint64_t *launches, double *ms, int cap);
void hb_profile_reset(hb_ctx_t *ctx);
/* the same summed over every pass of the last hb_stage_run, with the counters of hb_counters summed likewise */
int hb_stage_profile(const hb_ctx_t *ctx, const char **names, uint64_t *launches, double *ms, int cap, uint64_t *counters, int ccap); /* the counters also restart at the beginning of every pass */
/* device time (ms) of the last pass, from CUDA events recorded on the context's
 * own stream around the whole pass (first launch .. last result resident)     */
int hb_last_pass_ms(const hb_ctx_t *ctx, double *ms);
/* algorithmic byte counters of the last pass:
 * c[0]=reads, c[1]=bases, c[2]=minimizers, c[3]=anchors, c[4]=groups, c[5]=chains */
int hb_counters(const hb_ctx_t *ctx, uint64_t *c, int cap);

#ifdef __cplusplus
}
#endif
#endif
