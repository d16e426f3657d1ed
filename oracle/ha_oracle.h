/* oracle/ha_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of hifiasm's overlap hot path
 * (reference v0.25.0-r726), written from the reference's behaviour; every
 * function cites the reference file:line it follows.  It is the checker for the
 * CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline
 * leg may load it.  The product (hifiasm_b200/) never includes or links it.
 *
 * Pinned against the reference itself: tests/test_oracle_vs_ref.py compares
 * every function here with dumps produced by oracle/_ref/refdump (the
 * unmodified reference objects) — see tests/golden/.
 */
#ifndef HA_ORACLE_H
#define HA_ORACLE_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- read store: All_reads (Process_Read.h:115-146), flattened ---------- */
typedef struct {
	uint64_t n;               /* total_reads */
	const uint64_t *len;      /* read_length[n] */
	const uint64_t *off;      /* byte offset of read i in packed[], n+1 */
	const uint8_t *packed;    /* 2-bit, b0<<6|b1<<4|b2<<2|b3, len/4+1 bytes */
	const uint64_t *n_off;    /* N_site offsets n+1 (may be NULL) */
	const uint64_t *n_pos;    /* N positions */
} hao_reads_t;

/* ha_mz1_t (htab.h:13-18): x + {rid:28,pos:27,rev:1,span:8} */
typedef struct { uint64_t x, info; } hao_mz_t;
#define HAO_MZ_RID(m)  ((uint32_t)((m).info & 0xfffffffULL))
#define HAO_MZ_POS(m)  ((uint32_t)(((m).info >> 28) & 0x7ffffffULL))
#define HAO_MZ_REV(m)  ((uint32_t)(((m).info >> 55) & 1ULL))
#define HAO_MZ_SPAN(m) ((uint32_t)((m).info >> 56))

/* k_mer_hit (Hash_Table.h:116-120) */
typedef struct { uint32_t id_strand; /* readID:31 | strand<<31 */ uint32_t offset, self_offset, cnt; } hao_hit_t;

/* ma_hit_t (Overlaps.h:116-124), unpacked */
typedef struct {
	uint64_t qns; uint32_t qe, tn, ts, te;
	uint32_t ml, rev, bl, del; uint8_t el, no_l_indel;
} hao_ma_t;

/* the overlap_region fields the path defines (Hash_Table.h:78-106) */
typedef struct {
	uint32_t x_pos_s, x_pos_e, y_id, y_pos_s, y_pos_e, y_pos_strand;
	int32_t shared_seed;
	uint32_t align_length, non_homopolymer_errors, overlapLen;
	uint8_t is_match, without_large_indel; int8_t strong; uint8_t pad;
	uint32_t fc_off, fc_n; /* Fake_Cigar entries in the per-read fc pool */
} hao_ovlp_t;

typedef struct hao_ft_s hao_ft_t; /* yak_ft_t: hash -> count */
typedef struct hao_pt_s hao_pt_t; /* ha_pt_t : hash -> position list */

typedef struct {
	int k, w, is_hpc;          /* asm_opt.k_mer_length, mz_win, !(flag&HA_F_NO_HPC) */
	int mz_sample_dist, mz_rewin; /* CommandLines.cpp:266-268 */
	int min_hist_kmer_cnt;     /* 5, CommandLines.cpp:277 */
	double high_factor;        /* 5.0 */
	int max_kmer_cnt;          /* 2000 */
	int max_n_chain;           /* asm_opt.max_n_chain */
	int hom_cov, het_cov;
} hao_opt_t;

void hao_opt_default(hao_opt_t *o);
/* ha_opt_update_cov (CommandLines.cpp:411) */
void hao_opt_update_cov(hao_opt_t *o, int hom_cov);

/* recover_UC_Read / recover_UC_Read_sub_region (Process_Read.cpp:716,524) */
void hao_decode(const hao_reads_t *r, uint64_t id, char *out);
void hao_decode_sub(const hao_reads_t *r, uint64_t id, int64_t start, int64_t len, int strand, char *out);

/* mz1_ha_sketch (sketch.cpp:454-579) incl. mz1_select_mz_h; *out malloc'd */
int hao_sketch(const char *s, int len, int w, int k, uint32_t rid, int is_hpc, const hao_ft_t *ft,
               int sample_dist, int rewin, hao_mz_t **out, uint32_t *n_out);

/* ha_ft_gen (htab.cpp:1136) with exact counting (-f0) */
hao_ft_t *hao_ft_gen(const hao_reads_t *r, const hao_opt_t *o, int *hom_cov);
hao_ft_t *hao_ft_gen_bf(const hao_reads_t *r, const hao_opt_t *o, int bf_shift, int *hom_cov); /* -f bf_shift: the Bloom-filtered counting of ha_ft_gen */
uint64_t hao_all_kmers(const hao_reads_t *r, const hao_opt_t *o, uint64_t *out, uint64_t cap);
int32_t hao_ft_cnt(const hao_ft_t *ft, uint64_t y); /* ha_ft_cnt htab.cpp:1064 */
uint64_t hao_ft_size(const hao_ft_t *ft);
void hao_ft_destroy(hao_ft_t *ft);

/* ha_pt_gen (htab.cpp:1232), normal mode */
hao_pt_t *hao_pt_gen(const hao_reads_t *r, const hao_ft_t *ft, const hao_opt_t *o, int *hom_cov, int *het_cov);
const uint64_t *hao_pt_get(const hao_pt_t *pt, uint64_t hash, int *n); /* ha_pt_get htab.cpp:518 */
uint64_t hao_pt_tot_pos(const hao_pt_t *pt);
uint64_t hao_pt_n_keys(const hao_pt_t *pt);
void hao_pt_destroy(hao_pt_t *pt);
/* ha_analyze_count (hist.cpp:74) */
int hao_analyze_count(int n_cnt, int start_cnt, int m_peak_hom, const int64_t *cnt, int *peak_het);

/* minimizers_qgen0 (anchor.cpp:987-1081): cl->list; *out malloc'd */
int hao_anchors(const hao_reads_t *r, const hao_pt_t *pt, const hao_mz_t *mz, uint32_t n_mz,
                uint32_t high_occ, uint32_t low_occ, hao_hit_t **out, uint64_t *n_out);

/* h_ec_lchain minus the seeding (anchor.cpp:2302 -> lchain_qgen_mcopy_fast 1920):
 * hits[] is cl->list (modified in place: chain hits compacted to the front,
 * *n_hits updated); chains malloc'd; fc = Fake_Cigar pool malloc'd (u64) */
int hao_lchain(const hao_reads_t *r, uint32_t rid, hao_hit_t *hits, uint64_t *n_hits, double bw_rate, int mz_k,
               int max_n_chain, hao_ovlp_t **out, uint32_t *n_out, uint64_t **fc, uint64_t *n_fc);

/* worker_hap_dc_ec_gen_new_idx (ecovlp.cpp:3948-3992) for one read.
 * in0/in1 = previous paf / reverse_paf of the read (in0 is modified: el reset,
 * ecovlp.cpp:5103); out0/out1 must hold n_chains_max + n0 entries. */
int hao_final_read(const hao_reads_t *r, const hao_pt_t *pt, const hao_ft_t *ft, const hao_opt_t *o, uint32_t rid,
                   hao_ma_t *in0, uint32_t n0, const hao_ma_t *in1, uint32_t n1,
                   hao_ma_t **out0, uint32_t *m0, hao_ma_t **out1, uint32_t *m1);

/* ed_band_cal_semi_64_w_absent_diag (Levenshtein_distance.h:3727-3776):
 * returns err (INT32_MAX if > thre semantics of the reference) and *pe */
int hao_ed_semi_64_absent_diag(const char *pstr, int32_t pn, const char *tstr, int32_t tn, int32_t thre, int32_t abs_diag, int32_t *pe);

/* per-window work of align_hc_ed_post_extz (Correct.cpp:12951-13011) for every window of every chain */
typedef struct { int32_t chain, q_s, q_e, t_s, t_pri_l, thre, aux_beg, aux_end, err, pe; } hao_win_rec_t;

void hao_free(void *p);
#ifdef __cplusplus
}
#endif
#endif
