// hb_shim.cpp — the reference-side binding of libhifiasm_b200.so: hifiasm's own entry points of the overlap / error-correction stage, with the
// reference's signatures and side effects on its globals, implemented over the C-ABI of include/hifiasm_b200.h.
//
//   cal_ec_r (ecovlp.h:12; called from ha_ec, Assembly.cpp:1021)      one error-correction round
//   cal_ov_r (ecovlp.h:14; called from ha_ec_ff, Assembly.cpp:1956)   the final overlap pass
//
// Built by oracle/Makefile (target dropin) against the reference's headers and linked with the reference's OWN objects, in which these two symbols
// are weakened (objcopy) so that the definitions below win: oracle/_ref/hifiasm_b200_dropin is hifiasm's unmodified main() / ha_assemble() with the
// stage's compute on the GPU.  tests/test_gpu_dropin.py compares the files it writes with the stock binary's.  Everything the reference does around the
// two calls stays: option parsing, ingest (inside its first counting pass), its CPU index builds (ha_ec still calls ha_pt_gen: the reads arrive through
// it; the GPU builds its own index from R_INF), writers, and the downstream graph code, which consumes R_INF.paf / reverse_paf unchanged
// (malloc-owned buffers with size >= length: Overlaps.cpp:1050-1073).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "Process_Read.h"
#include "CommandLines.h"
#include "Overlaps.h"
#include "htab.h"
#include "Hash_Table.h"
#include "../include/hifiasm_b200.h"

static hb_ctx_t *g_hb;
static int g_have_ft;

static void hb_die(const char *what) { fprintf(stderr, "[hifiasm_b200] %s: %s\n", what, hb_last_error(g_hb)); exit(1); } // hifiasm's error convention

static void hb_sync_opt()
{
	hb_opt_t o; hb_opt_init(&o);
	o.k_mer_length = asm_opt.k_mer_length; o.mz_win = asm_opt.mz_win; o.is_hpc = !(asm_opt.flag & HA_F_NO_HPC);
	o.mz_sample_dist = asm_opt.mz_sample_dist; o.mz_rewin = asm_opt.mz_rewin; o.min_hist_kmer_cnt = asm_opt.min_hist_kmer_cnt;
	o.max_kmer_cnt = asm_opt.max_kmer_cnt; o.max_n_chain = asm_opt.max_n_chain; o.high_factor = asm_opt.high_factor;
	o.hom_cov = asm_opt.hom_cov; o.het_cov = asm_opt.het_cov; o.is_ont = asm_opt.is_ont;
	o.bf_shift = asm_opt.bf_shift;
	if (!g_hb) { if (hb_create(&g_hb, 0, &o)) { fprintf(stderr, "[hifiasm_b200] no CUDA device (there is no CPU fallback)\n"); exit(1); } }
	else if (hb_set_opt(g_hb, &o)) hb_die("hb_set_opt");
}

// the reads of R_INF into HBM, the filter table on the first call (raw reads: ha_ft_gen's input, Assembly.cpp:2083), the index of this pass
static void hb_load_and_index()
{
	int hom = 0, het = 0;
	hb_sync_opt();
	if (hb_reads_upload_ptrs(g_hb, R_INF.total_reads, R_INF.read_length, R_INF.read_sperate, R_INF.N_site)) hb_die("hb_reads_upload_ptrs");
	if (!g_have_ft) {
		if (ha_flt_tab) { if (hb_ft_gen(g_hb, &hom)) hb_die("hb_ft_gen"); } // (no table when the reference built none: -f / high-occ filtering disabled)
		g_have_ft = 1;
	}
	if (hb_pt_gen(g_hb, &hom, &het)) hb_die("hb_pt_gen");
	if (hom != asm_opt.hom_cov || het != asm_opt.het_cov) fprintf(stderr, "[hifiasm_b200] warning: coverage peaks differ from the reference's index (%d/%d vs %d/%d)\n", hom, het, asm_opt.hom_cov, asm_opt.het_cov);
	hb_sync_opt(); // asm_opt.hom_cov / het_cov are the reference's (ha_ec set them from its own ha_pt_gen)
}

static void flatten(ma_hit_t_alloc *a, uint64_t n, std::vector<hb_ma_hit_t> &rec, std::vector<uint64_t> &off)
{
	off.assign(n + 1, 0);
	for (uint64_t i = 0; i < n; i++) off[i + 1] = off[i] + a[i].length;
	rec.resize(off[n] + 1);
	for (uint64_t i = 0; i < n; i++) for (uint32_t k = 0; k < a[i].length; k++) {
		const ma_hit_t &s = a[i].buffer[k]; hb_ma_hit_t &d = rec[off[i] + k]; memset(&d, 0, sizeof(d));
		d.qns = s.qns; d.qe = s.qe; d.tn = s.tn; d.ts = s.ts; d.te = s.te; d.ml = s.ml; d.rev = s.rev; d.bl = s.bl; d.del = s.del; d.el = s.el; d.no_l_indel = s.no_l_indel;
	}
}
static void unflatten(const std::vector<hb_ma_hit_t> &rec, const std::vector<uint64_t> &off, ma_hit_t_alloc *a, uint64_t n)
{
	for (uint64_t i = 0; i < n; i++) {
		const uint32_t m = (uint32_t)(off[i + 1] - off[i]);
		if (m > a[i].size) { a[i].size = m; a[i].buffer = (ma_hit_t *)realloc(a[i].buffer, m * sizeof(ma_hit_t)); } // size >= length, malloc-owned (push_ff_ovlp, ecovlp.cpp:2651)
		a[i].length = m;
		for (uint32_t k = 0; k < m; k++) {
			const hb_ma_hit_t &s = rec[off[i] + k]; ma_hit_t &d = a[i].buffer[k];
			d.qns = s.qns; d.qe = s.qe; d.tn = s.tn; d.ts = s.ts; d.te = s.te; d.ml = s.ml; d.rev = s.rev; d.bl = s.bl; d.del = s.del; d.el = s.el; d.no_l_indel = s.no_l_indel; d.cc = 0;
		}
	}
}

// cal_ov_r, ecovlp.cpp:6385: replaces R_INF.paf[] / R_INF.reverse_paf[] with the final overlaps and prints ha_print_ovlp_stat_0's lines (ecovlp.cpp:6173-6180)
void cal_ov_r(uint64_t n_thre, uint64_t n_a, uint64_t new_idx)
{
	(void)n_thre; (void)new_idx;
	const double tt0 = yak_realtime_0();
	hb_load_and_index();
	std::vector<hb_ma_hit_t> p0, p1, o0, o1; std::vector<uint64_t> f0, f1, g0(n_a + 1), g1(n_a + 1); uint64_t st[8];
	flatten(R_INF.paf, R_INF.total_reads, p0, f0); flatten(R_INF.reverse_paf, R_INF.total_reads, p1, f1);
	uint64_t cap = 4 * (p0.size() + p1.size()) + 256 * n_a; int rc;
	for (int attempt = 0;; attempt++) {
		o0.resize(cap); o1.resize(cap);
		rc = hb_cal_ov_r(g_hb, 0, n_a, p0.data(), f0.data(), p1.data(), f1.data(), o0.data(), g0.data(), cap, o1.data(), g1.data(), cap, st);
		if (rc != HB_E_OVERFLOW || attempt >= 3) break;
		cap *= 4;
	}
	if (rc) hb_die("hb_cal_ov_r");
	unflatten(o0, g0, R_INF.paf, n_a); unflatten(o1, g1, R_INF.reverse_paf, n_a);
	fprintf(stderr, "[M::ha_print_ovlp_stat_0] # overlaps: %lu\n", (unsigned long)st[0]);          fprintf(stderr, "[M::ha_print_ovlp_stat_0] # strong overlaps: %lu\n", (unsigned long)st[2]);
	fprintf(stderr, "[M::ha_print_ovlp_stat_0] # weak overlaps: %lu\n", (unsigned long)st[3]);     fprintf(stderr, "[M::ha_print_ovlp_stat_0] # exact overlaps: %lu\n", (unsigned long)st[4]);
	fprintf(stderr, "[M::ha_print_ovlp_stat_0] # inexact overlaps: %lu\n", (unsigned long)st[6]);  fprintf(stderr, "[M::ha_print_ovlp_stat_0] # overlaps without large indels: %lu\n", (unsigned long)st[5]);
	fprintf(stderr, "[M::ha_print_ovlp_stat_0] # reverse overlaps: %lu\n", (unsigned long)st[1]);
	fprintf(stderr, "[M::ha_print_ovlp_stat_0] # running time: %.3f\n", yak_realtime_0() - tt0);
}

// cal_ec_r, ecovlp.cpp:6268: reads, both lists and the two read flags in R_INF, *tot_b / *tot_e, the [M::pec] lines (ecovlp.cpp:6089, 6180)
void cal_ec_r(uint64_t n_thre, uint64_t round, uint64_t n_round, uint64_t n_a, uint64_t is_sv, uint64_t *tot_b, uint64_t *tot_e)
{
	(void)n_thre;
	const double tt0 = yak_realtime_0();
	hb_load_and_index();
	std::vector<hb_ma_hit_t> p0, o0, o1; std::vector<uint64_t> f0, g0(n_a + 1), g1(n_a + 1); std::vector<uint8_t> fl(2 * n_a + 2), st(n_a + 1);
	flatten(R_INF.paf, n_a, p0, f0);                                             // gen_hc_r_alin_ea reads the previous round's paf[i] (ecovlp.cpp:3288)
	uint64_t cap = p0.size() * 2 + 128 * n_a + 4096, n_ex = 0, n_inex = 0; int rc;
	for (int attempt = 0;; attempt++) {
		o0.resize(cap); o1.resize(cap);
		rc = hb_cal_ec_r(g_hb, round, n_round, is_sv, asm_opt.max_ov_diff_ec, WINDOW_HC, p0.data(), f0.data(), o0.data(), g0.data(), cap, o1.data(), g1.data(), cap, fl.data(), st.data(), tot_b, tot_e, &n_ex, &n_inex);
		if (rc != HB_E_OVERFLOW || attempt >= 3) break;
		cap *= 4;
	}
	if (rc) hb_die("hb_cal_ec_r");
	for (uint64_t i = 0; i < n_a; i++) if (st[i]) { fprintf(stderr, "[hifiasm_b200] read %lu could not be finished on the device (status %d)\n", (unsigned long)i, (int)st[i]); exit(1); }
	fprintf(stderr, "[M::pec::%.3f] # bases: %lu; # corrected bases: %lu\n", yak_realtime_0() - tt0, (unsigned long)*tot_b, (unsigned long)*tot_e);   // ecovlp.cpp:6089
	fprintf(stderr, "[M::pec::%.3f] # exact o: %lu; # non-exact o: %lu\n", yak_realtime_0() - tt0, (unsigned long)n_ex, (unsigned long)n_inex);       // ecovlp.cpp:6108
	unflatten(o0, g0, R_INF.paf, n_a); unflatten(o1, g1, R_INF.reverse_paf, n_a);
	for (uint64_t i = 0; i < n_a; i++) { R_INF.paf[i].is_fully_corrected = fl[2 * i]; R_INF.paf[i].is_abnormal = fl[2 * i + 1]; R_INF.trio_flag[i] = AMBIGU; }
	// corrected reads back into R_INF (malloc-owned, read_size >= length: worker_sl_ec's realloc rule, ecovlp.cpp:6017)
	std::vector<uint64_t> len(n_a), noff(n_a + 1);
	if (hb_reads_download(g_hb, len.data(), 0, 0, noff.data(), 0, 0)) hb_die("hb_reads_download");
	uint64_t pb = 0; for (uint64_t i = 0; i < n_a; i++) pb += len[i] / 4 + 1;
	std::vector<uint8_t> pk(pb + 8); std::vector<uint64_t> npos(noff[n_a] + 1);
	if (hb_reads_download(g_hb, 0, pk.data(), pk.size(), 0, npos.data(), npos.size())) hb_die("hb_reads_download");
	uint64_t tb = 0;
	for (uint64_t i = 0, o = 0; i < n_a; i++) {
		if (R_INF.read_size[i] < len[i]) { R_INF.read_size[i] = len[i]; R_INF.read_sperate[i] = (uint8_t *)realloc(R_INF.read_sperate[i], len[i] / 4 + 1); }
		R_INF.read_length[i] = len[i]; memcpy(R_INF.read_sperate[i], pk.data() + o, len[i] / 4 + 1); o += len[i] / 4 + 1; tb += len[i];
		free(R_INF.N_site[i]); R_INF.N_site[i] = NULL;
		if (uint64_t c = noff[i + 1] - noff[i]) { R_INF.N_site[i] = (uint64_t *)malloc((c + 1) * 8); R_INF.N_site[i][0] = c; memcpy(R_INF.N_site[i] + 1, npos.data() + noff[i], c * 8); }
	}
	(void)tb;
}
