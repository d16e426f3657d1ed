// hb_internal.h — context and host-side helpers shared by engine.cu / index.cu.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "hb_common.cuh"

#define HB_CUDA(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { hb_set_err(ctx, HB_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); return HB_E_CUDA; } } while (0)

struct ProfEntry { const char *name; uint64_t launches; double ms; };

#define HB_MAX_LANES 4
struct hb_ctx {
	int device; cudaStream_t stream; hb_opt_t opt; std::string err;
	int sm_count;
	// read store
	uint64_t n_reads, total_bases, packed_bytes, n_npos;
	uint8_t *d_packed; uint64_t *d_roff; uint32_t *d_rlen; uint64_t *d_noff; uint32_t *d_npos;
	std::vector<uint32_t> h_rlen;
	// filter table
	uint64_t ft_n, ft_cap; uint64_t *d_ft_key; int32_t *d_ft_val;
	// position index
	uint64_t pt_keys, pt_npos, pt_cap; ulonglong2 *d_pt_slot; uint64_t *d_pt_pos;
	// staged previous overlaps for the resident final pass
	hb_ma_hit_t *d_prev0, *d_prev1; uint64_t *d_prev0_off, *d_prev1_off; uint64_t n_prev0, n_prev1;
	std::vector<uint64_t> h_prev0_off, h_prev1_off;
	// edit scripts of the current EC round (scc.a[i], ecovlp.cpp:101), staged by hb_ec_stage_scc
	uint16_t *d_scc; uint64_t *d_scc_off; uint64_t scc_reads, scc_total;
	// results of the last final pass (device resident)
	hb_ma_hit_t *d_out0, *d_out1; uint64_t *d_out0_off, *d_out1_off; uint64_t n_out0, n_out1, out_reads;
	// instrumentation
	std::vector<ProfEntry> prof, prof_stage; uint64_t counters[12], stage_counters[12]; int in_stage; // prof_stage / stage_counters: sums over all passes of the running hb_stage_run
	uint64_t anchor_budget; // anchors per batch
	uint64_t ecb_path_words; int32_t ecb_cig_words; // step B: trace words per warp of alignment tier 1, cigar words of the first merge launch (HB_ECB_PATH_WORDS / HB_ECB_CIG_WORDS, read once in hb_create: the tests shrink them so that the deferral paths run)
	int ft_chunk_bits; // HB_FT_CHUNK_BITS: exact k-mer counting in 2^bits hash ranges (-1 = as many as the free memory asks for), read once in hb_create
	int no_kmer_flt; // HB_NO_KMER_FLT (hifiasm --no-kmer-flt), read once in hb_create
	int trace, trace_ec; // HB_TRACE / HB_TRACE_EC (diagnostics), read once in hb_create
	uint32_t cns_g_nodes, cns_g_arcs; // arena of the graph consensus per warp (HB_CNS_G_NODES / HB_CNS_G_ARCS, read once in hb_create: the overflow report is tested with tiny ones)
	double last_pass_ms;
	// workspace: one device allocation used as a double-ended stack (lo: scoped scratch,
	// hi: buffers that live until the end of the pass); grown between passes only
	uint8_t *ws; size_t ws_cap, ws_lo, ws_hi, ws_need, ws_virt;
	// capacity of the resident result arrays
	uint64_t out0_cap, out1_cap, outoff_cap;
	// cached pinned staging buffer and capacities of the read-store arrays
	uint8_t *h_stage; uint64_t h_stage_cap, packed_cap, reads_cap, npos_cap;
	void *stage_buf; // hb_stage_run's host-side lists (stage.cu)
	// the sketch of the resident reads that built the position index (hb_pt_gen), kept for the pass that follows: same reads, same filter table, same
	// parameters, so the pass reads its minimizers instead of sketching every read a second time.  Invalid (sk_reads = 0) whenever the index is.
	hb_mz_t *d_sk_mz; uint64_t *d_sk_off; uint64_t sk_reads, sk_total, sk_mz_cap, sk_off_cap; std::vector<uint64_t> h_sk_off; int sk_reuse;
	int n_lanes; hb_ctx *lane[3]; // batches of a pass on n_lanes streams (HB_LANES, read once in hb_create; engine.cu run_batches); lane[] = the other lanes' contexts, owned
};
void hb_lane_free(hb_ctx *ctx);
void hb_ws_release(hb_ctx *ctx); // frees the workspaces of all lanes (between passes only)
void hb_stage_buf_free(hb_ctx *ctx);
struct GroupDir;
int hb_group_sort(hb_ctx *ctx, uint64_t nb, uint64_t r0_batch, const uint64_t *d_aoff, uint64_t a_base, uint64_t B, const hb_hit_t *d_raw, hb_hit_t *d_hits,
                  GroupDir **dir_out, uint32_t *d_dirn, uint32_t *h_dirn, uint32_t *d_sc, int mcopy_num, int cutoff); // group.cu
int hb_sort_pairs_u32(hb_ctx *ctx, const uint32_t *d_keys, const uint32_t *d_vals_in, uint32_t *d_vals_out, uint64_t n, int bits); // group.cu
#define HB_E_WS (-100) /* internal: workspace too small, the caller grows it and reruns */
int hb_ws_grow(hb_ctx *ctx);
void hb_ws_reset(hb_ctx *ctx);
void *hb_ws_lo(hb_ctx *ctx, size_t bytes);
void *hb_ws_hi(hb_ctx *ctx, size_t bytes);
void *hb_ws_hi_packed(hb_ctx *ctx, size_t bytes);

void hb_set_err(hb_ctx *ctx, int code, const char *fmt, ...);
DevReads hb_dev_reads(const hb_ctx *ctx);
DevFt hb_dev_ft(const hb_ctx *ctx);
DevPt hb_dev_pt(const hb_ctx *ctx);

// profiling scope: cudaEvent pair on the context stream, accumulated by name
struct ProfScope {
	hb_ctx *ctx; const char *name; cudaEvent_t a, b;
	ProfScope(hb_ctx *c, const char *n);
	~ProfScope();
};
void hb_prof_reset(hb_ctx *ctx);
void hb_stage_prof_begin(hb_ctx *ctx);
void hb_stage_prof_end(hb_ctx *ctx);

// device exclusive scan helpers (CUB, stream-ordered); out has n+1 entries (out[n] = total)
int hb_scan_u32_to_u64(hb_ctx *ctx, const uint32_t *d_in, uint64_t *d_out, uint64_t n);

// batch sketch of reads [r0,r1): dense minimizer arrays + per-read offsets (device, cudaMallocAsync'd; caller frees)
struct DevSketch { hb_mz_t *mz; uint64_t *off; uint64_t total; };
int hb_run_sketch(hb_ctx *ctx, uint64_t r0, uint64_t r1, int rid_mode, DevSketch *out);
int hb_run_sketch_retry(hb_ctx *ctx, uint64_t r0, uint64_t r1, DevSketch *out); // outputs live on the workspace's hi side until hb_ws_reset
