// hb_rechain_launch.h — launch record of k_ecb_rechain (ecrechain.cu), filled by the step-B host code in engine.cu.
#pragma once
#include <cuda_runtime.h>
#include "hb_ecaln.cuh"

struct RcLaunch {
	// the batch (same arrays as EcCigArgs)
	DevReads R; uint64_t r0; const OvDesc *desc; const hb_chain_t *ch; const uint64_t *fc; const uint64_t *fc_grp_base; const hb_aln_t *aln; const hb_wl_t *wlA; const uint16_t *poolA;
	double e_rate; int32_t w_l; int gaps;
	hb_alnb_t *out; hb_wl_t *wl; const uint64_t *wl_off; uint16_t *pool; unsigned long long *pool_used; uint64_t pool_cap; int *err;
	const uint32_t *rc_q; const unsigned int *rc_n;   // the overlaps the merge kernel queued
	// per-thread scratch, thread t uses slice t of each
	uint64_t *path; uint64_t path_words; uint64_t *vec; uint16_t *cig3; int32_t cig_words;     // multi-word aligner: trace, 11 x HB_MW_MAXW vectors, 3 cigar buffers (alignment | window | gap output)
	hb_wl_t *zw; int32_t zcap; uint16_t *zc; uint64_t zc_cap; unsigned long long *zc_used;      // private copy of step A's windows + the cigars traced here
	uint64_t *path1; uint16_t *cig1;                                                           // one-word traced aligner (5 * w_l words, HB_EC_CIG_TMP runs)
	hb_hit_t *h; int32_t hcap; int64_t *t, *p; int32_t *f; int32_t *rs_b; void *rs_f;           // re-seeded hits, chaining state, radix-sort scratch (512 ints, HB_RS_STACK frames)
	double pen_gap, pen_skip;                                                                  // set_lchain_dp_op(1, E_KHIT): host expf
	unsigned blocks;                                                                           // x 32 threads
};
// bytes of scratch one thread needs (the host sizes the arena with it)
size_t hb_rechain_scratch_bytes(int32_t w_l, int32_t cig_words, int32_t zcap, uint64_t zc_cap, int32_t hcap);
cudaError_t hb_launch_ecb_rechain(const RcLaunch &L, cudaStream_t stream);
