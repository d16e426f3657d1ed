// kern_ecb.cu — steps A and B of the alignment stage of an EC round (rows a9-a11) as their own translation unit (see hb_kernels.cuh).
#define HB_KERNELS_ECB
#include <cuda_runtime.h>
#include "hb_kernels.cuh"

void hb_k_ecaln(int slow, unsigned grid, cudaStream_t st, const EcAlnArgs &A)
{ if (slow) k_ec_overlap<<<grid, 64, 0, st>>>(A); else k_ec_overlap_fast<<<grid, 128, 0, st>>>(A); }
void hb_k_ecb(int which, unsigned grid, unsigned block, cudaStream_t st, const EcCigArgs &A)
{
	switch (which) {
	case HB_K_ECB_PREP: k_ecb_prep<<<grid, block, 0, st>>>(A); break;
	case HB_K_ECB_SEG_FAST: k_ecb_seg_fast<<<grid, block, 0, st>>>(A); break;
	case HB_K_ECB_SEG: k_ecb_seg<<<grid, block, 0, st>>>(A); break;
	case HB_K_ECB_SEG_G: k_ecb_seg_g<<<grid, block, 0, st>>>(A); break;
	case HB_K_ECB_SEG_W: k_ecb_seg_w<<<grid, block, 0, st>>>(A); break;
	default: k_ecb_merge<<<grid, block, 0, st>>>(A); break;
	}
}
