// kern_ecp.cu — phasing and window consensus of an EC round (rows a13, a14) as their own translation unit (see hb_kernels.cuh).
#define HB_KERNELS_ECP
#include <cuda_runtime.h>
#include "hb_kernels.cuh"

void hb_k_ph(int decide, unsigned grid, cudaStream_t st, const PhArgs &A)
{ if (decide) k_ph_decide<<<grid, PH_WARPS * 32, 0, st>>>(A); else k_ph_count<<<grid, PH_WARPS * 32, 0, st>>>(A); }
int hb_k_cns(int graph, unsigned grid, cudaStream_t st, const CnsArgs &A)
{
	const size_t smem = (size_t)CNS_WARPS * HB_CNS_SMEM_WORDS * 8; cudaError_t e;
	if (graph) { e = cudaFuncSetAttribute(k_ec_cns_w<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e == cudaSuccess) k_ec_cns_w<true><<<grid, CNS_WARPS * 32, smem, st>>>(A); }
	else { e = cudaFuncSetAttribute(k_ec_cns_w<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e == cudaSuccess) k_ec_cns_w<false><<<grid, CNS_WARPS * 32, smem, st>>>(A); }
	return e == cudaSuccess ? 0 : 1;
}
