// group.cu — anchors of a batch grouped by (query read, target read) with one device-wide stable radix sort.
//
// What the reference does per read with radix_sort_ha_an1 + radix_sort_ha_an3 (anchor.cpp:1046-1049) and the per-target loop of lchain_qgen_mcopy_fast
// (anchor.cpp:1928-1934): the read's anchors ordered by target, strand 0 before strand 1, query-minimizer order inside; one chaining call per target
// with both strand blocks.  k_expand leaves the anchors of a read contiguous and in query-minimizer order, so ONE stable LSD sort of the whole batch on
//      key = read (batch-local) | target id | strand        [bits(nb) + bits(n_reads) + 1 bits: 4 radix passes at benchmark scale]
// gives every read's list in exactly that order (stability keeps the minimizer order), a run-length encoding of key >> 1 gives the (read, target) groups,
// and two scans give each group its place in the read's anchor range and its chain slots.  Work is spread by anchor, not by read, so a repeat-rich read
// with 10^6 anchors costs what its anchors cost (the round-1 kernel gave a warp to each read: the heaviest read set the launch time).
#include "hb_internal.h"
#include "hb_chain.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_run_length_encode.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#define GRP_EMPTY 0xffffffffu
static inline unsigned nblk(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }
static inline int bits_for(uint64_t n) { int b = 1; while ((1ull << b) < n) b++; return b; }

__global__ void k_gs_keys(uint64_t B, uint64_t nb, const uint64_t *__restrict__ a_off, uint64_t a_base, const hb_hit_t *__restrict__ raw, int id_bits, uint64_t *__restrict__ key)
{ // a_off[0 .. nb]: anchor offsets of the batch's reads; the read of anchor i by binary search (the anchors of a read are contiguous)
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= B) return;
	uint64_t lo = 0, hi = nb; const uint64_t p = a_base + i;
	while (hi - lo > 1) { const uint64_t mid = (lo + hi) >> 1; if (a_off[mid] <= p) lo = mid; else hi = mid; }
	const uint32_t ids = __ldg(&raw[i].id_strand);
	key[i] = lo << (id_bits + 1) | (uint64_t)(ids & 0x7fffffffu) << 1 | (ids >> 31);
}
struct KeyToGroup { __host__ __device__ uint64_t operator()(uint64_t k) const { return k >> 1; } };

// per group (run of equal key >> 1): chain slots, and the per-read totals
__global__ void k_gs_slots(uint32_t n_runs, const uint64_t *__restrict__ gkey, const uint32_t *__restrict__ cnt, int id_bits, uint64_t r0, int mcopy_num, int cutoff, uint32_t *__restrict__ ns, uint32_t *__restrict__ sc, uint32_t *__restrict__ first_run)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; if (g >= n_runs) return;
	const uint64_t k = gkey[g]; const uint32_t rd = (uint32_t)(k >> id_bits), id = (uint32_t)(k & ((1ull << id_bits) - 1));
	const uint32_t s = id == (uint32_t)(r0 + rd) ? 0u : (cnt[g] >= (uint32_t)cutoff ? (uint32_t)mcopy_num : 1u); // a read's hits on itself are not chained (anchor.cpp:1931)
	ns[g] = s;
	if (s) atomicAdd(&sc[rd], s);
	if (g == 0 || (uint32_t)(gkey[g - 1] >> id_bits) != rd) first_run[rd] = g;
}
__global__ void k_gs_dir(uint32_t n_runs, const uint64_t *__restrict__ gkey, const uint32_t *__restrict__ cnt, const uint32_t *__restrict__ ns, const uint64_t *__restrict__ start, const uint64_t *__restrict__ slotg,
                         const uint32_t *__restrict__ first_run, int id_bits, const uint64_t *__restrict__ a_off, uint64_t a_base, GroupDir *__restrict__ dir)
{
	const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; if (g >= n_runs) return;
	const uint32_t rd = (uint32_t)(gkey[g] >> id_bits);
	GroupDir e; e.read = rd; e.start = (uint32_t)(start[g] - (a_off[rd] - a_base)); e.count = cnt[g];
	e.slot = ns[g] ? (uint32_t)(slotg[g] - slotg[first_run[rd]]) : GRP_EMPTY;
	dir[g] = e;
}

struct U32ToU64g { __host__ __device__ uint64_t operator()(uint32_t v) const { return v; } };

// raw -> hits (grouped); *dir_out (workspace, lives until the batch's arena is released), *h_dirn groups; sc[nb] = chain slots per read (zeroed here)
int hb_group_sort(hb_ctx *ctx, uint64_t nb, uint64_t r0_batch, const uint64_t *d_aoff, uint64_t a_base, uint64_t B, const hb_hit_t *d_raw, hb_hit_t *d_hits,
                  GroupDir **dir_out, uint32_t *d_dirn, uint32_t *h_dirn, uint32_t *d_sc, int mcopy_num, int cutoff)
{
	*dir_out = 0; *h_dirn = 0;
	HB_CUDA(cudaMemsetAsync(d_sc, 0, (nb + 1) * 4, ctx->stream)); HB_CUDA(cudaMemsetAsync(d_dirn, 0, 4, ctx->stream));
	if (!B) return HB_OK;
	const int id_bits = bits_for(ctx->n_reads + 1), end_bit = bits_for(nb + 1) + id_bits + 1;
	uint64_t *d_k0 = (uint64_t *)hb_ws_lo(ctx, (B + 1) * 8), *d_k1 = (uint64_t *)hb_ws_lo(ctx, (B + 1) * 8);
	if (!d_k0 || !d_k1) return HB_E_WS;
	{
		ProfScope ps(ctx, "k_group_keys");
		k_gs_keys<<<nblk(B, 256), 256, 0, ctx->stream>>>(B, nb, d_aoff, a_base, d_raw, id_bits, d_k0);
	}
	HB_CUDA(cudaGetLastError());
	{
		size_t tb = 0;
		HB_CUDA(cub::DeviceRadixSort::SortPairs(0, tb, d_k0, d_k1, (const uint4 *)d_raw, (uint4 *)d_hits, (int64_t)B, 0, end_bit, ctx->stream));
		void *tmp = hb_ws_lo(ctx, tb + 256); if (!tmp) return HB_E_WS;
		ProfScope ps(ctx, "sort_anchors");
		HB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, d_k0, d_k1, (const uint4 *)d_raw, (uint4 *)d_hits, (int64_t)B, 0, end_bit, ctx->stream));
	}
	// groups = runs of key >> 1 (at most one per anchor; the unique keys reuse the unsorted key array)
	uint64_t *d_gkey = d_k0; uint32_t *d_cnt = (uint32_t *)hb_ws_lo(ctx, (B + 1) * 4);
	if (!d_cnt) return HB_E_WS;
	{
		cub::TransformInputIterator<uint64_t, KeyToGroup, const uint64_t *> in(d_k1, KeyToGroup()); size_t tb = 0;
		HB_CUDA(cub::DeviceRunLengthEncode::Encode(0, tb, in, d_gkey, d_cnt, d_dirn, (int64_t)B, ctx->stream));
		void *tmp = hb_ws_lo(ctx, tb + 256); if (!tmp) return HB_E_WS;
		ProfScope ps(ctx, "k_group_runs");
		HB_CUDA(cub::DeviceRunLengthEncode::Encode(tmp, tb, in, d_gkey, d_cnt, d_dirn, (int64_t)B, ctx->stream));
	}
	uint32_t n_runs = 0; HB_CUDA(cudaMemcpyAsync(&n_runs, d_dirn, 4, cudaMemcpyDeviceToHost, ctx->stream)); HB_CUDA(cudaStreamSynchronize(ctx->stream));
	uint32_t *d_ns = (uint32_t *)hb_ws_lo(ctx, ((uint64_t)n_runs + 1) * 4), *d_first = (uint32_t *)hb_ws_lo(ctx, (nb + 1) * 4);
	uint64_t *d_start = (uint64_t *)hb_ws_lo(ctx, ((uint64_t)n_runs + 2) * 8), *d_slotg = (uint64_t *)hb_ws_lo(ctx, ((uint64_t)n_runs + 2) * 8);
	GroupDir *d_dir = (GroupDir *)hb_ws_lo(ctx, ((uint64_t)n_runs + 1) * sizeof(GroupDir));
	if (!d_ns || !d_first || !d_start || !d_slotg || !d_dir) return HB_E_WS;
	{
		ProfScope ps(ctx, "k_group_dir");
		k_gs_slots<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_gkey, d_cnt, id_bits, r0_batch, mcopy_num, cutoff, d_ns, d_sc, d_first);
		size_t tb = 0; cub::TransformInputIterator<uint64_t, U32ToU64g, const uint32_t *> ic(d_cnt, U32ToU64g()), is(d_ns, U32ToU64g());
		HB_CUDA(cub::DeviceScan::ExclusiveSum(0, tb, ic, d_start, (int64_t)n_runs, ctx->stream));
		void *tmp = hb_ws_lo(ctx, tb + 256); if (!tmp) return HB_E_WS;
		HB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, ic, d_start, (int64_t)n_runs, ctx->stream));
		HB_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, is, d_slotg, (int64_t)n_runs, ctx->stream));
		k_gs_dir<<<nblk(n_runs, 256), 256, 0, ctx->stream>>>(n_runs, d_gkey, d_cnt, d_ns, d_start, d_slotg, d_first, id_bits, d_aoff, a_base, d_dir);
	}
	HB_CUDA(cudaGetLastError());
	*dir_out = d_dir; *h_dirn = n_runs;
	return HB_OK;
}

// stable sort of (key, value) pairs of 32-bit words on the low `bits` key bits (workspace scratch; vals_out receives the values in key order)
int hb_sort_pairs_u32(hb_ctx *ctx, const uint32_t *d_keys, const uint32_t *d_vals_in, uint32_t *d_vals_out, uint64_t n, int bits)
{
	if (!n) return HB_OK;
	uint32_t *d_k2 = (uint32_t *)hb_ws_lo(ctx, (n + 1) * 4); size_t tb = 0;
	HB_CUDA(cub::DeviceRadixSort::SortPairs(0, tb, d_keys, d_k2, d_vals_in, d_vals_out, (int64_t)n, 0, bits, ctx->stream));
	void *tmp = hb_ws_lo(ctx, tb + 256); if (!d_k2 || !tmp) return HB_E_WS;
	ProfScope ps(ctx, "sort_seg_queue");
	HB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tb, d_keys, d_k2, d_vals_in, d_vals_out, (int64_t)n, 0, bits, ctx->stream));
	return HB_OK;
}
