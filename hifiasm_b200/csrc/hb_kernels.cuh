// hb_kernels.cuh — the sm_100a kernels of the overlap hot path.
//
//   k_sketch        one thread per read, candidate ring in shared memory      (a2)
//   k_probe_count   warp per read: 128-bit bucket loads of the GPU ha_pt_t    (a4)
//   k_expand        half-warp per minimizer: 128-bit loads of the packed
//                   position lists -> k_mer_hit records (THE probe kernel)     (a4/a5)
//   k_group         warp per read: shared-memory hash grouping by (target,
//                   strand) + ordered scatter — replaces the anchor radix sort (a5)
//   k_chain         thread per (query,target) group: quick check + DP          (a6)
//   k_post          thread per read: chain capping / sorting / shadow filter   (a7)
//   k_exact         warp per chain: 2-bit packed exact-overlap test            (a19)
//   k_merge         thread per read: merge with previous overlaps, emit        (a19)
//   k_ed_semi64     thread per window: banded Myers in one 64-bit register     (a8)
//
// No tensor-core work exists on this path (integer / byte streams); the design
// rules that matter are coalesced 128-bit access, shared-memory staging and
// grids sized from the SM count.
#pragma once
#include "hb_sketch.cuh"
#include "hb_final.cuh"

#define HB_FULL 0xffffffffu
static __device__ __forceinline__ int hb_lane() { return threadIdx.x & 31; }

// ----------------------------------------------------------------------------
// sketch
// ----------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_sketch(DevReads R, DevFt ft, SketchPar P, uint64_t r0, uint64_t nR, const uint64_t *__restrict__ cap_off,
                                                   hb_mz_t *mz, uint32_t *mz_l, uint32_t *mz_n, int *err)
{
	extern __shared__ uint64_t sk_smem[];
	uint64_t r = (uint64_t)blockIdx.x * BLOCK + threadIdx.x;
	if (r >= nR) return;
	RingRef<uint64_t> rx = { sk_smem + threadIdx.x, BLOCK }, rm = { sk_smem + (size_t)P.w * BLOCK + threadIdx.x, BLOCK };
	RingRef<uint32_t> rl = { (uint32_t *)(sk_smem + (size_t)2 * P.w * BLOCK) + threadIdx.x, BLOCK };
	SketchOut o; o.mz = mz + cap_off[r]; o.l = mz_l + cap_off[r]; o.cap = (uint32_t)(cap_off[r + 1] - cap_off[r]); o.n = 0; o.ovf = 0;
	hb_sketch_read(R, ft, P, r0 + r, (uint32_t)(r0 + r), rx, rm, rl, o);
	mz_n[r] = o.ovf ? 0 : o.n;
	if (o.ovf) atomicOr(err, 1);
}

// strided per-read slices -> dense array (warp per read)
__global__ void k_compact_mz(uint64_t nR, const uint64_t *__restrict__ cap_off, const uint64_t *__restrict__ off, const hb_mz_t *__restrict__ in, hb_mz_t *__restrict__ out)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	const uint4 *s = (const uint4 *)(in + cap_off[w]); uint4 *d = (uint4 *)(out + off[w]);
	uint32_t n = (uint32_t)(off[w + 1] - off[w]);
	for (uint32_t i = hb_lane(); i < n; i += 32) d[i] = s[i];
}

// ----------------------------------------------------------------------------
// probe: ha_pt_get for every query minimizer (htab.cpp:518; anchor.cpp:1011-1016)
// seeds[i] = offset<<12 | n ; spre[i] = anchors of the read before minimizer i
// ----------------------------------------------------------------------------
__global__ void k_probe_count(DevPt pt, uint64_t nR, const uint64_t *__restrict__ mz_off, const hb_mz_t *__restrict__ mz,
                              uint64_t *__restrict__ seeds, uint32_t *__restrict__ spre, uint32_t *__restrict__ a_cnt)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	int lane = hb_lane();
	uint64_t b = mz_off[w], e = mz_off[w + 1]; uint32_t run = 0;
	for (uint64_t i = b; i < e; i += 32) {
		uint64_t s = i + lane, off = 0; uint32_t c = 0;
		if (s < e) c = hb_pt_lookup(pt, mz[s].x, &off);
		uint32_t inc = c;
		for (int d = 1; d < 32; d <<= 1) { uint32_t v = __shfl_up_sync(HB_FULL, inc, d); if (lane >= d) inc += v; }
		if (s < e) { seeds[s] = off << 12 | c; spre[s] = run + inc - c; }
		run += __shfl_sync(HB_FULL, inc, 31);
	}
	if (lane == 0) a_cnt[w] = run;
}

// ----------------------------------------------------------------------------
// expand: the seed-hash probe kernel proper.  Half-warp per query minimizer:
// 16 lanes x 128-bit loads walk the minimizer's ha_idxpos_t list; every hit
// becomes a k_mer_hit (anchor.cpp:1023-1037 + 1055-1076) written with 128-bit
// stores at its final place in the read's (unsorted) anchor run.
// Algorithmic bytes: 16 (minimizer) + 8 (seed) + 4 (prefix) per minimizer,
// 8 in + 16 out per hit.
// ----------------------------------------------------------------------------
__global__ void k_expand(DevReads R, DevPt pt, uint64_t r0, uint64_t n_mz, const hb_mz_t *__restrict__ mz, const uint64_t *__restrict__ seeds,
                         const uint32_t *__restrict__ spre, const uint64_t *__restrict__ a_off, uint64_t a_base, const uint32_t *__restrict__ w_tab, hb_hit_t *__restrict__ raw)
{ // mz/seeds/spre: the batch's minimizers; a_off: anchor offsets indexed by (read - r0); a_base = a_off of the batch's first read
	uint64_t hw = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4, n_hw = ((uint64_t)gridDim.x * blockDim.x) >> 4;
	int l16 = threadIdx.x & 15;
	for (uint64_t s = hw; s < n_mz; s += n_hw) {
		uint64_t sd = __ldg(&seeds[s]); uint32_t n = (uint32_t)(sd & 0xfff);
		if (n == 0) continue;
		uint64_t off = sd >> 12, zi = __ldg(&mz[s].info);
		uint32_t zrev = HB_MZ_REV(zi), zpos = HB_MZ_POS(zi), zspan = HB_MZ_SPAN(zi), rid = HB_MZ_RID(zi);
		uint32_t cnt = __ldg(&w_tab[n]) << 8 | (zspan <= 255 ? zspan : 255);
		hb_hit_t *dst = raw + (a_off[rid - r0] - a_base) + spre[s];
		uint64_t a0 = off & ~1ULL; uint32_t head = (uint32_t)(off - a0), span = head + n;
		for (uint32_t e = 2 * l16; e < span; e += 32) {
			ulonglong2 v = __ldg((const ulonglong2 *)(pt.pos + a0 + e));
#pragma unroll
			for (int h = 0; h < 2; h++) {
				int32_t j = (int32_t)(e + h) - (int32_t)head;
				if (j < 0 || j >= (int32_t)n) continue;
				uint64_t y = h ? v.y : v.x;
				uint32_t tid = HB_MZ_RID(y), rev = zrev ^ HB_MZ_REV(y), tl = __ldg(&R.len[tid]);
				uint4 o;
				o.x = tid | rev << 31;
				o.y = rev ? tl - 1 - (HB_MZ_POS(y) + 1 - HB_MZ_SPAN(y)) : HB_MZ_POS(y);
				o.z = zpos; o.w = cnt;
				*(uint4 *)(dst + j) = o;
			}
		}
	}
}

// ----------------------------------------------------------------------------
// group: warp per read.  Groups the read's anchors by key = target<<1|strand
// with a hash table in shared memory (count -> order keys -> ordered scatter),
// which is what radix_sort_ha_an1 + per-run radix_sort_ha_an3 (anchor.cpp:
// 1046-1049) achieve on the CPU: anchors leave in key order and, inside a key,
// in query-minimizer order (the chain kernel finishes the (self_offset,offset)
// order, which only same-minimizer ties can violate).  Reads with more distinct
// targets than the table holds fall back to tables in a global arena.
// ----------------------------------------------------------------------------
#define GRP_TS 2048
#define GRP_MAXG 1024
#define GRP_EMPTY 0xffffffffu
#define GRP_WARPS 4

static __device__ __forceinline__ uint32_t grp_key(uint32_t id_strand) { return id_strand << 1 | id_strand >> 31; }
static __device__ __forceinline__ uint32_t grp_hash(uint32_t key, uint32_t mask) { return (key * 2654435761u >> 7) & mask; }

// returns slot, or GRP_EMPTY when the table is full
static __device__ __forceinline__ uint32_t grp_find_or_insert(uint32_t *keys, uint32_t mask, uint32_t key, bool *fresh)
{
	uint32_t h = grp_hash(key, mask);
	for (uint32_t p = 0; p <= mask; p++) {
		uint32_t old = atomicCAS(&keys[h], GRP_EMPTY, key);
		if (old == GRP_EMPTY) { *fresh = true; return h; }
		if (old == key) { *fresh = false; return h; }
		h = (h + 1) & mask;
	}
	return GRP_EMPTY;
}
static __device__ __forceinline__ uint32_t grp_find(const uint32_t *keys, uint32_t mask, uint32_t key)
{
	uint32_t h = grp_hash(key, mask);
	while (keys[h] != key) h = (h + 1) & mask;
	return h;
}

struct GroupArgs {
	uint64_t nR, r0; const uint64_t *a_off; uint64_t a_base; const hb_hit_t *raw; hb_hit_t *hits; // r0: id of the batch's first read; a_off indexed by batch-local read
	GroupDir *dir; uint32_t *dir_n; uint32_t dir_cap; uint32_t *sc; // chain slots per read
	uint32_t *arena; unsigned long long *arena_used; uint64_t arena_words;
	int32_t mcopy_num, mcopy_khit_cutoff; int *err;
};

#define GRP_SMEM_BYTES (GRP_WARPS * (2 * GRP_TS + 2 * GRP_MAXG) * 4)
__global__ void __launch_bounds__(GRP_WARPS * 32) k_group(GroupArgs A)
{
	extern __shared__ uint32_t grp_smem[];
	__shared__ uint32_t s_nd[GRP_WARPS];
	const int wib = threadIdx.x >> 5, lane = hb_lane();
	uint32_t *const w_smem = grp_smem + (size_t)wib * (2 * GRP_TS + 2 * GRP_MAXG);
	uint64_t r = (uint64_t)blockIdx.x * GRP_WARPS + wib;
	if (r >= A.nR) return;
	const uint64_t base = A.a_off[r] - A.a_base; const uint32_t n = (uint32_t)(A.a_off[r + 1] - A.a_off[r]), rid = (uint32_t)(A.r0 + r);
	if (n == 0) { if (lane == 0) A.sc[r] = 0; return; }
	const hb_hit_t *raw = A.raw + base; hb_hit_t *out = A.hits + base;
	uint32_t *keys = w_smem, *vals = w_smem + GRP_TS, *gk = w_smem + 2 * GRP_TS, *gc = gk + GRP_MAXG, mask = GRP_TS - 1, maxg = GRP_MAXG;
	bool use_global = false;
	for (;;) { // phase A: count anchors per key
		for (uint32_t i = lane; i <= mask; i += 32) { keys[i] = GRP_EMPTY; vals[i] = 0; }
		if (lane == 0) s_nd[wib] = 0;
		__syncwarp();
		bool ovf = false;
		for (uint32_t q = lane; q < n; q += 32) {
			bool fresh; uint32_t s = grp_find_or_insert(keys, mask, grp_key(raw[q].id_strand), &fresh);
			if (s == GRP_EMPTY) { ovf = true; continue; }
			atomicAdd(&vals[s], 1u);
			if (fresh && atomicAdd(&s_nd[wib], 1u) >= maxg) ovf = true;
		}
		__syncwarp();
		if (!__any_sync(HB_FULL, ovf)) break;
		if (use_global) { if (lane == 0) atomicOr(A.err, 2); return; } // cannot happen: global tables are sized from n
		// fall back: tables sized for n distinct keys in the global arena
		uint32_t ts = 64; while (ts < 2 * n) ts <<= 1;
		unsigned long long need = 3ull * ts, at = 0;
		if (lane == 0) at = atomicAdd(A.arena_used, need);
		at = __shfl_sync(HB_FULL, at, 0);
		if (at + need > A.arena_words) { if (lane == 0) { atomicOr(A.err, 4); A.sc[r] = 0; } return; }
		keys = A.arena + at; vals = keys + ts; gk = vals + ts; gc = gk + ts / 2; mask = ts - 1; maxg = ts / 2; use_global = true;
	}
	const uint32_t G = s_nd[wib];
	uint32_t Gp = 32; while (Gp < G) Gp <<= 1;
	__syncwarp();
	// phase B: collect the distinct keys, order them (bitonic, padded with +inf)
	if (lane == 0) s_nd[wib] = 0;
	__syncwarp();
	for (uint32_t i = lane; i <= mask; i += 32)
		if (keys[i] != GRP_EMPTY) { uint32_t p = atomicAdd(&s_nd[wib], 1u); gk[p] = keys[i]; gc[p] = vals[i]; }
	for (uint32_t i = G + lane; i < Gp; i += 32) { gk[i] = GRP_EMPTY; gc[i] = 0; }
	__syncwarp();
	for (uint32_t k = 2; k <= Gp; k <<= 1)
		for (uint32_t j = k >> 1; j > 0; j >>= 1) {
			for (uint32_t i = lane; i < Gp; i += 32) {
				uint32_t x = i ^ j;
				if (x > i) {
					uint32_t a = gk[i], b = gk[x]; bool asc = (i & k) == 0;
					if ((a > b) == asc) { gk[i] = b; gk[x] = a; uint32_t t = gc[i]; gc[i] = gc[x]; gc[x] = t; }
				}
			}
			__syncwarp();
		}
	// one directory entry per group; a group whose target is the read itself gets no
	// chain slot (anchor.cpp:1931) but is still ordered (slot = GRP_EMPTY)
	uint32_t dbase = 0;
	if (lane == 0) dbase = atomicAdd(A.dir_n, G);
	dbase = __shfl_sync(HB_FULL, dbase, 0);
	const bool dir_ok = (uint64_t)dbase + G <= A.dir_cap;
	if (!dir_ok && lane == 0) atomicOr(A.err, 8);
	uint32_t run_a = 0, run_s = 0;
	for (uint32_t g0 = 0; g0 < G; g0 += 32) {
		uint32_t g = g0 + lane, c = 0, ns = 0, key = 0;
		if (g < G) { key = gk[g]; c = gc[g]; if ((key >> 1) != rid) ns = c >= (uint32_t)A.mcopy_khit_cutoff ? A.mcopy_num : 1; }
		uint32_t ia = c, is = ns;
		for (int d = 1; d < 32; d <<= 1) {
			uint32_t va = __shfl_up_sync(HB_FULL, ia, d), vs = __shfl_up_sync(HB_FULL, is, d);
			if (lane >= d) { ia += va; is += vs; }
		}
		if (g < G) {
			uint32_t start = run_a + ia - c;
			vals[grp_find(keys, mask, key)] = start; // the key's write cursor
			if (dir_ok) { GroupDir e; e.read = (uint32_t)r; e.start = start; e.count = c; e.slot = ns ? run_s + is - ns : GRP_EMPTY; A.dir[dbase + g] = e; }
		}
		run_a += __shfl_sync(HB_FULL, ia, 31); run_s += __shfl_sync(HB_FULL, is, 31);
	}
	if (lane == 0) A.sc[r] = run_s;
	__syncwarp();
	// phase C: ordered scatter (32 anchors at a time, in minimizer order)
	for (uint32_t q0 = 0; q0 < n; q0 += 32) {
		uint32_t q = q0 + lane;
		if (q < n) {
			uint4 h = *(const uint4 *)(raw + q);
			uint32_t d = atomicAdd(&vals[grp_find(keys, mask, grp_key(h.x))], 1u);
			*(uint4 *)(out + d) = h;
		}
		__syncwarp();
	}
}

// ----------------------------------------------------------------------------
// chain: one thread per (query, target) group
// ----------------------------------------------------------------------------
struct ChainArgs {
	DevReads R; uint64_t r0; const GroupDir *dir; const uint32_t *dir_n; const uint64_t *a_off; uint64_t a_base; const uint64_t *c_off;
	hb_hit_t *hits, *chits; int32_t *f, *p, *ii; int64_t *t; hb_chain_t *ch; uint32_t *slot_read; uint64_t *fc; ChainPar P; int *err;
};
__global__ void k_chain(ChainArgs A)
{
	const uint32_t n = *A.dir_n;
	for (uint32_t g = blockIdx.x * blockDim.x + threadIdx.x; g < n; g += gridDim.x * blockDim.x) {
		GroupDir d = A.dir[g];
		uint64_t ab = A.a_off[d.read] - A.a_base + d.start;
		if (d.slot == GRP_EMPTY) { hb_order_group(A.hits + ab, (int32_t)d.count); continue; }
		uint64_t cb = A.c_off[d.read] + d.slot;
		int32_t ns = d.count >= (uint32_t)A.P.mcopy_khit_cutoff ? A.P.mcopy_num : 1;
		uint32_t tid = HB_HIT_ID(A.hits[ab]);
		FcOut fc; fc.n = 0; fc.ovf = 0; fc.buf = A.fc ? A.fc + ab + 2 * cb : 0; fc.cap = d.count + 2 * ns;
		hb_chain_group(A.hits + ab, (int32_t)d.count, A.chits + ab, ab, A.f + ab, A.p + ab, A.t + ab, A.ii + ab, A.P,
		               (int64_t)A.R.len[A.r0 + d.read], (int64_t)A.R.len[tid], A.ch + cb, ns, fc);
		for (int32_t s = 0; s < ns; s++) A.slot_read[cb + s] = d.read;
		if (fc.ovf) atomicOr(A.err, 16);
	}
}

// ----------------------------------------------------------------------------
// post: one thread per read
// ----------------------------------------------------------------------------
struct PostArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *c_off; hb_chain_t *ch; const hb_hit_t *chits; uint32_t *idx; uint32_t *n_ol; uint8_t *keep;
	uint64_t *cc; const uint64_t *cc_off; ChainPar P;
};
__global__ void k_post(PostArgs A)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= A.nR) return;
	uint64_t cb = A.c_off[r]; uint32_t ns = (uint32_t)(A.c_off[r + 1] - cb);
	uint32_t n = hb_chain_post(A.ch + cb, ns, A.chits, A.idx + cb, A.cc + A.cc_off[r], A.R.len[A.r0 + r], A.P);
	A.n_ol[r] = n;
	for (uint32_t i = 0; i < n; i++) A.keep[cb + A.idx[cb + i]] = 1;
}

// ----------------------------------------------------------------------------
// exact: warp per chain slot.  exact_ec_check (ecovlp.cpp:2803) of the query
// interval against the strand-oriented target interval, 16 bases (one 32-bit
// word of 2-bit codes) per lane per step, straight from the packed reads.
// ----------------------------------------------------------------------------
static __device__ __forceinline__ uint32_t ex_fwd16(const uint8_t *p, uint64_t pos)
{ // 16 bases starting at `pos`, first base in the top two bits
	const uint8_t *b = p + (pos >> 2); uint32_t sh = (uint32_t)(pos & 3) << 1;
	uint64_t v = (uint64_t)b[0] << 32 | (uint64_t)b[1] << 24 | (uint64_t)b[2] << 16 | (uint64_t)b[3] << 8 | b[4];
	return (uint32_t)(v >> (8 - sh));
}
static __device__ __forceinline__ uint32_t ex_rc16(uint32_t w)
{ // reverse the 16 2-bit groups and complement
	w = __brev(w); w = ((w >> 1) & 0x55555555u) | ((w & 0x55555555u) << 1);
	return ~w;
}
struct ExactArgs { DevReads R; uint64_t r0, n_slots; const hb_chain_t *ch; const uint32_t *slot_read; const uint8_t *keep; uint8_t *exact; };
__global__ void k_exact(ExactArgs A)
{
	uint64_t s = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; int lane = hb_lane();
	if (s >= A.n_slots || !A.keep[s]) return;
	const hb_chain_t c = A.ch[s];
	uint64_t qid = A.r0 + A.slot_read[s], tid = c.y_id, qs = c.x_pos_s, qe = (uint64_t)c.x_pos_e + 1, ts = c.y_pos_s, te = (uint64_t)c.y_pos_e + 1;
	int ok = 1;
	if (qe - qs != te - ts) ok = 0;
	else {
		const uint8_t *q = A.R.packed + A.R.off[qid], *t = A.R.packed + A.R.off[tid];
		uint64_t n = qe - qs, tl = A.R.len[tid], nw = n >> 4;
		for (uint64_t w = lane; w < nw && ok; w += 32) {
			uint32_t a = ex_fwd16(q, qs + (w << 4)), b;
			if (!c.y_pos_strand) b = ex_fwd16(t, ts + (w << 4));
			else b = ex_rc16(ex_fwd16(t, tl - (ts + (w << 4)) - 16));
			if (a != b) ok = 0;
		}
		for (uint64_t i = (nw << 4) + lane; i < n && ok; i += 32) {
			int a = hb_base(q, qs + i), b = c.y_pos_strand ? 3 - hb_base(t, tl - 1 - (ts + i)) : hb_base(t, ts + i);
			if (a != b) ok = 0;
		}
		ok = __all_sync(HB_FULL, ok);
		if (ok && (A.R.noff[qid] != A.R.noff[qid + 1] || A.R.noff[tid] != A.R.noff[tid + 1])) { // N bases: rare, one lane
			if (lane == 0) ok = hb_exact_seq(A.R, qid, qs, qe, tid, ts, te, (int)c.y_pos_strand);
			ok = __shfl_sync(HB_FULL, ok, 0);
		}
	}
	ok = __all_sync(HB_FULL, ok);
	if (lane == 0) A.exact[s] = (uint8_t)ok;
}

// ----------------------------------------------------------------------------
// merge: one thread per read
// ----------------------------------------------------------------------------
struct MergeArgs {
	DevReads R; uint64_t r0, nR; const uint64_t *c_off; const hb_chain_t *ch; const uint32_t *idx; const uint32_t *n_ol; const uint8_t *exact;
	hb_ma_hit_t *in0; const uint64_t *in0_off; const hb_ma_hit_t *in1; const uint64_t *in1_off; // offsets indexed by batch-local read, relative to in0/in1
	FinOv *ov; uint64_t *srt; const uint64_t *o_off; // o_off[r] = sum(slots + n0) before r
	hb_ma_hit_t *out0, *out1; uint32_t *m0, *m1; unsigned long long *stat;
};
__global__ void k_merge(MergeArgs A)
{
	uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= A.nR) return;
	uint64_t cb = A.c_off[r], ob = A.o_off[r], i0 = A.in0_off[r], i1 = A.in1_off[r];
	uint32_t n0 = (uint32_t)(A.in0_off[r + 1] - i0), n1 = (uint32_t)(A.in1_off[r + 1] - i1), m0, m1;
	unsigned long long st[7];
	hb_final_merge(A.R, (uint32_t)(A.r0 + r), A.ch + cb, A.idx + cb, A.n_ol[r], A.exact + cb, A.in0 + i0, n0, A.in1 + i1, n1,
	               A.ov + ob, A.srt + i0 + i1, A.out0 + ob, &m0, A.out1 + ob, &m1, st);
	A.m0[r] = m0; A.m1[r] = m1;
	for (int b = 0; b < 7; b++) if (st[b]) atomicAdd(&A.stat[b], st[b]);
}

// per-read slices (at o_off) -> dense arrays (at d_off); warp per read
__global__ void k_gather_ma(uint64_t nR, const uint64_t *__restrict__ o_off, const uint64_t *__restrict__ d_off, const hb_ma_hit_t *__restrict__ in, hb_ma_hit_t *__restrict__ out)
{
	uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (w >= nR) return;
	const uint4 *s = (const uint4 *)(in + o_off[w]); uint4 *d = (uint4 *)(out + d_off[w]);
	uint32_t n = (uint32_t)(d_off[w + 1] - d_off[w]) * 3; // 48-byte records = 3 x 16 B
	for (uint32_t i = hb_lane(); i < n; i += 32) d[i] = s[i];
}

// ----------------------------------------------------------------------------
// window alignment: ed_band_cal_semi_64_w_absent_diag
// (Levenshtein_distance.h:3727-3776, ed_core_64 3116-3125), one thread per
// window, the whole band (<= 63 bits) in one 64-bit register pair.
// ----------------------------------------------------------------------------
static __device__ __forceinline__ int ed_nt4(char c)
{
	switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': case 'U': case 'u': return 3; default: return 4; }
}
__global__ void k_ed_semi64(uint64_t n_cases, const char *__restrict__ pat, const uint64_t *__restrict__ pat_off, const char *__restrict__ txt, const uint64_t *__restrict__ txt_off,
                            const int32_t *__restrict__ thre_a, const int32_t *__restrict__ abs_a, int32_t *__restrict__ err_o, int32_t *__restrict__ pe_o)
{
	uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (c >= n_cases) return;
	const char *pstr = pat + pat_off[c], *tstr = txt + txt_off[c];
	int32_t pn = (int32_t)(pat_off[c + 1] - pat_off[c]), tn = (int32_t)(txt_off[c + 1] - txt_off[c]), thre = thre_a[c], abs_diag = abs_a[c];
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
	int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = thre + (thre << 1), best = INT32_MAX, pe = -1, site, ai, uge = INT32_MAX, ch;
	bool dead = pn > tn + cut || tn > pn + cut;
	if (!dead) {
		bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
		for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[ed_nt4(pstr[i])] |= mm; mm <<= 1; }
		i_bd = (thre << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
		Peq[4] = 0; mm = 1ULL << (thre << 1);
		for (i = 0; i <= tn0; i++) {
			X = Peq[ed_nt4(tstr[i])] | VN;
			D0 = ((VP + (X & VP)) ^ VP) | X;
			HN = VP & D0; HP = VN | ~(VP | D0);
			X = D0 >> 1;
			VN = X & HP; VP = HN | ~(X | HP);
			if (!(D0 & 1ULL)) { ++err; if (err > cut) { dead = true; break; } }
			if (i == tn0) break;
			Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
			++i_bd; ch = 4;
			if (i_bd < pn) ch = ed_nt4(pstr[i_bd]);
			if (ch < 4) Peq[ch] |= mm;
		}
	}
	if (!dead) {
		site = tn - 1 - abs_diag; ai = pn - tn + abs_diag;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
		if (err <= thre && err <= best) { best = err; pe = site; }
		site -= i;
		while (i < ai) {
			err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
			if (err <= thre && err <= best) { best = err; pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == best) pe = site + thre;
	}
	err_o[c] = best; pe_o[c] = pe;
}
