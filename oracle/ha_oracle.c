/* oracle/ha_oracle.c — TEST INFRASTRUCTURE ONLY (see ha_oracle.h).
 *
 * CPU restatement of the overlap hot path of hifiasm v0.25.0-r726.
 * Single-threaded, simple data structures (sorted arrays + binary search
 * instead of khashl); every routine names the reference lines it restates.
 * Compile with -ffp-contract=off: chain scoring must not be FMA-contracted.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <assert.h>
#include "ha_oracle.h"

#define MALLOC_N(T, n) ((T *)malloc((size_t)((n) > 0 ? (n) : 1) * sizeof(T)))
#define CALLOC_N(T, n) ((T *)calloc((size_t)((n) > 0 ? (n) : 1), sizeof(T)))

void hao_free(void *p) { free(p); }

void hao_opt_default(hao_opt_t *o)
{ /* init_opt, CommandLines.cpp:243-380 */
	memset(o, 0, sizeof(*o));
	o->k = 51; o->w = 51; o->is_hpc = 1;
	o->mz_sample_dist = 500; o->mz_rewin = 1000;
	o->min_hist_kmer_cnt = 5; o->high_factor = 5.0; o->max_kmer_cnt = 2000;
	o->max_n_chain = 100; o->hom_cov = 20; o->het_cov = -1024;
}

void hao_opt_update_cov(hao_opt_t *o, int hom_cov)
{ /* ha_opt_update_cov, CommandLines.cpp:411-418 */
	int m = (int)(hom_cov * o->high_factor + .499);
	o->hom_cov = hom_cov;
	if (o->max_n_chain < m) o->max_n_chain = m;
}

/* ------------------------------------------------------------------ */
/* 2-bit read store                                                    */
/* ------------------------------------------------------------------ */

static inline int rd_base(const hao_reads_t *r, uint64_t id, int64_t p)
{ /* ha_compress_base layout, Process_Read.cpp:792-850 */
	uint8_t b = r->packed[r->off[id] + (uint64_t)(p >> 2)];
	return (b >> ((3 - (p & 3)) << 1)) & 3;
}

void hao_decode(const hao_reads_t *r, uint64_t id, char *out)
{ /* recover_UC_Read, Process_Read.cpp:716-760 */
	int64_t i, L = (int64_t)r->len[id];
	for (i = 0; i < L; i++) out[i] = "ACGT"[rd_base(r, id, i)];
	if (r->n_off) {
		uint64_t j;
		for (j = r->n_off[id]; j < r->n_off[id + 1]; j++) out[r->n_pos[j]] = 'N';
	}
}

void hao_decode_sub(const hao_reads_t *r, uint64_t id, int64_t start, int64_t len, int strand, char *out)
{ /* recover_UC_Read_sub_region, Process_Read.cpp:524-616: start/len are on the
     requested strand; the reverse strand is the reverse complement */
	int64_t i, L = (int64_t)r->len[id];
	if (!strand) {
		for (i = 0; i < len; i++) out[i] = "ACGT"[rd_base(r, id, start + i)];
	} else {
		for (i = 0; i < len; i++) out[i] = "TGCA"[rd_base(r, id, L - 1 - (start + i))];
	}
	if (r->n_off) {
		uint64_t j;
		for (j = r->n_off[id]; j < r->n_off[id + 1]; j++) {
			int64_t p = (int64_t)r->n_pos[j];
			if (strand) p = L - 1 - p;
			if (p >= start && p < start + len) out[p - start] = 'N';
		}
	}
}

/* ------------------------------------------------------------------ */
/* hashing                                                             */
/* ------------------------------------------------------------------ */

static inline uint64_t hash64(uint64_t key)
{ /* yak_hash64_64, htab.h:150-160 */
	key = ~key + (key << 21);
	key = key ^ key >> 24;
	key = (key + (key << 3)) + (key << 8);
	key = key ^ key >> 14;
	key = (key + (key << 2)) + (key << 4);
	key = key ^ key >> 28;
	key = key + (key << 31);
	return key;
}

static inline int nt4(char c)
{ /* seq_nt4_table, htab.cpp:17-34 */
	switch (c) {
	case 'A': case 'a': return 0;
	case 'C': case 'c': return 1;
	case 'G': case 'g': return 2;
	case 'T': case 't': case 'U': case 'u': return 3;
	default: return 4;
	}
}

/* ------------------------------------------------------------------ */
/* filter table: sorted (hash,count)                                   */
/* ------------------------------------------------------------------ */

struct hao_ft_s { uint64_t n; uint64_t *key; int16_t *val; };

int32_t hao_ft_cnt(const hao_ft_t *ft, uint64_t y)
{ /* ha_ft_cnt, htab.cpp:1064-1070 */
	uint64_t lo = 0, hi;
	if (ft == 0) return 0;
	hi = ft->n;
	while (lo < hi) {
		uint64_t mid = (lo + hi) >> 1;
		if (ft->key[mid] < y) lo = mid + 1; else hi = mid;
	}
	if (lo < ft->n && ft->key[lo] == y) return ft->val[lo] == INT16_MAX ? INT32_MAX : ft->val[lo];
	return 0;
}
uint64_t hao_ft_size(const hao_ft_t *ft) { return ft ? ft->n : 0; }
void hao_ft_destroy(hao_ft_t *ft) { if (ft) { free(ft->key); free(ft->val); free(ft); } }

/* ------------------------------------------------------------------ */
/* minimizer sketch                                                    */
/* ------------------------------------------------------------------ */

typedef struct { uint64_t x; uint32_t cnt; uint32_t pos; uint8_t rev, span; } cand_t;

static inline int cand_cmp(const cand_t *a, const cand_t *b)
{ /* mz1_mzcmp, sketch.cpp:184: by count (kept in .rid) then by hash */
	if (a->cnt != b->cnt) return a->cnt < b->cnt ? -1 : 1;
	return (a->x > b->x) - (a->x < b->x);
}

typedef struct { hao_mz_t *a; uint64_t *l; uint32_t n, m; } mzv_t;

static void mzv_push(mzv_t *v, const cand_t *c, uint32_t l)
{
	if (v->n == v->m) {
		v->m = v->m ? v->m << 1 : 256;
		v->a = (hao_mz_t *)realloc(v->a, v->m * sizeof(hao_mz_t));
		v->l = (uint64_t *)realloc(v->l, v->m * sizeof(uint64_t));
	}
	v->a[v->n].x = c->x;
	v->a[v->n].info = (uint64_t)(c->cnt & 0xfffffff) | (uint64_t)(c->pos & 0x7ffffff) << 28 | (uint64_t)(c->rev & 1) << 55 | (uint64_t)c->span << 56;
	v->l[v->n] = l;
	v->n++;
}

/* --- high-occurrence thinning: mz1_select_mz_h, sketch.cpp:194-330 --- */
#define MZ_CNT(v, i) HAO_MZ_RID((v)->a[i])
#define MZ_SETCNT0(v, i) ((v)->a[i].info &= ~0xfffffffULL)
#define MZ_L(v, i) ((int64_t)(uint32_t)(v)->l[i])
#define MZ_ACT(v, i) ((i) >= 0 && MZ_CNT(v, i) > 0)

static inline int mz_cmp(const hao_mz_t *a, const hao_mz_t *b)
{
	uint32_t ca = HAO_MZ_RID(*a), cb = HAO_MZ_RID(*b);
	if (ca != cb) return ca < cb ? -1 : 1;
	return (a->x > b->x) - (a->x < b->x);
}
static inline int mz_cmp_l(const mzv_t *v, int32_t ai, int32_t bi)
{ /* mz1_mzcmp_l, sketch.cpp:217-225 */
	if (ai >= 0 && bi >= 0) {
		const hao_mz_t *a = &v->a[ai], *b = &v->a[bi];
		if (HAO_MZ_RID(*a) > 0 && HAO_MZ_RID(*b) > 0) return mz_cmp(a, b);
		return (HAO_MZ_RID(*a) == 0) - (HAO_MZ_RID(*b) == 0);
	}
	return (ai < 0) - (bi < 0);
}
#define MARK 0x100000000ULL

typedef struct { hao_mz_t m; int32_t idx; } hsel_t;
static inline int hsel_lt(const hsel_t *a, const hsel_t *b) { return mz_cmp(&a->m, &b->m) < 0; }
static void hsel_down(size_t i, size_t n, hsel_t *l)
{ /* ks_heapdown (ksort.h:43-53) on mz ordering */
	size_t k = i; hsel_t tmp = l[i];
	while ((k = (k << 1) + 1) < n) {
		if (k != n - 1 && hsel_lt(&l[k], &l[k + 1])) ++k;
		if (hsel_lt(&l[k], &tmp)) break;
		l[i] = l[k]; i = k;
	}
	l[i] = tmp;
}
static void hf_select(mzv_t *p, int32_t si, int32_t ei, int32_t n, int32_t len, int32_t sample_dist)
{ /* mz1_hf_select, sketch.cpp:194-216 (force=0) */
	hsel_t b[16];
	int32_t ps, pe, j, k, max_occ; size_t i;
	if (ei - si <= 1) return;
	ps = si < 0 ? 0 : (int32_t)HAO_MZ_POS(p->a[si]);
	pe = ei == n ? len : (int32_t)HAO_MZ_POS(p->a[ei]);
	max_occ = (int32_t)((double)(pe - ps) / sample_dist + .499);
	if (max_occ > 16) max_occ = 16;
	for (j = si + 1, k = 0; j < ei && k < max_occ; ++j, ++k) b[k].m = p->a[j], b[k].idx = j;
	for (i = ((size_t)k >> 1) - 1; i != (size_t)(-1); --i) hsel_down(i, k, b);
	for (; j < ei; ++j) {
		hsel_t c; c.m = p->a[j]; c.idx = j;
		if (hsel_lt(&c, &b[0])) { b[0] = c; hsel_down(0, k, b); }
	}
	for (j = 0; j < k; ++j)
		if ((int32_t)HAO_MZ_RID(b[j].m) < pe - ps) MZ_SETCNT0(p, b[j].idx);
}

static void rescan_win(mzv_t *p, int32_t si, int32_t i, int32_t *mi, int skip_inactive_first)
{ /* the "find new min in [si,i] then mark all equal" block used four times in
     mz1_select_mz_h / mz1_qfw (sketch.cpp:232-241, 282-290, 297-305) */
	int32_t m;
	for (m = si, *mi = -1; m <= i; m++) {
		if (skip_inactive_first && !MZ_ACT(p, m)) continue;
		if (mz_cmp_l(p, *mi, m) >= 0) *mi = m;
	}
	if (MZ_ACT(p, *mi)) {
		for (m = si; m <= i; m++) {
			if (!MZ_ACT(p, m)) continue;
			if (mz_cmp_l(p, *mi, m) == 0) p->l[m] |= MARK;
		}
	}
}

static void select_mz_h(mzv_t *p, int len, int sample_dist, int32_t w, int32_t k, int32_t tot_l)
{ /* mz1_select_mz_h, sketch.cpp:247-330 */
	int32_t i, mi = -1, si, last0, n = (int32_t)p->n, m = 0, ws = w + k - 1, any = 0, dense = 0;
	if (n == 0) return;
	for (i = 0, last0 = -1; i <= n; ++i) { /* 252-265 */
		if (i == n || MZ_CNT(p, i) == 0) {
			if (i - last0 > 1) {
				int32_t ps = last0 < 0 ? 0 : (int32_t)HAO_MZ_POS(p->a[last0]);
				int32_t pe = i == n ? len : (int32_t)HAO_MZ_POS(p->a[i]);
				if (((int32_t)((double)(pe - ps) / sample_dist + .499)) > 0) { dense = 1; any = 1; break; }
			}
			last0 = i;
		}
	}
	if (!any) return;      /* 266: no high-frequency streak: nothing is removed */
	(void)dense;
	/* 268-270: mz1_qfw — first window */
	for (si = 0, i = 0, mi = -1; i < n; i++) {
		if (MZ_L(p, i) >= ws || (i + 1 < n && MZ_L(p, i) < ws && MZ_L(p, i + 1) > ws) ||
		    (i + 1 == n && tot_l >= ws && MZ_L(p, i) < ws)) {
			rescan_win(p, si, i, &mi, 1);
			break;
		}
	}
	if (i == n) goto squeeze;
	for (si = 0, i++; i < n; i++) { /* 271-292 */
		for (; si < i; si++)
			if (MZ_L(p, si) + w > MZ_L(p, i)) break;
		if (mz_cmp_l(p, i, mi) <= 0) {
			if (MZ_ACT(p, mi)) p->l[mi] |= MARK;
			mi = i;
		} else if (si > mi) {
			if (MZ_ACT(p, mi)) p->l[mi] |= MARK;
			rescan_win(p, si, i, &mi, 0);
		}
	}
	if (MZ_ACT(p, mi)) p->l[mi] |= MARK; /* 293 */
	for (i = n - 1; si < n && MZ_L(p, si) + w <= tot_l + 1; si++) { /* 294-307 */
		if (si > mi) {
			if (MZ_ACT(p, mi)) p->l[mi] |= MARK;
			rescan_win(p, si, i, &mi, 0);
		}
	}
	for (i = 0, last0 = -1; i <= n; ++i) { /* 310-324 */
		if (i == n || MZ_CNT(p, i) == 0) {
			if (i - last0 > 1) {
				int32_t ps = last0 < 0 ? 0 : (int32_t)HAO_MZ_POS(p->a[last0]);
				int32_t pe = i == n ? len : (int32_t)HAO_MZ_POS(p->a[i]);
				if (((int32_t)((double)(pe - ps) / sample_dist + .499)) > 0) {
					for (m = last0 + 1, mi = 0; m < i; ++m)
						if (p->l[m] & MARK) MZ_SETCNT0(p, m), mi++;
					if (mi == 0) hf_select(p, last0, i, n, len, sample_dist);
				}
			}
			last0 = i;
		}
	}
squeeze:
	for (i = n = 0; i < (int32_t)p->n; ++i) /* 326-329 */
		if (MZ_CNT(p, i) == 0) p->a[n++] = p->a[i];
	p->n = n;
}

int hao_sketch(const char *s, int len, int w, int k, uint32_t rid, int is_hpc, const hao_ft_t *ft,
               int sample_dist, int rewin, hao_mz_t **out, uint32_t *n_out)
{ /* mz1_ha_sketch, sketch.cpp:454-579 */
	static const cand_t dummy = { UINT64_MAX, (1u << 28) - 1, 0, 0, 0 };
	uint64_t shift1 = k - 1, mask = (1ULL << k) - 1, pl[4] = { 0, 0, 0, 0 };
	int i, j, l = 0, tl = 0, bp = 0, min_bp = 0, span = 0;
	int q[64], qf = 0, qc = 0; /* tiny_queue_t, htab.h:39-57 */
	cand_t ring[256], min = dummy; uint32_t ring_l[256], min_l = (uint32_t)-1;
	mzv_t v; memset(&v, 0, sizeof(v));
	if (!(len > 0 && w > 0 && w < 256 && k > 0 && k <= 63)) return -1;
	memset(ring, 0xff, sizeof(ring)); /* sketch.cpp:470 */
	for (j = 0; j < w; j++) ring[j].cnt = (1u << 28) - 1, ring[j].pos = (1u << 27) - 1, ring[j].rev = 1, ring[j].span = 255;
	for (i = 0; i < len; ++i) {
		int c = nt4(s[i]);
		cand_t info = dummy;
		if (c < 4) {
			int z;
			if (is_hpc) { /* 480-492 */
				int run = 1;
				while (i + run < len && nt4(s[i + run]) == c) ++run;
				i += run - 1;
				q[(qc++ + qf) & 0x3f] = run;
				span += run;
				if (qc > k) { span -= q[qf++]; qf &= 0x3f; --qc; }
			} else span = l + 1 < k ? l + 1 : k;
			pl[0] = (pl[0] << 1 | (uint64_t)(c & 1)) & mask;
			pl[1] = (pl[1] << 1 | (uint64_t)(c >> 1)) & mask;
			pl[2] = pl[2] >> 1 | (uint64_t)(1 - (c & 1)) << shift1;
			pl[3] = pl[3] >> 1 | (uint64_t)(1 - (c >> 1)) << shift1;
			if (pl[1] == pl[3]) continue; /* 502: strand unknown; nothing else advances */
			z = pl[1] < pl[3] ? 0 : 1;
			++l; ++tl;
			if (l >= k && span < 256) {
				uint64_t y = hash64(pl[z << 1 | 0]) + hash64(pl[z << 1 | 1]);
				int32_t cnt = ft ? hao_ft_cnt(ft, y) : 0;
				if (!(cnt >= 1 << 28)) { info.x = y; info.cnt = (uint32_t)cnt; info.pos = (uint32_t)i; info.rev = (uint8_t)z; info.span = (uint8_t)span; }
			}
		} else l = 0, qc = qf = 0, span = 0;
		ring[bp] = info; ring_l[bp] = (uint32_t)l;
		if (l == w + k - 1 && min.x != UINT64_MAX) { /* 523-534 */
			for (j = bp + 1; j < w; ++j)
				if (cand_cmp(&min, &ring[j]) == 0 && ring[j].pos != min.pos) mzv_push(&v, &ring[j], ring_l[j]);
			for (j = 0; j < bp; ++j)
				if (cand_cmp(&min, &ring[j]) == 0 && ring[j].pos != min.pos) mzv_push(&v, &ring[j], ring_l[j]);
		}
		if (cand_cmp(&min, &info) >= 0) { /* 543-547 */
			if (l >= w + k && min.x != UINT64_MAX) mzv_push(&v, &min, min_l);
			min = info; min_bp = bp; min_l = ring_l[bp];
		} else if (bp == min_bp) { /* 548-568 */
			if (l >= w + k - 1 && min.x != UINT64_MAX) mzv_push(&v, &min, min_l);
			for (j = bp + 1, min = dummy; j < w; ++j)
				if (cand_cmp(&min, &ring[j]) >= 0) min = ring[j], min_bp = j, min_l = ring_l[j];
			for (j = 0; j <= bp; ++j)
				if (cand_cmp(&min, &ring[j]) >= 0) min = ring[j], min_bp = j, min_l = ring_l[j];
			if (l >= w + k - 1 && min.x != UINT64_MAX) {
				for (j = bp + 1; j < w; ++j)
					if (cand_cmp(&min, &ring[j]) == 0 && min.pos != ring[j].pos) mzv_push(&v, &ring[j], ring_l[j]);
				for (j = 0; j <= bp; ++j)
					if (cand_cmp(&min, &ring[j]) == 0 && min.pos != ring[j].pos) mzv_push(&v, &ring[j], ring_l[j]);
			}
		}
		if (++bp == w) bp = 0;
	}
	if (min.x != UINT64_MAX) mzv_push(&v, &min, min_l);
	if (sample_dist > w) select_mz_h(&v, len, sample_dist, rewin, k, tl); /* 575 */
	for (i = 0; i < (int)v.n; ++i) v.a[i].info = (v.a[i].info & ~0xfffffffULL) | (rid & 0xfffffff); /* 577 */
	free(v.l);
	*out = v.a; *n_out = v.n;
	return 0;
}

/* ------------------------------------------------------------------ */
/* k-mer histogram peaks                                               */
/* ------------------------------------------------------------------ */

static int adj_peak_hom(int m_peak_hom, int max_i, int max2_i, int max3_i, int *peak_het)
{ /* adj_m_peak_hom, hist.cpp:46-72 */
	int64_t mm[3], d, min_i = -1, min_d = -1, i;
	mm[0] = max2_i; mm[1] = max_i; mm[2] = max3_i;
	for (i = 0; i < 3; i++) {
		if (mm[i] <= 0) continue;
		d = mm[i] >= m_peak_hom ? mm[i] - m_peak_hom : m_peak_hom - mm[i];
		if (min_d == -1 || min_d > d || (min_d == d && i == 1)) min_d = d, min_i = i;
	}
	if (min_i < 0) return m_peak_hom;
	if (mm[min_i] < m_peak_hom) {
		d = m_peak_hom - mm[min_i];
		if (d >= mm[min_i] * 0.51) { *peak_het = (int)mm[min_i]; return m_peak_hom; }
	}
	for (i = min_i - 1; i >= 0; i--) {
		if (mm[i] <= 0) continue;
		*peak_het = (int)mm[i];
		break;
	}
	return (int)mm[min_i];
}

int hao_analyze_count(int n_cnt, int start_cnt, int m_peak_hom, const int64_t *cnt, int *peak_het)
{ /* ha_analyze_count, hist.cpp:74-157 (printing dropped) */
	int i, start, low_i, max_i, max2_i, max3_i; int64_t max, max2, max3, min;
	*peak_het = -1;
	start = cnt[1] > 0 ? 1 : 2;
	low_i = start > start_cnt ? start : start_cnt;
	for (i = low_i + 1; i < n_cnt; ++i)
		if (cnt[i] > cnt[i - 1]) break;
	low_i = i - 1;
	if (low_i == n_cnt - 1) return -1;
	max_i = low_i + 1, max = cnt[max_i];
	for (i = low_i + 1; i < n_cnt; ++i)
		if (cnt[i] > max) max = cnt[i], max_i = i;
	max2 = -1; max2_i = -1;
	for (i = max_i - 1; i > low_i; --i)
		if (cnt[i] >= cnt[i - 1] && cnt[i] >= cnt[i + 1])
			if (cnt[i] > max2) max2 = cnt[i], max2_i = i;
	if (max2_i > low_i && max2_i < max_i) {
		for (i = max2_i + 1, min = max; i < max_i; ++i)
			if (cnt[i] < min) min = cnt[i];
		if (max2 < max * 0.05 || min > max2 * 0.95) max2 = -1, max2_i = -1;
	}
	max3 = -1; max3_i = -1;
	for (i = max_i + 1; i < n_cnt - 1; ++i)
		if (cnt[i] >= cnt[i - 1] && cnt[i] >= cnt[i + 1])
			if (cnt[i] > max3) max3 = cnt[i], max3_i = i;
	if (max3_i > max_i) {
		for (i = max_i + 1, min = max; i < max3_i; ++i)
			if (cnt[i] < min) min = cnt[i];
		if (max3 < max * 0.05 || min > max3 * 0.95 || max3_i > max_i * 2.5) max3 = -1, max3_i = -1;
	}
	if (m_peak_hom > 0) return adj_peak_hom(m_peak_hom, max_i, max2_i, max3_i, peak_het);
	if (max3_i > 0) { *peak_het = max_i; return max3_i; }
	if (max2_i > 0) *peak_het = max2_i;
	return max_i;
}

/* ------------------------------------------------------------------ */
/* counting helpers                                                    */
/* ------------------------------------------------------------------ */

static void radix_sort_u64(uint64_t *a, uint64_t *tmp, uint64_t n)
{ /* plain LSD radix sort (stable) */
	int pass;
	for (pass = 0; pass < 8; pass++) {
		uint64_t c[257], i; int sh = pass * 8;
		memset(c, 0, sizeof(c));
		for (i = 0; i < n; i++) c[((a[i] >> sh) & 0xff) + 1]++;
		for (i = 0; i < 256; i++) c[i + 1] += c[i];
		for (i = 0; i < n; i++) tmp[c[(a[i] >> sh) & 0xff]++] = a[i];
		memcpy(a, tmp, n * sizeof(uint64_t));
	}
}

typedef struct { uint64_t *a; uint64_t n, m; } u64v_t;
static inline void u64v_push(u64v_t *v, uint64_t x)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 1 << 16; v->a = (uint64_t *)realloc(v->a, v->m * 8); }
	v->a[v->n++] = x;
}

static void count_hpc_kmers(u64v_t *v, int k, int len, const char *seq, int is_hpc)
{ /* mz1_count_seq_buf_HPC / mz1_count_seq_buf, htab.cpp:608-645: every k-mer,
     hashed by yak_hash_long (htab.h:162-167) */
	int i, l, last = -1; uint64_t x[4] = { 0, 0, 0, 0 }, mask = (1ULL << k) - 1, shift = k - 1;
	for (i = l = 0; i < len; ++i) {
		int c = nt4(seq[i]);
		if (c < 4) {
			if (!is_hpc || c != last) {
				x[0] = (x[0] << 1 | (uint64_t)(c & 1)) & mask;
				x[1] = (x[1] << 1 | (uint64_t)(c >> 1)) & mask;
				x[2] = x[2] >> 1 | (uint64_t)(1 - (c & 1)) << shift;
				x[3] = x[3] >> 1 | (uint64_t)(1 - (c >> 1)) << shift;
				if (++l >= k) {
					int j = x[1] < x[3] ? 0 : 1;
					u64v_push(v, hash64(x[j << 1 | 0]) + hash64(x[j << 1 | 1]));
				}
				last = c;
			}
		} else l = 0, last = -1, x[0] = x[1] = x[2] = x[3] = 0;
	}
}

/* yak_bf_insert, htab.cpp:98-115: blocked Bloom filter, one 512-bit block per k-mer, n_hashes bits; returns how many of them were set */
static int bf_insert(uint8_t *b, int n_shift, int n_hashes, uint64_t hash)
{
	int x = n_shift - 9, i, z, cnt = 0;
	uint64_t y = hash & ((1ULL << x) - 1);
	int h1 = (int)(hash >> x & 511), h2 = (int)(hash >> n_shift & 511);
	uint8_t *p = &b[y << 6];
	if ((h2 & 31) == 0) h2 = (h2 + 1) & 511;
	for (i = 0, z = h1; i < n_hashes; z = (z + h2) & 511) {
		uint8_t *q = &p[z >> 3], u = (uint8_t)(1 << (z & 7));
		cnt += !!(*q & u); *q |= u; ++i;
	}
	return cnt;
}

hao_ft_t *hao_ft_gen_bf(const hao_reads_t *r, const hao_opt_t *o, int bf_shift, int *hom_cov)
{ /* ha_ft_gen, htab.cpp:1136-1169.  bf_shift = asm_opt.bf_shift (-f): with bf_shift - 12 >= 9 every one of the 4096 sub-tables has a blocked
     Bloom filter of 2^(bf_shift-12) bits (ha_ct_init 140-158, yak_bf_init 78-91) and ha_ct_insert_list (181-214) only counts an occurrence
     whose 4 bits were all set before — the first occurrence of a k-mer is lost unless it is a false positive — and a k-mer enters the
     table with count 1 + 1.  The k-mers go through it in read order, position order (the pipeline fills the 4096 buffers sequentially,
     826-833, and each buffer is inserted by one thread, 690-698), so the counts are a function of the input alone. */
	u64v_t v = { 0, 0, 0 }; uint64_t i, j, maxl = 0, *tmp; char *buf;
	int64_t cnt[4096]; int peak_hom, peak_het, cutoff, max_cnt;
	const int bf_local = bf_shift - 12, use_bf = bf_shift > 12 && bf_local >= 9 && bf_local + 9 <= 64;
	hao_ft_t *ft = CALLOC_N(hao_ft_t, 1);
	for (i = 0; i < r->n; i++) if (r->len[i] > maxl) maxl = r->len[i];
	buf = MALLOC_N(char, maxl + 1);
	for (i = 0; i < r->n; i++) {
		hao_decode(r, i, buf);
		count_hpc_kmers(&v, o->k, (int)r->len[i], buf, o->is_hpc);
	}
	free(buf);
	if (use_bf) { /* keep the occurrences the filter lets through; one extra copy of each k-mer that got in stands for the initial count of 1 */
		const uint64_t sub_bytes = 1ULL << (bf_local - 3); uint64_t m = 0;
		uint8_t *bf = (uint8_t *)calloc(4096, sub_bytes);
		for (i = 0; i < v.n; i++)
			if (bf_insert(bf + (v.a[i] & 0xfff) * sub_bytes, bf_local, 4, v.a[i] >> 12) == 4) v.a[m++] = v.a[i];
		free(bf);
		v.n = m;
	}
	tmp = MALLOC_N(uint64_t, v.n);
	radix_sort_u64(v.a, tmp, v.n);
	free(tmp);
	memset(cnt, 0, sizeof(cnt));
	for (i = 0; i < v.n; i = j) { /* ha_ct_hist, htab.cpp:240; counts saturate at 4095 (181-214) */
		for (j = i + 1; j < v.n && v.a[j] == v.a[i]; j++) {}
		cnt[j - i + use_bf > 4095 ? 4095 : j - i + use_bf]++;
	}
	peak_hom = hao_analyze_count(4096, o->min_hist_kmer_cnt, -1, cnt, &peak_het);
	if (hom_cov) *hom_cov = peak_hom;
	cutoff = (int)(peak_hom * o->high_factor);
	if (cutoff > 4094) cutoff = 4094;
	max_cnt = o->max_kmer_cnt; /* gen_hh, htab.cpp:1038-1062 */
	if (max_cnt > 4094) max_cnt = 4094;
	ft->key = MALLOC_N(uint64_t, v.n); ft->val = MALLOC_N(int16_t, v.n);
	for (i = 0; i < v.n; i = j) {
		int c;
		for (j = i + 1; j < v.n && v.a[j] == v.a[i]; j++) {}
		c = j - i + use_bf > 4095 ? 4095 : (int)(j - i + use_bf);
		if (c >= cutoff && c <= 4095) { /* ha_ct_shrink(h, cutoff, YAK_MAX_COUNT) */
			ft->key[ft->n] = v.a[i];
			ft->val[ft->n] = c > max_cnt ? INT16_MAX : (int16_t)c;
			ft->n++;
		}
	}
	free(v.a);
	return ft;
}
hao_ft_t *hao_ft_gen(const hao_reads_t *r, const hao_opt_t *o, int *hom_cov) { return hao_ft_gen_bf(r, o, 0, hom_cov); } /* exact counting (-f0) */

/* every HPC k-mer hash of the store in read / position order (test tooling: the probes of the filter-table parity tests) */
uint64_t hao_all_kmers(const hao_reads_t *r, const hao_opt_t *o, uint64_t *out, uint64_t cap)
{
	u64v_t v = { 0, 0, 0 }; uint64_t i, maxl = 0, n; char *buf;
	for (i = 0; i < r->n; i++) if (r->len[i] > maxl) maxl = r->len[i];
	buf = MALLOC_N(char, maxl + 1);
	for (i = 0; i < r->n; i++) { hao_decode(r, i, buf); count_hpc_kmers(&v, o->k, (int)r->len[i], buf, o->is_hpc); }
	free(buf);
	n = v.n;
	if (out) memcpy(out, v.a, (n < cap ? n : cap) * 8);
	free(v.a);
	return n;
}

/* ------------------------------------------------------------------ */
/* position index                                                      */
/* ------------------------------------------------------------------ */

struct hao_pt_s { uint64_t n_keys, tot_pos; uint64_t *key, *off; uint32_t *cnt; uint64_t *pos; };

typedef struct { uint64_t x, info, ord; } mzo_t;
static int mzo_cmp(const void *a, const void *b)
{
	const mzo_t *p = (const mzo_t *)a, *q = (const mzo_t *)b;
	if (p->x != q->x) return p->x < q->x ? -1 : 1;
	return (p->ord > q->ord) - (p->ord < q->ord);
}

hao_pt_t *hao_pt_gen(const hao_reads_t *r, const hao_ft_t *ft, const hao_opt_t *o, int *hom_cov, int *het_cov)
{ /* ha_pt_gen, htab.cpp:1232-1287 (normal mode, reads from the store) */
	uint64_t i, j, n = 0, m = 0, maxl = 0, nk = 0, np = 0; mzo_t *a = 0; char *buf;
	int64_t cnt[4096]; int peak_hom, peak_het, lo = 2, hi = 4094;
	hao_pt_t *pt = CALLOC_N(hao_pt_t, 1);
	for (i = 0; i < r->n; i++) if (r->len[i] > maxl) maxl = r->len[i];
	buf = MALLOC_N(char, maxl + 1);
	for (i = 0; i < r->n; i++) { /* mz1_worker_for_mz, htab.cpp:685-696 */
		hao_mz_t *mz; uint32_t nm, t;
		hao_decode(r, i, buf);
		hao_sketch(buf, (int)r->len[i], o->w, o->k, (uint32_t)i, o->is_hpc, ft, o->mz_sample_dist, o->mz_rewin, &mz, &nm);
		if (n + nm > m) { m = (n + nm) * 2; a = (mzo_t *)realloc(a, m * sizeof(mzo_t)); }
		for (t = 0; t < nm; t++) { a[n].x = mz[t].x; a[n].info = mz[t].info; a[n].ord = n; n++; }
		free(mz);
	}
	free(buf);
	qsort(a, n, sizeof(mzo_t), mzo_cmp); /* positions end up in read-id, in-read order (htab.cpp:646-672) */
	memset(cnt, 0, sizeof(cnt));
	for (i = 0; i < n; i = j) {
		for (j = i + 1; j < n && a[j].x == a[i].x; j++) {}
		cnt[j - i > 4095 ? 4095 : j - i]++;
	}
	peak_hom = hao_analyze_count(4096, o->min_hist_kmer_cnt, -1, cnt, &peak_het);
	if (hom_cov) *hom_cov = peak_hom;
	if (het_cov) *het_cov = peak_het;
	if (ft == 0) { /* htab.cpp:1259-1264 */
		hi = (int)(peak_hom * o->high_factor);
		if (hi > 4094) hi = 4094;
	}
	for (i = 0; i < n; i = j) {
		for (j = i + 1; j < n && a[j].x == a[i].x; j++) {}
		if ((int64_t)(j - i) >= lo && (int64_t)(j - i) <= hi) nk++, np += j - i;
	}
	pt->key = MALLOC_N(uint64_t, nk); pt->off = MALLOC_N(uint64_t, nk + 1); pt->cnt = MALLOC_N(uint32_t, nk);
	pt->pos = MALLOC_N(uint64_t, np);
	nk = np = 0;
	for (i = 0; i < n; i = j) {
		for (j = i + 1; j < n && a[j].x == a[i].x; j++) {}
		if ((int64_t)(j - i) >= lo && (int64_t)(j - i) <= hi) {
			uint64_t t;
			pt->key[nk] = a[i].x; pt->off[nk] = np; pt->cnt[nk] = (uint32_t)(j - i);
			for (t = i; t < j; t++) pt->pos[np++] = a[t].info; /* ha_idxpos_t = the info word */
			nk++;
		}
	}
	pt->off[nk] = np; pt->n_keys = nk; pt->tot_pos = np;
	free(a);
	return pt;
}

const uint64_t *hao_pt_get(const hao_pt_t *pt, uint64_t hash, int *n)
{ /* ha_pt_get, htab.cpp:518-527 */
	uint64_t lo = 0, hi = pt->n_keys;
	*n = 0;
	while (lo < hi) {
		uint64_t mid = (lo + hi) >> 1;
		if (pt->key[mid] < hash) lo = mid + 1; else hi = mid;
	}
	if (lo < pt->n_keys && pt->key[lo] == hash) { *n = (int)pt->cnt[lo]; return pt->pos + pt->off[lo]; }
	return 0;
}
uint64_t hao_pt_tot_pos(const hao_pt_t *pt) { return pt->tot_pos; }
uint64_t hao_pt_n_keys(const hao_pt_t *pt) { return pt->n_keys; }
void hao_pt_destroy(hao_pt_t *pt) { if (pt) { free(pt->key); free(pt->off); free(pt->cnt); free(pt->pos); free(pt); } }

/* ------------------------------------------------------------------ */
/* anchors: minimizers_qgen0 (anchor.cpp:987-1081)                      */
/* ------------------------------------------------------------------ */

typedef struct { uint64_t srt; uint32_t self_off, other_off, cnt; } anc_t;
static int anc_cmp(const void *a, const void *b)
{ /* radix_sort_ha_an1 on srt, then radix_sort_ha_an3 on other_off inside equal
     srt runs (anchor.cpp:1046-1049): a total order up to identical records */
	const anc_t *p = (const anc_t *)a, *q = (const anc_t *)b;
	if (p->srt != q->srt) return p->srt < q->srt ? -1 : 1;
	return (p->other_off > q->other_off) - (p->other_off < q->other_off);
}

int hao_anchors(const hao_reads_t *r, const hao_pt_t *pt, const hao_mz_t *mz, uint32_t n_mz,
                uint32_t high_occ, uint32_t low_occ, hao_hit_t **out, uint64_t *n_out)
{
	uint64_t i, k, na = 0, max_cnt = high_occ, min_cnt = low_occ; int n, j;
	anc_t *a; hao_hit_t *h;
	if (max_cnt < 2) max_cnt = 2; /* anchor.cpp:992-999 */
	if (min_cnt < 2) min_cnt = 2;
	for (i = 0; i < n_mz; i++) { hao_pt_get(pt, mz[i].x, &n); na += n; }
	a = MALLOC_N(anc_t, na);
	for (i = k = 0; i < n_mz; i++) { /* anchor.cpp:1023-1037 */
		const uint64_t *y = hao_pt_get(pt, mz[i].x, &n);
		uint32_t zrev = HAO_MZ_REV(mz[i]), zpos = HAO_MZ_POS(mz[i]), zspan = HAO_MZ_SPAN(mz[i]);
		for (j = 0; j < n; j++) {
			hao_mz_t t; anc_t *an = &a[k++]; uint32_t rev;
			t.x = 0; t.info = y[j];
			rev = zrev == HAO_MZ_REV(t) ? 0 : 1;
			an->other_off = rev ? ((uint32_t)-1) - 1 - (HAO_MZ_POS(t) + 1 - HAO_MZ_SPAN(t)) : HAO_MZ_POS(t);
			an->self_off = zpos;
			an->cnt = (uint32_t)n; if (an->cnt > 0xffffffu) an->cnt = 0xffffffu;
			an->cnt <<= 8; an->cnt |= zspan <= 0xffu ? zspan : 0xffu;
			an->srt = (uint64_t)HAO_MZ_RID(t) << 33 | (uint64_t)rev << 32 | an->self_off;
		}
	}
	qsort(a, na, sizeof(anc_t), anc_cmp);
	h = MALLOC_N(hao_hit_t, na);
	for (i = 0; i < na; i++) { /* anchor.cpp:1055-1076 */
		uint32_t tid = (uint32_t)(a[i].srt >> 33), strand = (uint32_t)(a[i].srt >> 32) & 1, off, w, nocc = a[i].cnt >> 8;
		uint64_t tl = r->len[tid];
		if (!strand) off = a[i].other_off;
		else { off = ((uint32_t)-1) - a[i].other_off; off = (uint32_t)(tl - off); }
		if (nocc < max_cnt && nocc > min_cnt) w = 1;
		else if (nocc <= min_cnt) w = 2;
		else { w = (uint32_t)(1 + ((nocc + (max_cnt << 1) - 1) / (max_cnt << 1))); w = (uint32_t)pow((double)w, 1.1); }
		if (w > 0xffffffu) w = 0xffffffu;
		h[i].id_strand = tid | strand << 31; h[i].offset = off; h[i].self_offset = a[i].self_off;
		h[i].cnt = w << 8 | (a[i].cnt & 0xffu);
	}
	free(a);
	*out = h; *n_out = na;
	return 0;
}

/* ------------------------------------------------------------------ */
/* klib sorts whose tie order leaks (ksort.h:80-221)                   */
/* ------------------------------------------------------------------ */

typedef int (*ov_lt_f)(const hao_ovlp_t *, const hao_ovlp_t *);
#define OV_SWAP(a, b) do { hao_ovlp_t t_ = (a); (a) = (b); (b) = t_; } while (0)

static void ov_insertsort(hao_ovlp_t *s, hao_ovlp_t *t, ov_lt_f lt)
{ /* __ks_insertsort, ksort.h:80-87 */
	hao_ovlp_t *i, *j;
	for (i = s + 1; i < t; ++i)
		for (j = i; j > s && lt(j, j - 1); --j) OV_SWAP(*j, *(j - 1));
}
static void ov_combsort(size_t n, hao_ovlp_t *a, ov_lt_f lt)
{ /* ks_combsort, ksort.h:88-109 */
	const double shrink = 1.2473309501039786540366528676643;
	int do_swap; size_t gap = n; hao_ovlp_t *i, *j;
	do {
		if (gap > 2) { gap = (size_t)(gap / shrink); if (gap == 9 || gap == 10) gap = 11; }
		do_swap = 0;
		for (i = a; i < a + n - gap; ++i) { j = i + gap; if (lt(j, i)) { OV_SWAP(*i, *j); do_swap = 1; } }
	} while (do_swap || gap > 2);
	if (gap != 1) ov_insertsort(a, a + n, lt);
}
static void ov_introsort(size_t n, hao_ovlp_t *a, ov_lt_f lt)
{ /* ks_introsort, ksort.h:110-160 */
	int d; struct { hao_ovlp_t *l, *r; int d; } stack[128], *top = stack;
	hao_ovlp_t rp, *s, *t, *i, *j, *k;
	if (n < 1) return;
	if (n == 2) { if (lt(&a[1], &a[0])) OV_SWAP(a[0], a[1]); return; }
	for (d = 2; 1ul << d < n; ++d) {}
	s = a; t = a + (n - 1); d <<= 1;
	while (1) {
		if (s < t) {
			if (--d == 0) { ov_combsort(t - s + 1, s, lt); t = s; continue; }
			i = s; j = t; k = i + ((j - i) >> 1) + 1;
			if (lt(k, i)) { if (lt(k, j)) k = j; }
			else k = lt(j, i) ? i : j;
			rp = *k;
			if (k != t) OV_SWAP(*k, *t);
			for (;;) {
				do ++i; while (lt(i, &rp));
				do --j; while (i <= j && lt(&rp, j));
				if (j <= i) break;
				OV_SWAP(*i, *j);
			}
			OV_SWAP(*i, *t);
			if (i - s > t - i) {
				if (i - s > 16) { top->l = s; top->r = i - 1; top->d = d; ++top; }
				s = t - i > 16 ? i + 1 : t;
			} else {
				if (t - i > 16) { top->l = i + 1; top->r = t; top->d = d; ++top; }
				t = i - s > 16 ? i - 1 : s;
			}
		} else {
			if (top == stack) { ov_insertsort(a, a + n, lt); return; }
			--top; s = top->l; t = top->r; d = top->d;
		}
	}
}
static int ov_ss_lt(const hao_ovlp_t *a, const hao_ovlp_t *b) { return a->shared_seed > b->shared_seed; } /* anchor.cpp:35 */
static int ov_xs_lt(const hao_ovlp_t *a, const hao_ovlp_t *b)
{ return ((uint64_t)a->x_pos_s << 32 | a->x_pos_e) < ((uint64_t)b->x_pos_s << 32 | b->x_pos_e); } /* anchor.cpp:32 */

static void ov_rs_insertsort(hao_ovlp_t *beg, hao_ovlp_t *end)
{ /* rs_insertsort, ksort.h:176-186, key = y_id */
	hao_ovlp_t *i;
	for (i = beg + 1; i < end; ++i)
		if (i->y_id < (i - 1)->y_id) {
			hao_ovlp_t *j, tmp = *i;
			for (j = i; j > beg && tmp.y_id < (j - 1)->y_id; --j) *j = *(j - 1);
			*j = tmp;
		}
}
static void ov_rs_sort(hao_ovlp_t *beg, hao_ovlp_t *end, int n_bits, int s)
{ /* rs_sort, ksort.h:187-216: in-place MSD radix, unstable */
	hao_ovlp_t *i; int size = 1 << n_bits, m = size - 1;
	struct { hao_ovlp_t *b, *e; } *k, b[256], *be = b + size;
	for (k = b; k != be; ++k) k->b = k->e = beg;
	for (i = beg; i != end; ++i) ++b[i->y_id >> s & m].e;
	for (k = b + 1; k != be; ++k) k->e += (k - 1)->e - beg, k->b = (k - 1)->e;
	for (k = b; k != be;) {
		if (k->b != k->e) {
			__typeof__(k) l;
			if ((l = b + (k->b->y_id >> s & m)) != k) {
				hao_ovlp_t tmp = *k->b, swap;
				do { swap = tmp; tmp = *l->b; *l->b++ = swap; l = b + (tmp.y_id >> s & m); } while (l != k);
				*k->b++ = tmp;
			} else ++k->b;
		} else ++k;
	}
	for (b->b = beg, k = b + 1; k != be; ++k) k->b = (k - 1)->e;
	if (s) {
		s = s > n_bits ? s - n_bits : 0;
		for (k = b; k != be; ++k)
			if (k->e - k->b > 64) ov_rs_sort(k->b, k->e, n_bits, s);
			else if (k->e - k->b > 1) ov_rs_insertsort(k->b, k->e);
	}
}
static void ov_sort_y_id(hao_ovlp_t *a, size_t n)
{ /* overlap_region_sort_y_id -> radix_sort_overlap_region_sort (Hash_Table.cpp:12,22; ksort.h:217-221) */
	if (n <= 64) ov_rs_insertsort(a, a + n);
	else ov_rs_sort(a, a + n, 8, 4 * 8 - 8);
}

static int i64_cmp(const void *a, const void *b) { int64_t x = *(const int64_t *)a, y = *(const int64_t *)b; return (x > y) - (x < y); }
static int u64_cmp(const void *a, const void *b) { uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return (x > y) - (x < y); }

/* ------------------------------------------------------------------ */
/* chaining                                                            */
/* ------------------------------------------------------------------ */

#define HIT_ID(h) ((h).id_strand & 0x7fffffffu)
#define HIT_ST(h) ((h).id_strand >> 31)

static int64_t chain_len(int64_t xb, int64_t xe, int64_t xl, int64_t yb, int64_t ye, int64_t yl)
{ /* get_chainLen, Hash_Table.cpp:779-809 */
	int64_t xr, yr;
	if (xb <= yb) { yb -= xb; xb = 0; } else { xb -= yb; yb = 0; }
	xr = xl - xe - 1; yr = yl - ye - 1;
	if (xr <= yr) { xe = xl - 1; ye += xr; } else { xe += yr; ye = yl - 1; }
	(void)yb; (void)ye;
	return xe - xb + 1;
}

static int32_t link_bw(const hao_hit_t *ai, const hao_hit_t *aj, double bw_rate, int64_t sf_l, int64_t ot_l)
{ /* cal_bw, Hash_Table.cpp:1475-1488 */
	int64_t sf_s = aj->self_offset, sf_e = (int64_t)ai->self_offset + 1;
	int64_t ot_s = aj->offset, ot_e = (int64_t)ai->offset + 1;
	int64_t sf_r = sf_l - sf_e, ot_r = ot_l - ot_e;
	if (sf_s <= ot_s) sf_s = 0; else sf_s -= ot_s;
	if (sf_r <= ot_r) sf_e = sf_l; else sf_e += ot_r;
	return (int32_t)((sf_e - sf_s) * bw_rate);
}

#define LINK_FAIL INT32_MIN
static int32_t link_sc(const hao_hit_t *ai, const hao_hit_t *aj, double bw_rate, double pen_gap, double pen_skip, int64_t sl, int64_t ol, int64_t *pdd)
{ /* comput_sc_ch_ec, Hash_Table.cpp:1515-1541 */
	int32_t dq, dr, dd, dg, q_span, sc, w;
	dq = (int32_t)((int64_t)ai->self_offset - (int64_t)aj->self_offset);
	if (dq <= 0) return LINK_FAIL;
	dr = (int32_t)((int64_t)ai->offset - (int64_t)aj->offset);
	if (dr <= 0) return LINK_FAIL;
	dd = dr > dq ? dr - dq : dq - dr;
	if (dd > 16 && dd > link_bw(ai, aj, bw_rate, sl, ol)) return LINK_FAIL;
	dg = dr < dq ? dr : dq;
	q_span = ai->cnt & 0xffu;
	sc = q_span < dg ? q_span : dg;
	w = (int32_t)(ai->cnt >> 8);
	sc = sc >= w ? sc / w : 1; /* normal_w, Hash_Table.cpp:20 */
	if (dd || (dg > q_span && dg > 0)) {
		double lin_pen = pen_gap * (double)dd, a_pen = ((double)sc) * ((((double)dd) / ((double)dg)) / bw_rate);
		if (dd < 4) lin_pen = lin_pen > a_pen ? a_pen : lin_pen;
		else lin_pen = lin_pen < a_pen ? a_pen : lin_pen;
		lin_pen += pen_skip * (double)dg;
		sc -= (int32_t)lin_pen;
	}
	if (pdd) *pdd = dd;
	return sc;
}

typedef struct { hao_ovlp_t *a; uint32_t n, m; uint64_t *fc; uint64_t n_fc, m_fc; } ovv_t;
static hao_ovlp_t *ovv_push(ovv_t *v)
{
	if (v->n == v->m) { v->m = v->m ? v->m << 1 : 64; v->a = (hao_ovlp_t *)realloc(v->a, v->m * sizeof(hao_ovlp_t)); }
	memset(&v->a[v->n], 0, sizeof(hao_ovlp_t));
	return &v->a[v->n++];
}
static void fc_add(ovv_t *v, uint32_t site, int32_t shift)
{ /* add_fake_cigar, Hash_Table.cpp:1295-1328 */
	uint32_t e;
	if (v->n_fc == v->m_fc) { v->m_fc = v->m_fc ? v->m_fc << 1 : 256; v->fc = (uint64_t *)realloc(v->fc, v->m_fc * 8); }
	e = shift < 0 ? ((uint32_t)(-shift) << 1 | 1u) : (uint32_t)shift << 1;
	v->fc[v->n_fc++] = (uint64_t)site << 32 | e;
}
static int fc_shift(uint64_t e)
{ /* get_fake_gap_shift, Hash_Table.cpp:69-86 */
	uint32_t t = (uint32_t)e; int r = (int)(t >> 1);
	return (t & 1) ? -r : r;
}
static void gen_fcigar(ovv_t *v, hao_ovlp_t *o, const hao_hit_t *hit, int64_t n_hit)
{ /* gen_fake_cigar with apend_be = 1, Hash_Table.cpp:88-109 */
	int64_t k, dq, dr, dd, pdd = INT32_MAX;
	o->fc_off = (uint32_t)v->n_fc;
	fc_add(v, o->x_pos_s, 0);
	for (k = 0; k < n_hit; k++) {
		dq = (uint32_t)(hit[k].self_offset - o->x_pos_s);
		dr = (uint32_t)(hit[k].offset - o->y_pos_s);
		dd = dr - dq;
		if (dd != pdd) { pdd = dd; fc_add(v, hit[k].self_offset, (int32_t)pdd); }
	}
	if ((int64_t)(v->fc[v->n_fc - 1] >> 32) != (int64_t)o->x_pos_e)
		fc_add(v, o->x_pos_e, fc_shift(v->fc[v->n_fc - 1]));
	o->fc_n = (uint32_t)v->n_fc - o->fc_off;
}

static void push_chain(hao_ovlp_t *o, int64_t xl, int64_t yl, int64_t sc, const hao_hit_t *beg, const hao_hit_t *end)
{ /* push_ovlp_chain_qgen, Hash_Table.cpp:1752-1780 */
	int64_t xr, yr;
	o->y_id = HIT_ID(*beg); o->y_pos_strand = HIT_ST(*beg);
	o->x_pos_s = beg->self_offset; o->y_pos_s = beg->offset;
	o->x_pos_e = end->self_offset; o->y_pos_e = end->offset;
	if (o->x_pos_s <= o->y_pos_s) { o->y_pos_s -= o->x_pos_s; o->x_pos_s = 0; }
	else { o->x_pos_s -= o->y_pos_s; o->y_pos_s = 0; }
	xr = xl - o->x_pos_e - 1; yr = yl - o->y_pos_e - 1;
	if (xr <= yr) { o->x_pos_e = (uint32_t)(xl - 1); o->y_pos_e += (uint32_t)xr; }
	else { o->y_pos_e = (uint32_t)(yl - 1); o->x_pos_e += (uint32_t)yr; }
	o->shared_seed = (int32_t)sc; o->align_length = 0; o->is_match = 0; o->non_homopolymer_errors = 0; o->strong = 0; o->overlapLen = 0;
}

typedef struct { int32_t *f, *ii; int64_t *p, *t; int64_t m; hao_hit_t *swap; int64_t m_swap; } dp_t;
static void dp_reserve(dp_t *d, int64_t n)
{
	if (n > d->m) {
		d->m = n + (n >> 1) + 16;
		d->f = (int32_t *)realloc(d->f, d->m * 4); d->ii = (int32_t *)realloc(d->ii, d->m * 4);
		d->p = (int64_t *)realloc(d->p, d->m * 8); d->t = (int64_t *)realloc(d->t, d->m * 8);
	}
}

static void quick_check(const hao_hit_t *a, int64_t a_n, int64_t xl, int64_t yl, double pen_gap, double pen_skip, double bw_rate,
                        dp_t *d, int64_t *plus, int64_t *msc, int64_t *msc_i, int64_t *movl, int64_t *si, int64_t *ei)
{ /* quick_ck_lchain, Hash_Table.cpp:2007-2094: a strand block whose anchors are
     already co-linear is resolved by a linear scan */
	int64_t l, k, z, sorted = 1; int64_t *p = d->p, *t = d->t; int32_t *f = d->f, *ii = d->ii;
	*plus = 0; *msc = *msc_i = INT32_MIN; *movl = INT32_MAX; *si = 0; *ei = a_n;
	for (k = 1, l = 0; k <= a_n; k++) {
		if (k == a_n || HIT_ST(a[k]) != HIT_ST(a[l])) {
			t[k - 1] = 0; ii[k - 1] = 0;
			if (sorted) {
				int64_t plus0 = 0, msc0 = INT32_MIN, msc_i0 = INT32_MIN, movl0, ddt = 0;
				p[l] = -1; f[l] = a[l].cnt & 0xffu;
				if (f[l] >= msc0) { msc0 = f[l]; msc_i0 = l; }
				if (f[l] < plus0) plus0 = f[l];
				for (z = l + 1; z < k; z++) {
					int64_t dd = 0, sc, csc; int32_t s = link_sc(&a[z], &a[z - 1], bw_rate, pen_gap, pen_skip, xl, yl, &dd);
					if (s == LINK_FAIL) break;
					sc = (int64_t)s + f[z - 1]; csc = a[z].cnt & 0xffu;
					if (sc < csc) break;
					p[z] = z - 1; f[z] = (int32_t)sc; ddt += dd;
					if (f[z] >= msc0) { msc0 = f[z]; msc_i0 = z; }
					if (f[z] < plus0) plus0 = f[z];
				}
				if (z >= k && msc_i0 == k - 1) {
					if (k - l >= 2 && ddt > 16 && ddt > link_bw(&a[k - 1], &a[l], bw_rate, xl, yl)) msc_i0 = INT32_MIN;
					if (msc_i0 == k - 1) {
						if (msc0 >= *msc) {
							movl0 = chain_len(a[msc_i0].self_offset, a[msc_i0].self_offset, xl, a[msc_i0].offset, a[msc_i0].offset, yl);
							if (msc0 > *msc || movl0 < *movl) { *msc = msc0; *msc_i = msc_i0; *movl = movl0; }
						}
						if (plus0 < *plus) *plus = plus0;
						if (*ei > k) *si = k; else *ei = l;
					}
				}
			}
			l = k; sorted = 1;
		} else {
			if (a[k].self_offset <= a[k - 1].self_offset || a[k].offset <= a[k - 1].offset) sorted = 0;
			t[k - 1] = 0; ii[k - 1] = 0;
		}
	}
}

static int64_t chain_group(hao_hit_t *list, int64_t a_idx, int64_t a_n, int64_t des_idx, dp_t *d, ovv_t *res,
                           int64_t max_skip, int64_t max_iter, int64_t max_dis, double pen_gap, double pen_skip, double bw_rate,
                           int64_t xl, int64_t yl, int64_t mcopy_num, double mcopy_rate, int64_t mcopy_khit_cutoff)
{ /* lchain_qdp_mcopy_fast, Hash_Table.cpp:2097-2284 (quick_check=1, apend_be=1, gen_cigar=1, khit_n=1) */
	int64_t *p, *t, max_f, n_skip, st, max_j, end_j, sc, msc, msc_i, max_ii, ovl, movl, plus = 0, min_sc, ch_n, si, ei;
	int32_t *f, max, tmp, *ii; int64_t i, k, j, cL = 0; hao_hit_t *a, *des; hao_ovlp_t *z;
	if (a_n <= 0) return 0;
	dp_reserve(d, a_n);
	t = d->t; f = d->f; p = d->p; ii = d->ii;
	a = list + a_idx; des = list + des_idx;
	quick_check(a, a_n, xl, yl, pen_gap, pen_skip, bw_rate, d, &plus, &msc, &msc_i, &movl, &si, &ei);
	for (i = st = si, max_ii = -1; i < ei; ++i) { /* 2124-2176 */
		max_f = a[i].cnt & 0xffu;
		n_skip = 0; max_j = end_j = -1;
		if (i - st > max_iter) st = i - max_iter;
		while (HIT_ST(a[i]) != HIT_ST(a[st])) ++st;
		for (j = i - 1; j >= st; --j) {
			int32_t s = link_sc(&a[i], &a[j], bw_rate, pen_gap, pen_skip, xl, yl, 0);
			if (s == LINK_FAIL) continue;
			sc = (int64_t)s + f[j];
			if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
			else if (t[j] == (int32_t)i) { if (++n_skip > max_skip) break; }
			if (p[j] >= 0) t[p[j]] = i;
		}
		end_j = j;
		if (max_ii < 0 || a[i].self_offset > a[max_ii].self_offset + max_dis || HIT_ST(a[i]) != HIT_ST(a[max_ii])) {
			max = INT32_MIN; max_ii = -1;
			for (j = i - 1; j >= st && (int64_t)a[i].self_offset <= max_dis + (int64_t)a[j].self_offset && HIT_ST(a[i]) == HIT_ST(a[j]); --j)
				if (max < f[j]) max = f[j], max_ii = j;
		}
		if (max_ii >= 0 && max_ii < end_j && HIT_ST(a[i]) == HIT_ST(a[max_ii])) {
			tmp = link_sc(&a[i], &a[max_ii], bw_rate, pen_gap, pen_skip, xl, yl, 0);
			if (tmp != LINK_FAIL && max_f < (int64_t)tmp + f[max_ii]) max_f = (int64_t)tmp + f[max_ii], max_j = max_ii;
		}
		f[i] = (int32_t)max_f; p[i] = max_j;
		if (max_ii < 0 || ((int64_t)a[i].self_offset <= max_dis + (int64_t)a[max_ii].self_offset && HIT_ST(a[i]) == HIT_ST(a[max_ii]) && f[max_ii] < f[i])) max_ii = i;
		if (f[i] >= msc) {
			ovl = chain_len(a[i].self_offset, a[i].self_offset, xl, a[i].offset, a[i].offset, yl);
			if (f[i] > msc || ovl < movl) { msc = f[i]; msc_i = i; movl = ovl; }
		}
		if (f[i] < plus) plus = f[i];
		ii[i] = 0;
	}
	for (i = msc_i, cL = 0; i >= 0; i = p[i]) { ii[i] = 1; t[cL++] = i; } /* 2178 */

	if (mcopy_num > 1 && cL >= mcopy_khit_cutoff) { /* 2180-2270 */
		msc -= plus; min_sc = (int64_t)(msc * mcopy_rate); ii[msc_i] = 0;
		for (i = ch_n = 0; i < a_n; ++i) {
			f[i] -= (int32_t)plus; if (i >= ch_n) t[i] = 0;
			if (!ii[i] && f[i] >= min_sc) { t[ch_n] = (int64_t)(((uint64_t)f[i]) << 32); t[ch_n] += (i << 1); ch_n++; }
		}
		if (ch_n > 1) {
			int64_t n_v, n_v0, ni, n_u, n_u0 = res->n;
			qsort(t, ch_n, sizeof(int64_t), i64_cmp); /* radix_sort_hc64i: keys are distinct */
			for (k = ch_n - 1, n_v = n_u = 0; k >= 0 && n_u < mcopy_num; --k) {
				n_v0 = n_v;
				for (i = ((uint32_t)t[k]) >> 1; i >= 0 && (t[i] & 1) == 0;) { ii[n_v++] = (int32_t)i; t[i] |= 1; i = p[i]; }
				if (n_v0 == n_v) continue;
				sc = i < 0 ? (t[k] >> 32) : ((t[k] >> 32) - f[i]);
				if (sc >= min_sc) {
					z = ovv_push(res);
					push_chain(z, xl, yl, sc + plus, &a[ii[n_v - 1]], &a[ii[n_v0]]);
					if (!n_u || n_v - n_v0 > 1) { z->align_length = (uint32_t)(n_v - n_v0); z->overlapLen = (uint32_t)n_v0; /* x_id used as scratch at 2217 */ n_u++; }
					else { res->n--; n_v = n_v0; }
				} else n_v = n_v0;
			}
			n_u = res->n;
			if (n_v + 1 > d->m_swap) { d->m_swap = n_v + 64; d->swap = (hao_hit_t *)realloc(d->swap, d->m_swap * sizeof(hao_hit_t)); }
			for (k = n_u0, i = 0; k < n_u; k++) { /* 2230-2263, both branches */
				z = &res->a[k];
				z->non_homopolymer_errors = (uint32_t)(des_idx + i);
				n_v0 = z->overlapLen; ni = z->align_length; z->overlapLen = 0;
				for (j = 0; j < ni; j++, i++) {
					d->swap[i] = a[ii[n_v0 + (ni - j - 1)]];
					d->swap[i].id_strand = (d->swap[i].id_strand & 0x80000000u) | ((uint32_t)k & 0x7fffffffu);
				}
				gen_fcigar(res, z, d->swap + i - ni, ni);
			}
			memcpy(des, d->swap, i * sizeof(hao_hit_t));
			return i;
		}
		msc += plus; i = msc_i; cL = 0;
		while (i >= 0) { t[cL++] = i; i = p[i]; }
	}
	z = ovv_push(res); /* 2277-2283 */
	push_chain(z, xl, yl, msc, &a[t[cL - 1]], &a[t[0]]);
	for (i = 0; i < cL; i++) {
		des[i] = a[t[cL - i - 1]];
		des[i].id_strand = (des[i].id_strand & 0x80000000u) | ((uint32_t)(res->n - 1) & 0x7fffffffu);
	}
	z->non_homopolymer_errors = (uint32_t)des_idx;
	gen_fcigar(res, z, des, cL);
	z->align_length = (uint32_t)cL;
	return cL;
}

static int ov_type(const hao_ovlp_t *r, uint64_t len)
{ /* ha_ov_type, anchor.cpp:86-91 */
	if (r->x_pos_s == 0 && r->x_pos_e == len - 1) return 2;
	if (r->x_pos_s > 0 && r->x_pos_e < len - 1) return 3;
	return r->x_pos_s == 0 ? 0 : 1;
}

static void cov_add(uint64_t *cc, uint64_t cwn, uint64_t ocv_w, uint64_t rl, const hao_ovlp_t *o)
{ /* the coverage-window accumulation block, anchor.cpp:1986-1999 and 2028-2041 */
	uint64_t m = o->x_pos_s / ocv_w, rs = o->x_pos_s, re = (uint64_t)o->x_pos_e + 1, cws, cwe, os, oe;
	for (cws = m * ocv_w; m < cwn; m++) {
		cwe = cws + ocv_w; if (cwe > rl) cwe = rl;
		os = rs >= cws ? rs : cws; oe = re <= cwe ? re : cwe;
		if (oe <= os) break;
		if (((uint32_t)cc[m]) + (oe - os) < UINT32_MAX) cc[m] += oe - os;
		else { cc[m] >>= 32; cc[m] <<= 32; cc[m] |= UINT32_MAX; }
		cws += ocv_w;
	}
}

#define OFL 0.95   /* anchor.cpp:12-14 */
#define CH_OCC 4
#define CH_SC 16

int hao_lchain(const hao_reads_t *r, uint32_t rid, hao_hit_t *hits, uint64_t *n_hits, double bw_rate, int mz_k,
               int max_n_chain_, hao_ovlp_t **out, uint32_t *n_out, uint64_t **fc, uint64_t *n_fc)
{ /* set_lchain_dp_op(is_accurate=1) + lchain_qgen_mcopy_fast, anchor.cpp:2272-2285, 1920-2100
     with the arguments of ecovlp.cpp:3957 (apend_be=1, gen_off=1, mcopy_num=3, mcopy_rate=0.7,
     chain_cutoff=2, mcopy_khit_cut=32, ocv_w=3072) */
	const int64_t max_skip = 25, max_iter = 5000, max_dis = 5000; const uint32_t chain_cutoff = 2; const uint64_t ocv_w = 3072;
	double pen_gap = 0.5f, pen_skip = 0.0005f, tmp; uint64_t max_n_chain = (uint64_t)max_n_chain_;
	uint64_t i, k, l, m, cn = *n_hits, lch = 0, rl = r->len[rid], cwn = 0, *cc = 0; ovv_t ol; dp_t d; hao_ovlp_t t;
	memset(&ol, 0, sizeof(ol)); memset(&d, 0, sizeof(d));
	tmp = expf((float)(-0.01 * (double)mz_k));
	pen_gap *= tmp; pen_skip *= tmp;
	for (l = 0, k = 1, m = 0; k <= cn; k++) { /* 1929-1943 */
		if (k == cn || HIT_ID(hits[k]) != HIT_ID(hits[l])) {
			if (HIT_ID(hits[l]) != rid) {
				uint32_t yid = HIT_ID(hits[l]), ol0 = ol.n;
				m += chain_group(hits, l, k - l, m, &d, &ol, max_skip, max_iter, max_dis, pen_gap, pen_skip, bw_rate, rl, r->len[yid], 3, 0.7, 32);
				if (chain_cutoff >= 2 && !lch)
					for (i = ol0; i < ol.n && !lch; i++)
						if (ol.a[i].align_length < chain_cutoff) lch = 1;
			}
			l = k;
		}
	}
	*n_hits = m;
	k = ol.n;
	if (ol.n > max_n_chain) { /* 1954-2058 */
		int32_t w, n[4] = { 0, 0, 0, 0 }, s[4] = { 0, 0, 0, 0 };
		ov_introsort(ol.n, ol.a, ov_ss_lt);
		for (i = 0; i < ol.n; ++i) {
			w = ov_type(&ol.a[i], rl); ++n[w];
			if ((uint64_t)n[w] == max_n_chain) s[w] = ol.a[i].shared_seed;
		}
		if (s[0] > 0 || s[1] > 0 || s[2] > 0 || s[3] > 0) {
			if ((uint64_t)n[3] >= max_n_chain && rl >= ocv_w) {
				uint64_t cws, cwe;
				cwn = rl / ocv_w + (rl % ocv_w ? 1 : 0);
				cc = MALLOC_N(uint64_t, cwn);
				for (i = cws = 0; i < cwn; i++) {
					cwe = cws + ocv_w; if (cwe > rl) cwe = rl;
					cc[i] = (cwe - cws) * (max_n_chain >> 1);
					if (cc[i] > UINT32_MAX) cc[i] = UINT32_MAX;
					cc[i] <<= 32;
					cws += ocv_w;
				}
			}
			for (i = 0, k = 0, lch = 0; i < ol.n; ++i) {
				hao_ovlp_t *o = &ol.a[i];
				w = ov_type(o, rl);
				if (o->shared_seed >= s[w]) {
					if (cwn) cov_add(cc, cwn, ocv_w, rl, o);
					if (k != i) { t = ol.a[k]; ol.a[k] = ol.a[i]; ol.a[i] = t; }
					if (ol.a[k].align_length < chain_cutoff) lch = 1;
					++k;
				} else if (w == 3 && cwn > 0) {
					uint64_t mm = o->x_pos_s / ocv_w, cw0 = 0, cw1 = 0, rs = o->x_pos_s, re = (uint64_t)o->x_pos_e + 1, cws, cwe, os, oe;
					for (cws = mm * ocv_w; mm < cwn; mm++) {
						cwe = cws + ocv_w; if (cwe > rl) cwe = rl;
						os = rs >= cws ? rs : cws; oe = re <= cwe ? re : cwe;
						if (oe <= os) break;
						if ((oe - os) + ((uint64_t)((uint32_t)cc[mm])) >= (cc[mm] >> 32)) cw1 += oe - os; else cw0 += oe - os;
						cws += ocv_w;
					}
					if (cw0 >= ((cw0 + cw1) * 0.7)) {
						cov_add(cc, cwn, ocv_w, rl, o);
						if (k != i) { t = ol.a[k]; ol.a[k] = ol.a[i]; ol.a[i] = t; }
						if (ol.a[k].align_length < chain_cutoff) lch = 1;
						++k;
					}
				}
			}
			ol.n = (uint32_t)k;
		}
		free(cc);
	}
	ov_introsort(ol.n, ol.a, ov_xs_lt); /* 2060 */
	if (lch) { /* 2061-2096: drop tiny chains shadowed by a big one */
		uint64_t zs, ze, rs, re, ob, os, oe, ocn, pp, kn, ms, me, mm; int64_t osc;
		for (i = l = 0, cn = *n_hits; i < ol.n; ++i) {
			if (ol.a[i].align_length < chain_cutoff) {
				zs = ol.a[i].x_pos_s; ze = (uint64_t)ol.a[i].x_pos_e + 1;
				ob = (uint64_t)((ze - zs) * OFL); if (ob < 16) ob = 16;
				osc = (int64_t)ol.a[i].shared_seed * CH_SC;
				ocn = (uint64_t)ol.a[i].align_length << CH_OCC;
				for (k = 0; k < ol.n && ze > ol.a[k].x_pos_s; k++) {
					if (ol.a[k].align_length < chain_cutoff) continue;
					if (ol.a[k].align_length < ocn) continue;
					if (ol.a[k].shared_seed < osc) continue;
					rs = ol.a[k].x_pos_s; re = (uint64_t)ol.a[k].x_pos_e + 1;
					os = rs >= zs ? rs : zs; oe = re <= ze ? re : ze;
					if (oe > os && oe - os >= ob) {
						mm = ol.a[k].non_homopolymer_errors;
						pp = HIT_ID(hits[mm]); kn = 0;
						for (; mm < cn && HIT_ID(hits[mm]) == pp && kn < ocn; mm++) {
							me = hits[mm].self_offset; ms = me - (hits[mm].cnt & 0xffu);
							if (ms >= os && me <= oe) kn++;
						}
						if (kn >= ocn) break;
					}
				}
				if (k < ol.n && ze > ol.a[k].x_pos_s) continue;
			}
			if (l != i) { t = ol.a[l]; ol.a[l] = ol.a[i]; ol.a[i] = t; }
			l++;
		}
		ol.n = (uint32_t)l;
	}
	for (i = 0; i < ol.n; ++i) ol.a[i].align_length = 0; /* 2098 */
	free(d.f); free(d.ii); free(d.p); free(d.t); free(d.swap);
	*out = ol.a; *n_out = ol.n;
	if (fc) { *fc = ol.fc; *n_fc = ol.n_fc; } else free(ol.fc);
	return 0;
}

/* ------------------------------------------------------------------ */
/* final overlap pass for one read                                     */
/* ------------------------------------------------------------------ */

#define HA_KMER_GOOD_RATIO 0.333 /* ecovlp.cpp:9 */

int hao_final_read(const hao_reads_t *r, const hao_pt_t *pt, const hao_ft_t *ft, const hao_opt_t *o, uint32_t rid,
                   hao_ma_t *in0, uint32_t n0, const hao_ma_t *in1, uint32_t n1,
                   hao_ma_t **out0, uint32_t *m0, hao_ma_t **out1, uint32_t *m1)
{ /* worker_hap_dc_ec_gen_new_idx, ecovlp.cpp:3948-3992 */
	uint32_t high_occ = (uint32_t)(o->hom_cov * (2.0 - HA_KMER_GOOD_RATIO)), low_occ = (uint32_t)(o->hom_cov * HA_KMER_GOOD_RATIO);
	uint64_t rl = r->len[rid], n_hits, k, i, l, m, ns = 0, *srt; const double sh = 0.866666;
	char *qs = MALLOC_N(char, rl + 1), *ts = 0; uint64_t ts_m = 0;
	hao_mz_t *mz; uint32_t n_mz, n_ol, is_usrt = 0; hao_hit_t *hits; hao_ovlp_t *ol, t; int pass;

	hao_decode(r, rid, qs);
	hao_sketch(qs, (int)rl, o->w, o->k, 0, o->is_hpc, ft, o->mz_sample_dist, o->mz_rewin, &mz, &n_mz);
	hao_anchors(r, pt, mz, n_mz, high_occ, low_occ, &hits, &n_hits);
	hao_lchain(r, rid, hits, &n_hits, 0.001, o->k, o->max_n_chain, &ol, &n_ol, 0, 0);
	free(mz); free(hits);
	ol = (hao_ovlp_t *)realloc(ol, ((size_t)n_ol + n0 + 1) * sizeof(hao_ovlp_t));
	ov_sort_y_id(ol, n_ol); /* ecovlp.cpp:3959 */

	/* h_ec_lchain_fast_new, ecovlp.cpp:5047-5196 */
	srt = MALLOC_N(uint64_t, (uint64_t)n0 + n1);
	for (k = 0; k < n0; k++) srt[ns++] = ((uint64_t)in0[k].tn << 1 | (uint64_t)in0[k].rev) << 32 | (k << 1) | 0;
	for (k = 0; k < n1; k++) srt[ns++] = ((uint64_t)in1[k].tn << 1 | (uint64_t)in1[k].rev) << 32 | (k << 1) | 1;
	qsort(srt, ns, 8, u64_cmp); /* radix_sort_ec64: keys are distinct */
	for (k = m = 0, i = 0; k < n_ol; k++) {
		hao_ovlp_t *z = &ol[k]; uint64_t tid = z->y_id, trev = z->y_pos_strand, bq[2], bt[2]; int64_t om;
		z->non_homopolymer_errors = 0;
		bq[0] = z->x_pos_s; bq[1] = (uint64_t)z->x_pos_e + 1; bt[0] = z->y_pos_s; bt[1] = (uint64_t)z->y_pos_e + 1;
		for (; i < ns && (srt[i] >> 32) < (tid << 1 | trev); i++) {}
		if (i < ns && (srt[i] >> 32) == (tid << 1 | trev)) {
			hao_ma_t *p; uint32_t is_match; uint64_t aq[2], at[2], os, oe, ovlp;
			om = 1; z->shared_seed = 0; z->is_match = 0;
			if (srt[i] & 1) { p = (hao_ma_t *)&in1[((uint32_t)srt[i]) >> 1]; is_match = 2; }
			else { p = &in0[((uint32_t)srt[i]) >> 1]; is_match = 1; }
			z->strong = (int8_t)p->ml; z->without_large_indel = p->no_l_indel;
			aq[0] = (uint32_t)p->qns; aq[1] = p->qe; at[0] = p->ts; at[1] = p->te;
			os = aq[0] > bq[0] ? aq[0] : bq[0]; oe = aq[1] < bq[1] ? aq[1] : bq[1];
			ovlp = oe > os ? oe - os : 0;
			if (!(ovlp && ovlp >= (aq[1] - aq[0]) * sh && ovlp >= (bq[1] - bq[0]) * sh)) om = 0;
			z->non_homopolymer_errors += (uint32_t)((aq[1] - aq[0]) - ovlp);
			os = at[0] > bt[0] ? at[0] : bt[0]; oe = at[1] < bt[1] ? at[1] : bt[1];
			ovlp = oe > os ? oe - os : 0;
			if (!(ovlp && ovlp >= (at[1] - at[0]) * sh && ovlp >= (bt[1] - bt[0]) * sh)) om = 0;
			z->non_homopolymer_errors += (uint32_t)((at[1] - at[0]) - ovlp);
			if (om) {
				if (is_match == 1 && p->el == 1) p->el = 0;
				if (bt[1] - bt[0] + 1 > ts_m) { ts_m = bt[1] - bt[0] + 1; ts = (char *)realloc(ts, ts_m); }
				hao_decode_sub(r, tid, (int64_t)bt[0], (int64_t)(bt[1] - bt[0]), (int)trev, ts);
				/* exact_ec_check, ecovlp.cpp:2803-2808 */
				if (bq[1] - bq[0] == bt[1] - bt[0] && memcmp(qs + bq[0], ts, bq[1] - bq[0]) == 0) {
					if (is_match == 2) { z->strong = 0; z->without_large_indel = 1; }
					is_match = 1; z->shared_seed = 1;
				}
				z->is_match = (uint8_t)is_match;
			}
		} else {
			om = 0;
			if (bt[1] - bt[0] + 1 > ts_m) { ts_m = bt[1] - bt[0] + 1; ts = (char *)realloc(ts, ts_m); }
			hao_decode_sub(r, tid, (int64_t)bt[0], (int64_t)(bt[1] - bt[0]), (int)trev, ts);
			if (bq[1] - bq[0] == bt[1] - bt[0] && memcmp(qs + bq[0], ts, bq[1] - bq[0]) == 0) {
				z->strong = 0; z->without_large_indel = 1; z->shared_seed = 1; z->is_match = 1; om = 1;
			}
		}
		if (om) { if (m != k) { t = ol[m]; ol[m] = ol[k]; ol[k] = t; } m++; }
	}
	n_ol = (uint32_t)m;
	for (k = 0; k < n0; k++) { /* 5138-5165: previous exact overlaps not found again */
		if (in0[k].el) {
			hao_ovlp_t *z = &ol[n_ol++];
			memset(z, 0, sizeof(*z));
			z->y_id = in0[k].tn; z->y_pos_strand = in0[k].rev;
			z->x_pos_s = (uint32_t)in0[k].qns; z->x_pos_e = in0[k].qe - 1;
			z->y_pos_s = in0[k].ts; z->y_pos_e = in0[k].te - 1;
			z->align_length = z->overlapLen = z->x_pos_e + 1 - z->x_pos_s;
			z->non_homopolymer_errors = 0; z->is_match = 1; z->shared_seed = 1;
			z->strong = (int8_t)in0[k].ml; z->without_large_indel = in0[k].no_l_indel;
			is_usrt = 1;
		}
	}
	if (is_usrt) ov_sort_y_id(ol, n_ol);
	if (n_ol > 1) { /* 5169-5195: best chain per target */
		uint64_t mm_k, s; int64_t mm_sc, sc;
		for (k = 1, l = m = 0; k <= n_ol; k++) {
			if (k == n_ol || ol[k].y_id != ol[l].y_id) {
				mm_k = l;
				if (k - l > 1) {
					for (s = l, mm_sc = INT32_MIN, mm_k = (uint64_t)-1; s < k; s++) {
						hao_ovlp_t *z = &ol[s];
						sc = z->non_homopolymer_errors; sc = -sc;
						if (sc > mm_sc || (sc == mm_sc && (ol[mm_k].x_pos_e + 1 - ol[mm_k].x_pos_s) < (z->x_pos_e + 1 - z->x_pos_s))) { mm_sc = sc; mm_k = s; }
					}
				}
				if (mm_k != (uint64_t)-1) {
					if (mm_k != m) { t = ol[mm_k]; ol[mm_k] = ol[m]; ol[m] = t; }
					m++;
				}
				l = k;
			}
		}
		n_ol = (uint32_t)m;
	}
	/* push_ff_ovlp x2, ecovlp.cpp:2641-2694 */
	*out0 = MALLOC_N(hao_ma_t, n_ol); *out1 = MALLOC_N(hao_ma_t, n_ol); *m0 = *m1 = 0;
	for (pass = 1; pass <= 2; pass++) {
		hao_ma_t *dst = pass == 1 ? *out0 : *out1; uint32_t *dn = pass == 1 ? m0 : m1;
		for (k = 0; k < n_ol; k++) {
			hao_ma_t *z;
			if (ol[k].is_match != pass) continue;
			z = &dst[(*dn)++]; memset(z, 0, sizeof(*z));
			z->qns = (uint64_t)rid << 32 | ol[k].x_pos_s; z->tn = ol[k].y_id;
			z->qe = ol[k].x_pos_e + 1; z->ts = ol[k].y_pos_s; z->te = ol[k].y_pos_e + 1;
			z->rev = ol[k].y_pos_strand; z->bl = (uint32_t)r->len[ol[k].y_id] & 0x7fffffffu;
			z->ml = ol[k].strong & 1; z->no_l_indel = ol[k].without_large_indel; z->el = (uint8_t)ol[k].shared_seed;
			if (z->rev) { z->ts = z->bl - ol[k].y_pos_e - 1; z->te = z->bl - ol[k].y_pos_s; }
			z->del = 0;
		}
	}
	free(ol); free(srt); free(qs); free(ts);
	return 0;
}

/* ------------------------------------------------------------------ */
/* window alignment: banded semi-global Myers in one 64-bit word       */
/* ------------------------------------------------------------------ */

int hao_ed_semi_64_absent_diag(const char *pstr, int32_t pn, const char *tstr, int32_t tn, int32_t thre, int32_t abs_diag, int32_t *pe)
{ /* ed_band_cal_semi_64_w_absent_diag, Levenshtein_distance.h:3727-3776 with
     ed_core_64 (3116-3125).  Returns ez->err (INT32_MAX = not aligned), *pe = ez->pe */
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
	int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = thre + (thre << 1), best = INT32_MAX, site, ai, uge = INT32_MAX, c;
	*pe = -1;
	if (pn > tn + cut || tn > pn + cut) return best;
	bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
	for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[nt4(pstr[i])] |= mm; mm <<= 1; }
	i_bd = (thre << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
	Peq[4] = 0; mm = 1ULL << (thre << 1);
	for (i = 0; i <= tn0; i++) {
		X = Peq[nt4(tstr[i])] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0; HP = VN | ~(VP | D0);
		X = D0 >> 1;
		VN = X & HP; VP = HN | ~(X | HP);
		if (!(D0 & 1ULL)) { ++err; if (err > cut) return best; }
		if (i == tn0) break;
		Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
		++i_bd; c = 4;
		if (i_bd < pn) c = nt4(pstr[i_bd]);
		if (c < 4) Peq[c] |= mm;
	}
	site = tn - 1 - abs_diag; ai = pn - tn + abs_diag;
	for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
	if (err <= thre && err <= best) { best = err; *pe = site; }
	site -= i;
	while (i < ai) {
		err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
		if (err <= thre && err <= best) { best = err; *pe = site + i; }
		if (i == thre) uge = err;
	}
	if (uge <= thre && uge == best) *pe = site + thre;
	return best;
}

/* table dumps for the host-emulation harness (tests/hostemu) */
void hao_ft_dump(const hao_ft_t *ft, uint64_t *key, int32_t *val)
{
	uint64_t i;
	for (i = 0; i < ft->n; i++) { key[i] = ft->key[i]; val[i] = ft->val[i] == INT16_MAX ? INT32_MAX : ft->val[i]; }
}
void hao_pt_dump(const hao_pt_t *pt, uint64_t *key, uint64_t *off, uint32_t *cnt, uint64_t *pos)
{
	memcpy(key, pt->key, pt->n_keys * 8); memcpy(off, pt->off, pt->n_keys * 8);
	memcpy(cnt, pt->cnt, pt->n_keys * 4); memcpy(pos, pt->pos, pt->tot_pos * 8);
}

/* ------------------------------------------------------------------ */
/* window pass of the EC rounds (row a8)                               */
/* ------------------------------------------------------------------ */

static int64_t y_start_off(int64_t x_start, const uint64_t *fc, uint32_t n, int *bad)
{ /* y_start_offset, Hash_Table.h:165-189 */
	int64_t i;
	if (x_start == (int64_t)(fc[n - 1] >> 32)) return fc_shift(fc[n - 1]);
	for (i = 0; i < (int64_t)n; i++)
		if (x_start < (int64_t)(fc[i] >> 32)) break;
	if (i == 0 || i == (int64_t)n) { *bad = 1; return 0; }
	return fc_shift(fc[i - 1]);
}

static int init_waln_(int64_t err, int64_t s, int64_t l, int64_t w_l, int64_t *aux_beg, int64_t *aux_end, int64_t *r_s, int64_t *r_l)
{ /* init_waln, Correct.cpp:764-780 (THRESHOLD_MAX_SIZE = 31) */
	*aux_beg = *aux_end = *r_s = *r_l = -1;
	if (s < 0 || s >= l || (l - s + 2 * err + 31) < w_l) return 0;
	*aux_beg = *aux_end = 0;
	*r_s = s - err;
	*r_l = l - *r_s; if (*r_l > w_l) *r_l = w_l;
	*aux_end = w_l - *r_l;
	if (*r_s < 0) { *aux_beg = -*r_s; *r_s = 0; *r_l -= *aux_beg; }
	return 1;
}

typedef struct { int32_t chain, q_s, q_e, t_s, t_pri_l, thre, aux_beg, aux_end, err, pe; } hao_win_t;

int hao_windows(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const uint64_t *fc, double e_rate, int64_t w_l,
                hao_win_t **out, uint32_t *n_out)
{ /* the per-window work of align_hc_ed_post_extz (Correct.cpp:12951-13011) for every window of every chain,
     without its early exit: window grid (get_num_wins 783, get_win_se_by_normalize_xs Correct.h:1328),
     threshold (Adjust_Threshold Correct.h:46), target start from the fake cigar, init_waln, banded Myers */
	uint64_t ql = r->len[rid], tm = 0; uint32_t j, n = 0, m = 0; hao_win_t *w = 0; int bad = 0;
	char *qs = MALLOC_N(char, ql + 1), *ts = 0;
	hao_decode(r, rid, qs);
	for (j = 0; j < n_ch; j++) {
		const hao_ovlp_t *z = &ch[j];
		int64_t xs = z->x_pos_s, xe = z->x_pos_e, n_s = (xs / w_l) * w_l, nl = (xe + 1) - n_s, nw = nl / w_l + (nl % w_l > 0 ? 1 : 0), k;
		int64_t q_s = n_s < xs ? xs : n_s, q_e = n_s + w_l - 1 > xe ? xe : n_s + w_l - 1;
		for (k = 0; k < nw; k++) {
			int64_t q_l = 1 + q_e - q_s, thre = (int64_t)(q_l * e_rate), t_s, aln_l, aux_beg, aux_end, t_pri_l, t_tot_l = (int64_t)r->len[z->y_id];
			hao_win_t rec;
			if (thre == 0 && q_l >= 4) thre = 1;
			if (thre > 31) thre = 31;
			t_s = (q_s - xs) + z->y_pos_s;
			t_s += y_start_off(q_s, fc + z->fc_off, z->fc_n, &bad);
			aln_l = q_l + (thre << 1);
			rec.chain = (int32_t)j; rec.q_s = (int32_t)q_s; rec.q_e = (int32_t)q_e; rec.t_s = (int32_t)t_s; rec.t_pri_l = -1; rec.thre = (int32_t)thre;
			rec.aux_beg = rec.aux_end = 0; rec.err = INT32_MAX; rec.pe = -1;
			if (init_waln_(thre, t_s, t_tot_l, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) {
				int32_t pe;
				if ((uint64_t)t_pri_l + 1 > tm) { tm = t_pri_l + 64; ts = (char *)realloc(ts, tm); }
				hao_decode_sub(r, z->y_id, t_s, t_pri_l, (int)z->y_pos_strand, ts);
				rec.err = hao_ed_semi_64_absent_diag(ts, (int32_t)t_pri_l, qs + q_s, (int32_t)q_l, (int32_t)thre, (int32_t)aux_beg, &pe);
				rec.pe = pe; rec.t_s = (int32_t)t_s; rec.t_pri_l = (int32_t)t_pri_l; rec.aux_beg = (int32_t)aux_beg; rec.aux_end = (int32_t)aux_end;
			} else rec.t_s = -1; /* init_waln resets its r_s to -1 when it rejects the window (Correct.cpp:766) */
			if (n == m) { m = m ? m << 1 : 256; w = (hao_win_t *)realloc(w, m * sizeof(hao_win_t)); }
			w[n++] = rec;
			q_s = q_e + 1; q_e = q_s + w_l - 1;
			if (q_e >= xe) q_e = xe;
		}
	}
	free(qs); free(ts);
	*out = w; *n_out = n;
	return bad ? -1 : 0;
}

/* the alignment stage of the EC rounds (rows a8-a11) lives in its own file but shares this translation unit's helpers */
#include "ha_ec.c"
