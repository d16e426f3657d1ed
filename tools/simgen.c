/* simgen.c — seeded synthetic HiFi read sets at benchmark scale (bench / test tooling, host only; not part of the product).
 *
 * The shapes SURVEY.md §8(d) names: uniform-random diploid genome (haplotype 2 = haplotype 1 + SNPs) with repeat families (e.g. 5 % of the
 * genome in 50-copy 5 kb repeats, copies diverged by ~1 %), reads sampled uniformly from both haplotypes and strands, Gaussian lengths
 * clipped at min_len, HiFi errors split sub / ins / del 35 / 30 / 35 %, optional N bases.  Every read has its own counter-based random
 * stream (seed, read id), so the set is identical however many threads build it, and the two passes (sizes, then bases) agree.
 * Output: hifiasm's own 2-bit packing (ha_compress_base, Process_Read.cpp:792: len/4 + 1 bytes per read, N stored as A + a side list),
 * which is what hb_reads_upload takes, and optionally the same reads as FASTA for the reference binary.                               */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rng_next(rng_t *r) { uint64_t z = (r->s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
static inline double rng_u(rng_t *r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline rng_t rng_seed(uint64_t seed, uint64_t stream) { rng_t r; r.s = seed * 0xd1342543de82ef95ull + stream * 0x2545f4914f6cdd1dull + 0x632be59bd9b4e019ull; rng_next(&r); rng_next(&r); return r; }
static inline uint64_t rng_geom(rng_t *r, double log1mp) { /* failures before the first success */ double u = rng_u(r); if (u <= 0) u = 1e-300; return (uint64_t)(log(u) / log1mp); }

/* haplotypes as one base code (0..3) per byte */
void simgen_genome(uint64_t glen, uint64_t seed, double snp_rate, double repeat_frac, uint32_t repeat_len, uint32_t copies_per_family, double repeat_div,
                   uint8_t *hap1, uint8_t *hap2)
{
	rng_t r = rng_seed(seed, 1);
	for (uint64_t i = 0; i < glen; i += 32) { uint64_t z = rng_next(&r); for (int j = 0; j < 32 && i + j < glen; j++) hap1[i + j] = (z >> (2 * j)) & 3; }
	if (repeat_frac > 0 && glen > 4ull * repeat_len && copies_per_family >= 2) {
		uint64_t n_inst = (uint64_t)(glen * repeat_frac / repeat_len), n_fam = n_inst / copies_per_family; if (n_fam < 1) n_fam = 1;
		uint8_t *unit = (uint8_t *)malloc(repeat_len);
		double l1 = repeat_div > 0 ? log(1.0 - repeat_div) : 0;
		for (uint64_t f = 0; f < n_fam; f++) {
			rng_t q = rng_seed(seed, 1000 + f);
			for (uint32_t j = 0; j < repeat_len; j++) unit[j] = rng_next(&q) & 3;
			for (uint32_t c = 0; c < copies_per_family; c++) {
				uint64_t s = (uint64_t)(rng_u(&q) * (double)(glen - repeat_len));
				int rev = rng_next(&q) & 1;
				for (uint32_t j = 0; j < repeat_len; j++) hap1[s + j] = rev ? 3 - unit[repeat_len - 1 - j] : unit[j];
				if (repeat_div > 0) for (uint64_t j = rng_geom(&q, l1); j < repeat_len; j += 1 + rng_geom(&q, l1)) hap1[s + j] = (hap1[s + j] + 1 + rng_next(&q) % 3) & 3;
			}
		}
		free(unit);
	}
	memcpy(hap2, hap1, glen);
	if (snp_rate > 0) { rng_t q = rng_seed(seed, 2); double l1 = log(1.0 - snp_rate); for (uint64_t j = rng_geom(&q, l1); j < glen; j += 1 + rng_geom(&q, l1)) hap2[j] = (hap2[j] + 1 + rng_next(&q) % 3) & 3; }
}

typedef struct {
	uint64_t glen, seed, n_reads; double mean_len, sd_len; uint32_t min_len; double err, n_rate;
} simgen_par_t;

/* one read into out[] (codes 0..3, 4 = N); returns its length.  out == NULL: length only.  cap = capacity of out.                     */
static uint64_t gen_read(const simgen_par_t *p, const uint8_t *hap1, const uint8_t *hap2, uint64_t i, uint8_t *out, uint64_t cap)
{
	rng_t r = rng_seed(p->seed, (i << 1) | 1);
	double u1 = rng_u(&r), u2 = rng_u(&r); if (u1 <= 0) u1 = 1e-300;
	double g = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
	int64_t len = (int64_t)(p->mean_len + p->sd_len * g); if (len < (int64_t)p->min_len) len = p->min_len; if ((uint64_t)len > p->glen) len = (int64_t)p->glen;
	uint64_t start = (uint64_t)(rng_u(&r) * (double)(p->glen - (uint64_t)len + 1));
	uint64_t bits = rng_next(&r); const uint8_t *src = (bits & 1) ? hap2 : hap1; const int strand = (bits >> 1) & 1;
	const double l1 = p->err > 0 ? log(1.0 - p->err) : 0, ln = p->n_rate > 0 ? log(1.0 - p->n_rate) : 0;
	uint64_t next_err = p->err > 0 ? rng_geom(&r, l1) : UINT64_MAX, o = 0;
	for (uint64_t j = 0; j < (uint64_t)len; j++) {
		uint8_t c = strand ? 3 - src[start + (uint64_t)len - 1 - j] : src[start + j];
		if (j == next_err) {
			double k = rng_u(&r); next_err = j + 1 + rng_geom(&r, l1);
			if (k < 0.35) { c = (c + 1 + rng_next(&r) % 3) & 3; if (out && o < cap) out[o]= c; o++; }
			else if (k < 0.65) { if (out && o < cap) out[o] = c; o++; uint8_t x = rng_next(&r) & 3; if (out && o < cap) out[o] = x; o++; }
			/* else: deletion */
		} else { if (out && o < cap) out[o] = c; o++; }
	}
	if (p->n_rate > 0 && out) { rng_t q = rng_seed(p->seed ^ 0x5bd1e995u, (i << 1) | 1); for (uint64_t j = rng_geom(&q, ln); j < o && j < cap; j += 1 + rng_geom(&q, ln)) out[j] = 4; }
	return o;
}

/* pass 1: read lengths (after errors) */
void simgen_lengths(const simgen_par_t *p, const uint8_t *hap1, const uint8_t *hap2, uint64_t *length)
{
#pragma omp parallel for schedule(dynamic, 256)
	for (int64_t i = 0; i < (int64_t)p->n_reads; i++) length[i] = gen_read(p, hap1, hap2, (uint64_t)i, 0, 0);
}
/* how many N bases each read carries (0 when n_rate == 0) */
void simgen_ncount(const simgen_par_t *p, const uint8_t *hap1, const uint8_t *hap2, const uint64_t *length, uint64_t *n_cnt)
{
	uint64_t mx = 0; for (uint64_t i = 0; i < p->n_reads; i++) if (length[i] > mx) mx = length[i];
#pragma omp parallel
	{
		uint8_t *buf = (uint8_t *)malloc(mx + 8);
#pragma omp for schedule(dynamic, 256)
		for (int64_t i = 0; i < (int64_t)p->n_reads; i++) {
			uint64_t c = 0;
			if (p->n_rate > 0) { uint64_t L = gen_read(p, hap1, hap2, (uint64_t)i, buf, mx); for (uint64_t j = 0; j < L; j++) c += buf[j] == 4; }
			n_cnt[i] = c;
		}
		free(buf);
	}
}
/* pass 2: bases.  packed / byte_off as hb_reads_upload takes them; n_pos / n_off the N lists; fasta (optional) receives ">r<i>\n<bases>\n"
 * at fa_off[i] (fa_off computed by the caller from the lengths: name length + len + 3).                                                  */
void simgen_fill(const simgen_par_t *p, const uint8_t *hap1, const uint8_t *hap2, const uint64_t *length, const uint64_t *byte_off, uint8_t *packed,
                 const uint64_t *n_off, uint64_t *n_pos, const uint64_t *fa_off, char *fasta)
{
	uint64_t mx = 0; for (uint64_t i = 0; i < p->n_reads; i++) if (length[i] > mx) mx = length[i];
#pragma omp parallel
	{
		uint8_t *buf = (uint8_t *)malloc(mx + 8);
#pragma omp for schedule(dynamic, 256)
		for (int64_t i = 0; i < (int64_t)p->n_reads; i++) {
			uint64_t L = gen_read(p, hap1, hap2, (uint64_t)i, buf, mx), k = n_off ? n_off[i] : 0;
			if (fasta) {
				char *f = fasta + fa_off[i]; f += sprintf(f, ">r%lld\n", (long long)i);
				for (uint64_t j = 0; j < L; j++) f[j] = "ACGTN"[buf[j]];
				f[L] = '\n';
			}
			for (uint64_t j = 0; j < L; j++) if (buf[j] == 4) { if (n_pos) n_pos[k++] = j; buf[j] = 0; }
			buf[L] = buf[L + 1] = buf[L + 2] = buf[L + 3] = 0;
			uint8_t *o = packed + byte_off[i];
			for (uint64_t j = 0, b = 0; b < L / 4 + 1; j += 4, b++) o[b] = (uint8_t)(buf[j] << 6 | buf[j + 1] << 4 | buf[j + 2] << 2 | buf[j + 3]);
		}
		free(buf);
	}
}
