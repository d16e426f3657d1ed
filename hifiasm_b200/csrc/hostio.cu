// hostio.cu — the stage's on-disk products, written by the host side of the library in the reference's own formats
// (§8(f) row 2).  Host code only: plain buffered stdio over the flat arrays the C-ABI hands around.
//   <o>.ovlp.paf              Output_PAF              Assembly.cpp:1673-1717   (--write-paf; R_INF.paf only)
//   <o>.ec.fa                 Output_corrected_reads  Assembly.cpp:884-905     (--write-ec)
//   <o>.ovlp.{source,reverse}.bin  write_ma_hit_ts    Overlaps.cpp:23442-23465 (fields one by one: 42 bytes per record, Overlaps.cpp:23420-23439)
//   <o>.ec.bin                write_All_reads         Process_Read.cpp:69-125
// and the way in: FASTA / FASTQ (gzip or plain) -> the flat All_reads layout hb_reads_upload takes (hb_readset_load).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
#include <zlib.h>
#include "../../include/hifiasm_b200.h"

namespace {
struct Out { // buffered writer; any failure is sticky
	FILE *f; std::vector<char> b; size_t n; bool bad;
	Out(const char *path) : f(fopen(path, "wb")), b(1 << 22), n(0), bad(f == 0) {}
	void flush() { if (f && n) { if (fwrite(b.data(), 1, n, f) != n) bad = true; } n = 0; }
	void put(const void *p, size_t len)
	{
		if (bad) return;
		if (len >= b.size()) { flush(); if (fwrite(p, 1, len, f) != len) bad = true; return; }
		if (n + len > b.size()) flush();
		memcpy(b.data() + n, p, len); n += len;
	}
	void ch(char c) { put(&c, 1); }
	void dec(uint64_t v) { char t[24]; int k = 0; do { t[k++] = (char)('0' + v % 10); v /= 10; } while (v); char r[24]; for (int i = 0; i < k; i++) r[i] = t[k - 1 - i]; put(r, (size_t)k); }
	void sdec(int64_t v) { if (v < 0) { ch('-'); dec((uint64_t)(-v)); } else dec((uint64_t)v); }
	template <typename T> void raw(const T &v) { put(&v, sizeof(T)); }
	int close() { flush(); if (f && fclose(f)) bad = true; f = 0; return bad ? HB_E_IO : HB_OK; }
	~Out() { if (f) fclose(f); }
};
}

extern "C" {

int hb_write_paf(const char *path, uint64_t n_reads, const uint64_t *read_length, const char *names, const uint64_t *name_index,
                 const uint64_t *off, const hb_ma_hit_t *rec)
{
	Out o(path); if (o.bad) return HB_E_IO;
	for (uint64_t i = 0; i < n_reads; i++)
		for (uint64_t j = off[i]; j < off[i + 1]; j++) {
			const hb_ma_hit_t &h = rec[j]; const uint64_t qn = h.qns >> 32, tn = h.tn;
			if (qn >= n_reads || tn >= n_reads) return HB_E_ARG;
			o.put(names + name_index[qn], (size_t)(name_index[qn + 1] - name_index[qn])); o.ch('\t');
			o.dec(read_length[qn]); o.ch('\t');
			o.sdec((int32_t)(uint32_t)(h.qns & 0xffffffffu)); o.ch('\t');   // the reference prints qs / qe / ts / te / ml / bl with %d
			o.sdec((int32_t)h.qe); o.ch('\t');
			o.ch(h.rev ? '-' : '+'); o.ch('\t');
			o.put(names + name_index[tn], (size_t)(name_index[tn + 1] - name_index[tn])); o.ch('\t');
			o.dec(read_length[tn]); o.ch('\t');
			o.sdec((int32_t)h.ts); o.ch('\t'); o.sdec((int32_t)h.te); o.ch('\t'); o.sdec((int32_t)h.ml); o.ch('\t'); o.sdec((int32_t)h.bl); o.put("\t255\n", 5);
		}
	return o.close();
}

int hb_write_ec_fa(const char *path, uint64_t n_reads, const uint64_t *read_length, const uint8_t *packed, const uint64_t *byte_off,
                   const uint64_t *n_off, const uint64_t *n_pos, const char *names, const uint64_t *name_index)
{
	Out o(path); if (o.bad) return HB_E_IO;
	std::vector<char> seq;
	for (uint64_t i = 0; i < n_reads; i++) {
		const uint64_t L = read_length[i]; const uint8_t *p = packed + byte_off[i];
		if (seq.size() < L + 4) seq.resize(L + 4);
		for (uint64_t k = 0; k < (L + 3) / 4; k++) { const uint8_t b = p[k]; char *s = seq.data() + 4 * k; s[0] = "ACGT"[b >> 6]; s[1] = "ACGT"[(b >> 4) & 3]; s[2] = "ACGT"[(b >> 2) & 3]; s[3] = "ACGT"[b & 3]; } // recover_UC_Read, Process_Read.cpp:716
		for (uint64_t k = n_off[i]; k < n_off[i + 1]; k++) { if (n_pos[k] >= L) return HB_E_ARG; seq[n_pos[k]] = 'N'; }
		o.ch('>'); o.put(names + name_index[i], (size_t)(name_index[i + 1] - name_index[i])); o.ch('\n');
		o.put(seq.data(), (size_t)L); o.ch('\n');
	}
	return o.close();
}

int hb_write_ovlp_bin(const char *path, uint64_t n_reads, const uint64_t *off, const hb_ma_hit_t *rec, const uint8_t *is_fully_corrected, const uint8_t *is_abnormal)
{
	Out o(path); if (o.bad) return HB_E_IO;
	o.raw((int64_t)n_reads);
	for (uint64_t i = 0; i < n_reads; i++) {
		o.raw((uint8_t)(is_fully_corrected ? is_fully_corrected[i] : 0)); o.raw((uint8_t)(is_abnormal ? is_abnormal[i] : 0)); o.raw((uint32_t)(off[i + 1] - off[i]));
		for (uint64_t j = off[i]; j < off[i + 1]; j++) {
			const hb_ma_hit_t &h = rec[j];
			o.raw((uint64_t)h.qns); o.raw((uint32_t)h.qe); o.raw((uint32_t)h.tn); o.raw((uint32_t)h.ts); o.raw((uint32_t)h.te); o.raw((uint8_t)h.el); o.raw((uint8_t)h.no_l_indel);
			o.raw((uint32_t)h.ml); o.raw((uint32_t)h.rev); o.raw((uint32_t)h.bl); o.raw((uint32_t)h.del);
		}
	}
	return o.close();
}

int hb_write_ec_bin(const char *path, int32_t adapter_len, uint64_t index_size, uint64_t name_index_size, uint64_t n_reads, uint64_t total_reads_bases,
                    const uint64_t *read_length, const uint8_t *packed, const uint64_t *byte_off, const uint64_t *n_off, const uint64_t *n_pos,
                    const char *names, uint64_t total_name_length, const uint64_t *name_index, const uint8_t *trio_flag, int32_t hom_cov, int32_t het_cov)
{
	Out o(path); if (o.bad) return HB_E_IO;
	o.raw(adapter_len); o.raw(index_size); o.raw(name_index_size); o.raw(n_reads); o.raw(total_reads_bases); o.raw(total_name_length);
	for (uint64_t i = 0; i < n_reads; i++) { const uint64_t nn = n_off[i + 1] - n_off[i]; o.raw(nn); if (nn) o.put(n_pos + n_off[i], (size_t)nn * 8); }
	o.put(read_length, (size_t)n_reads * 8);
	for (uint64_t i = 0; i < n_reads; i++) {
		if (byte_off[i + 1] - byte_off[i] != read_length[i] / 4 + 1) return HB_E_ARG; // All_reads keeps len/4+1 bytes per read (Process_Read.cpp:104)
		o.put(packed + byte_off[i], (size_t)(read_length[i] / 4 + 1));
	}
	o.put(names, (size_t)total_name_length);
	o.put(name_index, (size_t)name_index_size * 8);
	if (trio_flag) o.put(trio_flag, (size_t)n_reads); else { const std::vector<uint8_t> z(n_reads, 0); o.put(z.data(), (size_t)n_reads); }
	o.raw(hom_cov); o.raw(het_cov);
	return o.close();
}

} // extern "C"

// ---- ingest: FASTA / FASTQ (plain or gzip) -> the All_reads layout (htab.cpp:761-813 step 0 of the counting pipeline with HAF_RS_WRITE_LEN /
// HAF_RS_WRITE_SEQ: ha_insert_read_len Process_Read.cpp:414, ha_compress_base 792, names up to the first blank like kseq).  HiFi mode:
// no length cut (CommandLines.cpp:1043), reads that are empty after trimming adapter_len bases from both ends are skipped (htab.cpp:762-763).
namespace {
struct Gz { // line reader over zlib (gzopen reads plain files too)
	gzFile f; std::vector<char> b; int n, i; bool eof;
	Gz(const char *p) : f(gzopen(p, "rb")), b(1 << 20), n(0), i(0), eof(false) {}
	~Gz() { if (f) gzclose(f); }
	int getc_() { if (i >= n) { if (eof) return -1; n = gzread(f, b.data(), (unsigned)b.size()); i = 0; if (n <= 0) { eof = true; n = 0; return -1; } } return (unsigned char)b[i++]; }
	bool line(std::string &s) { s.clear(); int c = getc_(); if (c < 0) return false; for (; c >= 0 && c != '\n'; c = getc_()) s.push_back((char)c); if (!s.empty() && s.back() == '\r') s.pop_back(); return true; }
};
template <typename T> T *dup_vec(const std::vector<T> &v) { T *p = (T *)malloc((v.size() ? v.size() : 1) * sizeof(T)); if (p && !v.empty()) memcpy(p, v.data(), v.size() * sizeof(T)); return p; }
}

extern "C" {

void hb_readset_free(hb_readset_t *rs)
{
	if (!rs) return;
	free(rs->read_length); free(rs->byte_off); free(rs->packed); free(rs->n_off); free(rs->n_pos); free(rs->names); free(rs->name_index);
	memset(rs, 0, sizeof(*rs));
}

int hb_readset_load(const char *const *paths, int n_paths, int32_t adapter_len, hb_readset_t *out)
{
	static const uint8_t nt6[256] = { // seq_nt6_table, Process_Read.cpp:12-29: ACGT / acgt -> 0..3, everything else is stored as an N site
#define R16 5,5,5,5,5,5,5,5,5,5,5,5,5,5,5,5
		R16, R16, R16, R16, 5,0,5,1,5,5,5,2,5,5,5,5,5,5,5,5, 5,5,5,5,3,5,5,5,5,5,5,5,5,5,5,5, 5,0,5,1,5,5,5,2,5,5,5,5,5,5,5,5, 5,5,5,5,3,5,5,5,5,5,5,5,5,5,5,5,
		R16, R16, R16, R16, R16, R16, R16, R16
#undef R16
	};
	std::vector<uint64_t> len, boff(1, 0), noff(1, 0), npos, nidx(1, 0); std::vector<uint8_t> pk; std::vector<char> names;
	uint64_t index_size = 1000, name_index_size = 1000, total = 0; // READ_INIT_NUMBER, Process_Read.h:13; growth rule of ha_insert_read_len
	memset(out, 0, sizeof(*out));
	if (adapter_len < 0) return HB_E_ARG;
	for (int fi = 0; fi < n_paths; fi++) {
		Gz g(paths[fi]); if (!g.f) return HB_E_IO;
		std::string ln, name, seq, q; bool have = g.line(ln);
		while (have) {
			if (ln.empty() || (ln[0] != '>' && ln[0] != '@')) { have = g.line(ln); continue; }
			const bool fq = ln[0] == '@';
			size_t e = 1; while (e < ln.size() && ln[e] != ' ' && ln[e] != '\t') e++;
			name.assign(ln, 1, e - 1); seq.clear();
			while ((have = g.line(ln))) { if (!ln.empty() && (ln[0] == '>' || ln[0] == '+' || (ln[0] == '@' && !fq))) break; seq += ln; }
			if (fq && have && !ln.empty() && ln[0] == '+') { q.clear(); while ((have = g.line(ln))) { q += ln; if (q.size() >= seq.size()) { have = g.line(ln); break; } } }
			const int64_t l = (int64_t)seq.size() - 2 * (int64_t)adapter_len;
			if (l <= 0) continue;
			if (len.size() >= (1u << 28)) return HB_E_OVERFLOW; // "this implementation supports no more than 2^28 reads", htab.cpp:765
			const char *sq = seq.data() + adapter_len; const uint64_t L = (uint64_t)l, o = pk.size();
			pk.resize(o + L / 4 + 1, 0); // len/4+1 bytes per read (malloc_All_reads, Process_Read.cpp:443): the extra byte of a read with len % 4 == 0 stays 0 here
			for (uint64_t i = 0; i < L; i++) { uint8_t c = nt6[(uint8_t)sq[i]]; if (c >= 4) { c = 0; npos.push_back(i); } pk[o + (i >> 2)] |= (uint8_t)(c << (6 - 2 * (i & 3))); }
			len.push_back(L); boff.push_back(pk.size()); noff.push_back(npos.size()); total += L;
			names.insert(names.end(), name.begin(), name.end()); nidx.push_back(names.size());
			if (index_size < len.size() + 2) { index_size = index_size * 2 + 2; name_index_size = name_index_size * 2 + 2; }
		}
	}
	out->n_reads = len.size(); out->total_bases = total; out->total_name_length = names.size(); out->index_size = index_size; out->name_index_size = name_index_size;
	nidx.resize(name_index_size, 0); // the reference writes name_index_size entries of which n_reads + 1 are defined
	out->read_length = dup_vec(len); out->byte_off = dup_vec(boff); out->packed = dup_vec(pk); out->n_off = dup_vec(noff); out->n_pos = dup_vec(npos);
	out->names = dup_vec(names); out->name_index = dup_vec(nidx);
	if (!out->read_length || !out->byte_off || !out->packed || !out->n_off || !out->n_pos || !out->names || !out->name_index) { hb_readset_free(out); return HB_E_NOMEM; }
	return HB_OK;
}

} // extern "C"
