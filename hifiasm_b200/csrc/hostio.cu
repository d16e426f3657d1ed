// hostio.cu — the stage's on-disk products, written by the host side of the library in the reference's own formats
// (§8(f) row 2).  Host code only: plain buffered stdio over the flat arrays the C-ABI hands around.
//   <o>.ovlp.paf              Output_PAF              Assembly.cpp:1673-1717   (--write-paf; R_INF.paf only)
//   <o>.ec.fa                 Output_corrected_reads  Assembly.cpp:884-905     (--write-ec)
//   <o>.ovlp.{source,reverse}.bin  write_ma_hit_ts    Overlaps.cpp:23442-23465 (fields one by one: 42 bytes per record, Overlaps.cpp:23420-23439)
//   <o>.ec.bin                write_All_reads         Process_Read.cpp:69-125
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
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
