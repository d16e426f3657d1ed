/* oracle/ha_ec.c — TEST INFRASTRUCTURE ONLY (part of libha_oracle.so; included by ha_oracle.c).
 *
 * CPU restatement of the alignment stage of an error-correction round of hifiasm v0.25.0-r726
 * (SURVEY.md §8 rows a8-a11): what gen_hc_r_alin (Correct.cpp:25617-25675) does to every overlap
 * h_ec_lchain returned.  Works on decoded strings like the reference; every routine names the lines it
 * restates.  Pinned against dumps of the unmodified reference (oracle/refdump.cpp step 6, tests/golden).
 */

#define EC_THRE_MAX 31   /* THRESHOLD_MAX_SIZE, Hash_Table.h:24 */
#define EC_OVLP_CUT 0.9  /* OVERLAP_THRESHOLD_HIFI_FILTER, Hash_Table.h:19 */

/* window_list, Hash_Table.h:54-62 (same 32-byte layout) */
typedef struct {
	int32_t x_start, x_end, y_start, y_end;
	int16_t extra_begin, extra_end, error, error_threshold;
	uint32_t cidx, clen;
} hao_wl_t;

typedef struct { int32_t st; uint32_t align_length; double rr; int64_t re; uint64_t w_off, w_n, c_off, c_n; } hao_alnA_t;

/* bit_extz_t (Levenshtein_distance.h:776-785), the fields this path uses */
typedef struct {
	int32_t ps, pe, pl, ts, te, tl, thre, err;
	uint16_t *cig; size_t cn, cm;
	uint64_t *path; size_t pn, pm;
} exz_t;

/* overlap_region + its window_list_alloc (Hash_Table.h:64-106) */
typedef struct {
	int64_t x_pos_s, x_pos_e, y_pos_s, y_pos_e; uint32_t y_id, rev;
	const uint64_t *fc; uint32_t fc_n;
	int64_t align_length;
	hao_wl_t *w; size_t wn, wm;
	uint16_t *c; size_t cn, cm;
} ecz_t;

typedef struct { const hao_reads_t *r; const char *qstr; char *tstr; size_t tm; exz_t ez; int bad; } ecb_t;

static void ec_tbuf(ecb_t *b, int64_t n) { if ((size_t)n + 8 > b->tm) { b->tm = (size_t)n + 64; b->tstr = (char *)realloc(b->tstr, b->tm); } }

static inline void cig_push(uint16_t **a, size_t *n, size_t *m, uint16_t v)
{
	if (*n == *m) { *m = *m ? *m << 1 : 64; *a = (uint16_t *)realloc(*a, *m * 2); }
	(*a)[(*n)++] = v;
}
static void ez_push_trace(exz_t *ez, uint16_t c, uint32_t len)
{ /* push_trace, Levenshtein_distance.h:522-531 */
	c <<= 14;
	while (len >= 0x3fff) { cig_push(&ez->cig, &ez->cn, &ez->cm, (uint16_t)(c + 0x3fff)); len -= 0x3fff; }
	if (len) cig_push(&ez->cig, &ez->cn, &ez->cm, (uint16_t)(c + len));
}

static void ez_gen_trace(exz_t *ez, int32_t ptrim, int reverse)
{ /* gen_trace, Levenshtein_distance.h:903-985, one word per column (nword = 1) */
	if (ez->err > ez->thre) return;
	ez->cn = 0;
	int32_t V, H, D, min, cur, tn = ez->te + 1 - ez->ts, pn = tn + (ez->thre << 1), bd = (ez->thre << 1) + 1;
	int32_t poff = ez->pe, sft = bd - (pn - ez->pe - ptrim);
	int32_t i = tn, low = bd - 1, d = 0, pd = -1, pdn = 0; cur = ez->err;
	while (i > 0 && cur > 0) {
		const uint64_t *D0 = ez->path + (size_t)(i - 1) * 5, *VP = D0 + 1, *VN = D0 + 2, *HP = D0 + 3, *HN = D0 + 4;
		D = cur - (int32_t)((~(D0[0] >> sft)) & 1ULL); d = 0; min = D;
		H = V = INT32_MAX;
		if (sft != low) {
			H = cur + (int32_t)((HN[0] >> sft) & 1ULL) - (int32_t)((HP[0] >> sft) & 1ULL);
			if (H + 1 == cur && H <= min) { min = H; d = 3; }
		}
		if (sft != 0) {
			V = cur + (int32_t)((VN[0] >> (sft - 1)) & 1ULL) - (int32_t)((VP[0] >> (sft - 1)) & 1ULL);
			if (V + 1 == cur && V <= min) { min = V; d = 2; }
		}
		if (d == 0) { if (D != cur) d = 1; i--; poff--; }
		else if (d == 2) { sft--; poff--; }
		else if (d == 3) { i--; sft++; }
		if (d == pd) pdn++;
		else { if (pdn > 0) ez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = 1; }
		cur = min;
	}
	if (i > 0) {
		d = 0; poff -= i;
		if (d == pd) pdn += i;
		else { if (pdn > 0) ez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	poff++;
	if (ez->ps < 0 || ez->ps >= ez->pl) ez->ps = poff;
	else if (poff > ez->ps) {
		d = 2; i = poff - ez->ps;
		if (d == pd) pdn += i;
		else { if (pdn > 0) ez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	if (pdn > 0) ez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn);
	if (reverse) {
		size_t k, h = ez->cn >> 1;
		for (k = 0; k < h; k++) { uint16_t t = ez->cig[k]; ez->cig[k] = ez->cig[ez->cn - k - 1]; ez->cig[ez->cn - k - 1] = t; }
	}
}

static void ez_semi64_trace(const char *pstr, int32_t pn, const char *tstr, int32_t tn, int32_t thre, int32_t abs_diag, exz_t *ez)
{ /* ed_band_cal_semi_64_w_absent_diag_trace, Levenshtein_distance.h:3778-3856 */
	ez->cn = 0;
	if (ez->err > thre) {
		ez->thre = thre; ez->err = INT32_MAX; ez->pl = pn; ez->tl = tn; /* init_base_ed */
		ez->ps = ez->pe = -1; ez->ts = 0; ez->te = tn - 1;
	} else if (ez->err == 0) {
		ez_push_trace(ez, 0, (uint32_t)(ez->te + 1 - ez->ts));
		ez->ps = ez->pe - (ez->te - ez->ts);
		return;
	}
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN, X, D0, HN, HP, mm;
	int32_t bd, i, err = abs_diag, i_bd, tn0 = tn - 1, cut = thre + (thre << 1), c;
	if (pn > tn + cut || tn > pn + cut) return;
	bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn;
	for (i = 0, mm = 1ULL << abs_diag; i < bd; i++) { Peq[nt4(pstr[i])] |= mm; mm <<= 1; }
	i_bd = (thre << 1) - abs_diag; VN = (1ULL << abs_diag) - 1;
	if ((size_t)tn * 5 > ez->pm) { ez->pm = (size_t)tn * 5 + 64; ez->path = (uint64_t *)realloc(ez->path, ez->pm * 8); }
	ez->pn = 0;
	Peq[4] = 0; mm = 1ULL << (thre << 1);
	for (i = 0; i <= tn0; i++) {
		X = Peq[nt4(tstr[i])] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0; HP = VN | ~(VP | D0);
		X = D0 >> 1;
		VN = X & HP; VP = HN | ~(X | HP);
		if (!(D0 & 1ULL)) { ++err; if (err > cut) return; }
		if (i < tn0) {
			Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
			++i_bd; c = 4;
			if (i_bd < pn) c = nt4(pstr[i_bd]);
			if (c < 4) Peq[c] |= mm;
		}
		ez->path[ez->pn++] = D0; ez->path[ez->pn++] = VP; ez->path[ez->pn++] = VN; ez->path[ez->pn++] = HP; ez->path[ez->pn++] = HN;
	}
	if (ez->err > thre) {
		int32_t site = tn - 1 - abs_diag, ai = pn - tn + abs_diag, uge = INT32_MAX;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); }
		if (err <= thre && err <= ez->err) { ez->err = err; ez->pe = site; }
		site -= i;
		while (i < ai) {
			err += (int32_t)((VP >> i) & 1ULL); err -= (int32_t)((VN >> i) & 1ULL); ++i;
			if (err <= thre && err <= ez->err) { ez->err = err; ez->pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == ez->err) ez->pe = site + thre;
	}
	ez_gen_trace(ez, abs_diag, 1);
}

static void ez_semi64(const char *pstr, int32_t pn, const char *tstr, int32_t tn, int32_t thre, int32_t abs_diag, exz_t *ez)
{ /* ed_band_cal_semi_64_w_absent_diag, Levenshtein_distance.h:3727-3776: sets thre, err, pl, tl, pe (ps = -1, ts = 0, te = tn-1) */
	int32_t pe;
	ez->thre = thre; ez->pl = pn; ez->tl = tn; ez->ps = -1; ez->ts = 0; ez->te = tn - 1; ez->cn = 0;
	ez->err = hao_ed_semi_64_absent_diag(pstr, pn, tstr, tn, thre, abs_diag, &pe);
	ez->pe = pe;
}

/* ---- Reserve_Banded_BPM_Extension[_REV], Levenshtein_distance.h:71-214, 216-359 --------------------
 * The reference keeps Peq in a 256-entry table indexed by the character; only the A/C/G/T rows are
 * shifted per column, so the row of any other pattern character ('N') keeps the bits it received at the
 * band's top position.  Row 4 below is that 'N' row. */
static void get_error_(int t_length, int errthold, int init_err, uint64_t VP, uint64_t VN, unsigned int *return_err, int *back_site)
{ /* get_error, Levenshtein_distance.h:20-69 */
	*return_err = (unsigned int)-1;
	int site = t_length - 1, return_site = -1, available_i = 2 * errthold, i = 0;
	unsigned int ungap_error = (unsigned int)-1;
	if (init_err <= errthold && (unsigned int)init_err <= *return_err) { *return_err = (unsigned int)init_err; return_site = site; }
	while (i < available_i) {
		init_err += (int)((VP >> i) & 1ULL);
		init_err -= (int)((VN >> i) & 1ULL);
		++i;
		if (init_err <= errthold && (unsigned int)init_err <= *return_err) { *return_err = (unsigned int)init_err; return_site = site + i; }
		if (i == errthold) ungap_error = (unsigned int)init_err;
	}
	if (ungap_error <= (unsigned int)errthold && ungap_error == *return_err) return_site = site + errthold;
	*back_site = return_site;
}

static int bpm_extension(const char *pattern, int p_length, const char *text, int t_length, int errthold, int rev,
                         unsigned int *return_err, int *return_p_end, int *return_t_end)
{
	*return_err = (unsigned int)-1; *return_p_end = -1; *return_t_end = -1;
	uint64_t Peq[5] = { 0, 0, 0, 0, 0 }, VP = 0, VN = 0, X, D0, HN, HP, Mask = 1ULL << (errthold << 1), b = 1;
	unsigned int line_error = (unsigned int)-1; int return_site, band_length = (errthold << 1) + 1, i, err = 0, i_bd = errthold << 1, last_high = errthold << 1;
#define PAT(j) nt4(pattern[rev ? p_length - (j) - 1 : (j)])
#define TXT(j) nt4(text[rev ? t_length - (j) - 1 : (j)])
	for (i = 0; i < band_length; i++) { Peq[PAT(i)] |= b; b <<= 1; }
	Peq[4] = 0; /* the memset keeps only the A/C/G/T rows */
	for (i = 0; i < t_length; i++) {
		X = Peq[TXT(i)] | VN;
		D0 = ((VP + (X & VP)) ^ VP) | X;
		HN = VP & D0; HP = VN | ~(VP | D0);
		X = D0 >> 1;
		VN = X & HP; VP = HN | ~(X | HP);
		if (!(D0 & 1ULL)) { ++err; if (err - last_high > errthold) return *return_t_end; }
		get_error_(i + 1, errthold, err, VP, VN, &line_error, &return_site);
		if (line_error != (unsigned int)-1) {
			*return_t_end = rev ? t_length - i - 1 : i;
			*return_p_end = rev ? p_length - return_site - 1 : return_site;
			*return_err = line_error;
		}
		if (i == t_length - 1) break;
		Peq[0] >>= 1; Peq[1] >>= 1; Peq[2] >>= 1; Peq[3] >>= 1;
		++i_bd;
		Peq[PAT(i_bd)] |= Mask;
	}
#undef PAT
#undef TXT
	return *return_t_end;
}

/* ---- window grid and thresholds ------------------------------------------------------------------ */
static inline int64_t win_by_s(const ecz_t *z, int64_t w_s, int64_t bs, int64_t *w_e)
{ /* get_win_id_by_s, Correct.h:1306-1314 */
	int64_t n_s = (z->x_pos_s / bs) * bs, wid = (w_s - n_s) / bs;
	if (w_e) { *w_e = n_s + (wid + 1) * bs - 1; if (*w_e > z->x_pos_e) *w_e = z->x_pos_e; }
	return wid;
}
static inline int64_t win_by_e(const ecz_t *z, int64_t w_e, int64_t bs, int64_t *w_s)
{ /* get_win_id_by_e, Correct.h:1317-1325 */
	int64_t n_s = (z->x_pos_s / bs) * bs, wid = (w_e - n_s) / bs;
	if (w_s) { *w_s = n_s + wid * bs; if (*w_s < z->x_pos_s) *w_s = z->x_pos_s; }
	return wid;
}
static inline int64_t adj_thre(int64_t t, int64_t len) { return (t == 0 && len >= 4) ? 1 : t; } /* Adjust_Threshold, Correct.h:46 */
static inline int64_t init_err_thres(int64_t len, double e_rate, int64_t bs, int64_t block_err)
{ /* get_init_err_thres, Correct.cpp:1042-1049 */
	if (len >= bs) return block_err;
	int64_t t = (int64_t)(len * e_rate); t = adj_thre(t, len);
	if (t > EC_THRE_MAX) t = EC_THRE_MAX;
	return t;
}
static inline int dbl_err_thres(int pre, int x_len)
{ /* double_error_threshold, Correct.cpp:917-934 */
	pre = (int)adj_thre(pre, x_len);
	int t = pre * 2;
	if (x_len >= 300 && t < EC_THRE_MAX) t = EC_THRE_MAX;
	if (t > EC_THRE_MAX) t = EC_THRE_MAX;
	return t;
}

static hao_wl_t *wl_pushp(ecz_t *z)
{
	if (z->wn == z->wm) { z->wm = z->wm ? z->wm << 1 : 16; z->w = (hao_wl_t *)realloc(z->w, z->wm * sizeof(hao_wl_t)); }
	return &z->w[z->wn++];
}
static void push_wcigar_(hao_wl_t *p, ecz_t *z, const exz_t *ez)
{ /* push_wcigar, Correct.cpp:4050-4055 */
	size_t k;
	p->cidx = (uint32_t)z->cn; p->clen = (uint32_t)ez->cn;
	for (k = 0; k < ez->cn; k++) cig_push(&z->c, &z->cn, &z->cm, ez->cig[k]);
}

static int recal_boundary_(ecb_t *b, const char *q_string, int64_t ql0, int64_t tl0, int64_t thres, int64_t toff, int64_t ts0, int64_t te0, int64_t err0,
                           uint32_t tid, uint32_t rev, int64_t *ts_r, int64_t *aux_beg_r, int64_t *aux_end_r)
{ /* recal_boundary_exz, Correct.cpp:2429-2469 */
	int64_t ts, aux_beg, aux_end, t_pri_l, aln_l = ql0 + (thres << 1), t_tot_l = (int64_t)b->r->len[tid];
	if (ts0 == 0) ts = toff;
	else if (te0 + 1 == tl0) ts = toff + te0 - ql0 + 1;
	else return 0;
	if (!init_waln_(thres, ts, t_tot_l, aln_l, &aux_beg, &aux_end, &ts, &t_pri_l)) return 0;
	if (ts == toff && tl0 == t_pri_l) return 0;
	ec_tbuf(b, t_pri_l);
	hao_decode_sub(b->r, tid, ts, t_pri_l, (int)rev, b->tstr);
	b->ez.err = INT32_MAX;
	ez_semi64_trace(b->tstr, (int32_t)t_pri_l, q_string, (int32_t)ql0, (int32_t)thres, (int32_t)aux_beg, &b->ez);
	if (b->ez.err <= b->ez.thre && b->ez.err < err0) { *aux_beg_r = aux_beg; *aux_end_r = aux_end; *ts_r = ts; return 1; }
	return 0;
}

static uint32_t aln_wlst_adv_(ecb_t *b, ecz_t *z, uint32_t max_err, int64_t qs, int64_t qe, int64_t t_s, int64_t bs, double e_rate, int is_cigar)
{ /* aln_wlst_adv_exz, Correct.cpp:4057-4127 (rref branch) */
	int64_t ql = qe + 1 - qs, tl, aln_l, t_tot_l = (int64_t)b->r->len[z->y_id], aux_beg, aux_end, t_pri_l, thres;
	const char *q_string; exz_t *ez = &b->ez;
	thres = dbl_err_thres((int)init_err_thres(ql, e_rate, bs, max_err), (int)ql);
	aln_l = ql + (thres << 1);
	if (!init_waln_(thres, t_s, t_tot_l, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) return 0;
	if (t_pri_l + thres < ql) return 0;
	q_string = b->qstr + qs;
	ec_tbuf(b, t_pri_l);
	hao_decode_sub(b->r, z->y_id, t_s, t_pri_l, (int)z->rev, b->tstr);
	tl = t_pri_l;
	if (is_cigar) { ez->err = INT32_MAX; ez_semi64_trace(b->tstr, (int32_t)tl, q_string, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez); }
	else { ez_semi64(b->tstr, (int32_t)tl, q_string, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez); ez->ps = 0; }
	if (ez->err <= ez->thre) {
		hao_wl_t *p = wl_pushp(z); size_t pi = z->wn - 1;
		p->x_start = (int32_t)qs; p->x_end = (int32_t)qe;
		p->y_start = (int32_t)(t_s + ez->ps); p->y_end = (int32_t)(t_s + ez->pe);
		p->error = (int16_t)ez->err; p->cidx = p->clen = 0;
		if (is_cigar) {
			push_wcigar_(p, z, ez);
			if ((ez->pe + 1 == tl || ez->ps == 0) && ez->err > 0) {
				if (recal_boundary_(b, q_string, ql, tl, thres, t_s, ez->ps, ez->pe, ez->err, z->y_id, z->rev, &t_s, &aux_beg, &aux_end)) {
					p = &z->w[pi];
					z->cn = p->cidx; push_wcigar_(p, z, ez);
					p->y_start = (int32_t)(t_s + ez->ps); p->y_end = (int32_t)(t_s + ez->pe); p->error = (int16_t)ez->err;
				}
			}
		}
		p = &z->w[pi];
		p->extra_begin = (int16_t)aux_beg; p->extra_end = (int16_t)aux_end; p->error_threshold = (int16_t)thres;
		z->align_length += ql;
		return 1;
	}
	return 0;
}

static uint32_t gen_backtrace_adv_(ecb_t *b, ecz_t *z, hao_wl_t *p)
{ /* gen_backtrace_adv_exz, Correct.cpp:12563-12641 (rref branch).  p may live outside z->w (the pending
     window of push_hc_wlst_exz) or inside it: the caller passes an index-stable pointer */
	if (p->error < 0 || p->y_end < 0) return 0;
	int64_t qs = p->x_start, qe = p->x_end, ql = qe + 1 - qs, tl, thres = p->error_threshold, aln_l = ql + (thres << 1), t_pri_l, ts = p->y_start;
	int64_t aux_beg = p->extra_begin, aux_end = p->extra_end, t_tot_l;
	const char *q_string = b->qstr + qs; exz_t *ez = &b->ez;
	if (aux_end >= 0) t_pri_l = aln_l - aux_beg - aux_end;
	else {
		t_tot_l = (int64_t)b->r->len[z->y_id];
		t_pri_l = ts + aln_l - aux_beg; if (t_pri_l > t_tot_l) t_pri_l = t_tot_l;
		t_pri_l -= ts;
	}
	tl = t_pri_l;
	ec_tbuf(b, t_pri_l);
	hao_decode_sub(b->r, z->y_id, ts, t_pri_l, (int)z->rev, b->tstr);
	ez->ts = 0; ez->te = p->x_end - p->x_start; ez->tl = (int32_t)ql;
	ez->ps = -1; ez->pe = p->y_end - p->y_start; ez->pl = (int32_t)tl;
	ez->err = p->error; ez->thre = p->error_threshold;
	ez_semi64_trace(b->tstr, (int32_t)tl, q_string, (int32_t)ql, (int32_t)thres, (int32_t)aux_beg, ez);
	if (ez->err <= ez->thre) {
		p->y_start = (int32_t)(ts + ez->ps); p->y_end = (int32_t)(ts + ez->pe); p->error = (int16_t)ez->err;
		push_wcigar_(p, z, ez);
		if ((ez->pe + 1 == tl || ez->ps == 0) && ez->err > 0) {
			if (recal_boundary_(b, q_string, ql, tl, thres, ts, ez->ps, ez->pe, ez->err, z->y_id, z->rev, &ts, &aux_beg, &aux_end)) {
				z->cn = p->cidx; push_wcigar_(p, z, ez);
				p->y_start = (int32_t)(ts + ez->ps); p->y_end = (int32_t)(ts + ez->pe); p->error = (int16_t)ez->err;
			}
		}
		p->extra_begin = (int16_t)aux_beg; p->extra_end = (int16_t)aux_end;
		return 1;
	}
	p->error = -1;
	return 0;
}

static uint32_t push_hc_wlst_(ecb_t *b, ecz_t *z, uint32_t max_err, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t tl,
                              int64_t aux_beg, int64_t aux_end, double e_rate, int64_t bs, double ovlp_cut)
{ /* push_hc_wlst_exz, Correct.cpp:12776-12835 (force_aln = 0) */
	hao_wl_t p, t, *a; int64_t w_e, w_s, ce = qs - 1, cs = z->x_pos_s, toff, ovl, ualn, aln, ys; size_t a_n, k;
	p.x_start = (int32_t)qs; p.x_end = (int32_t)qe; p.y_start = (int32_t)ts; p.y_end = (int32_t)te; p.error = (int16_t)b->ez.err;
	p.extra_begin = (int16_t)aux_beg; p.extra_end = (int16_t)aux_end; p.error_threshold = (int16_t)b->ez.thre; p.cidx = p.clen = 0;
	if (z->wn > 0) { /* forward from the end of the previous window */
		w_e = z->w[z->wn - 1].x_end; toff = z->w[z->wn - 1].y_end + 1;
		while (w_e < ce && toff < tl) {
			w_s = w_e + 1; win_by_s(z, w_s, bs, &w_e);
			if (aln_wlst_adv_(b, z, max_err, w_s, w_e, toff, bs, e_rate, 0)) toff = z->w[z->wn - 1].y_end + 1;
			else break;
		}
		cs = z->w[z->wn - 1].x_end + 1;
	}
	a_n = z->wn; w_s = qs;
	if (w_s > cs) { /* backward from the start of this window */
		gen_backtrace_adv_(b, z, &p);
		toff = p.y_start - 1;
		while (w_s > cs) {
			w_e = w_s - 1; win_by_e(z, w_e, bs, &w_s); ys = toff + 1 - (w_e + 1 - w_s);
			if (ys >= 0 && aln_wlst_adv_(b, z, max_err, w_s, w_e, ys, bs, e_rate, 1)) toff = z->w[z->wn - 1].y_start - 1;
			else break;
		}
	}
	z->align_length += qe + 1 - qs;
	ovl = z->x_pos_e + 1 - z->x_pos_s; ualn = (qe + 1 - z->x_pos_s) - z->align_length; aln = ovl - ualn;
	if (!(aln > 0 && (double)ovl * ovlp_cut <= (double)aln)) { *wl_pushp(z) = p; return 0; }
	if (z->wn > a_n) {
		a = z->w + a_n; a_n = z->wn - a_n; toff = (int64_t)a_n; a_n >>= 1;
		for (k = 0; k < a_n; k++) { t = a[k]; a[k] = a[toff - 1 - k]; a[toff - 1 - k] = t; }
	}
	*wl_pushp(z) = p;
	return 1;
}

static uint32_t align_hc_ed_post_(ecb_t *b, ecz_t *z, double e_rate, int64_t w_l, double ovlp_cut)
{ /* align_hc_ed_post_extz, Correct.cpp:12951-13011 (force_aln = 0) */
	int64_t q_s, q_e, nw, k, q_l, t_tot_l, aux_beg, aux_end, t_s, thre, aln_l, t_pri_l, n_s, nl;
	z->wn = 0; z->cn = 0; z->align_length = 0;
	n_s = (z->x_pos_s / w_l) * w_l; nl = (z->x_pos_e + 1) - n_s; nw = nl / w_l + (nl % w_l > 0 ? 1 : 0);
	q_s = n_s < z->x_pos_s ? z->x_pos_s : n_s; q_e = n_s + w_l - 1 > z->x_pos_e ? z->x_pos_e : n_s + w_l - 1;
	for (k = 0; k < nw; k++) {
		aux_beg = aux_end = 0; q_l = 1 + q_e - q_s;
		thre = (int64_t)(q_l * e_rate); thre = adj_thre(thre, q_l);
		if (thre > EC_THRE_MAX) thre = EC_THRE_MAX;
		t_s = (q_s - z->x_pos_s) + z->y_pos_s;
		t_s += y_start_off(q_s, z->fc, z->fc_n, &b->bad);
		aln_l = q_l + (thre << 1); t_tot_l = (int64_t)b->r->len[z->y_id];
		if (init_waln_(thre, t_s, t_tot_l, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) {
			ec_tbuf(b, t_pri_l);
			hao_decode_sub(b->r, z->y_id, t_s, t_pri_l, (int)z->rev, b->tstr);
			ez_semi64(b->tstr, (int32_t)t_pri_l, b->qstr + q_s, (int32_t)q_l, (int32_t)thre, (int32_t)aux_beg, &b->ez);
			if (b->ez.err <= b->ez.thre) {
				if (!push_hc_wlst_(b, z, EC_THRE_MAX, q_s, q_e, t_s, t_s + b->ez.pe, t_tot_l, aux_beg, aux_end, e_rate, w_l, ovlp_cut)) return 0;
			}
		}
		q_s = q_e + 1; q_e = q_s + w_l - 1;
		if (q_e >= z->x_pos_e) q_e = z->x_pos_e;
	}
	{
		int64_t ovl = z->x_pos_e + 1 - z->x_pos_s;
		if (!(z->align_length > 0 && (double)ovl * ovlp_cut <= (double)z->align_length)) return 0;
	}
	return 1;
}

static uint32_t ed_cut_(ecb_t *b, uint32_t rev, uint32_t id, int64_t qs, int64_t qe, int64_t t_s, int64_t bs, double e_rate, int64_t max_err,
                        uint32_t aln_dir, int64_t *r_err, int64_t *aln_qlen)
{ /* ed_cut, Correct.cpp:13063-13104 (rref branch) */
	*aln_qlen = 0; *r_err = INT32_MAX;
	int64_t ql = qe + 1 - qs, aln_l, t_tot_l = (int64_t)b->r->len[id], aux_beg, aux_end, t_pri_l, thres; unsigned int error; int t_end, q_end;
	thres = dbl_err_thres((int)init_err_thres(ql, e_rate, bs, max_err), (int)ql);
	aln_l = ql + (thres << 1);
	if (!init_waln_(thres, t_s, t_tot_l, aln_l, &aux_beg, &aux_end, &t_s, &t_pri_l)) return 0;
	ec_tbuf(b, aln_l);
	hao_decode_sub(b->r, id, t_s, t_pri_l, (int)rev, b->tstr + aux_beg); /* fill_subregion, Correct.cpp:270-277 */
	memset(b->tstr, 'N', (size_t)aux_beg); memset(b->tstr + aux_beg + t_pri_l, 'N', (size_t)aux_end);
	bpm_extension(b->tstr, (int)aln_l, b->qstr + qs, (int)ql, (int)thres, aln_dir ? 1 : 0, &error, &t_end, &q_end);
	if (t_end != -1 && q_end != -1) *aln_qlen = aln_dir ? ql - q_end : q_end + 1;
	*r_err = error;
	if (*aln_qlen == 0) return 0;
	return 1;
}

static int64_t gen_extend_err_0_(ecb_t *b, ecz_t *z, int64_t bs, double e_rate, int64_t max_err, int64_t qs, int64_t qe, int64_t pk)
{ /* gen_extend_err_0_exz, Correct.cpp:13224-13283 */
	int64_t tot_e = 0, ts, di[2], al[2], tb[2], an = (int64_t)z->wn, ql = qe + 1 - qs; double rr;
	ts = (qs - z->x_pos_s) + z->y_pos_s; ts += y_start_off(qs, z->fc, z->fc_n, &b->bad);
	di[0] = di[1] = al[0] = al[1] = 0; tb[0] = tb[1] = -1;
	if (pk > 0 && qs == z->w[pk].x_end + 1) {
		if (z->w[pk].clen == 0) gen_backtrace_adv_(b, z, &z->w[pk]);
		tb[0] = z->w[pk].y_end + 1;
	}
	if (pk + 1 < an && qe + 1 == z->w[pk + 1].x_start) {
		if (z->w[pk + 1].clen == 0) gen_backtrace_adv_(b, z, &z->w[pk + 1]);
		tb[1] = z->w[pk + 1].y_start - ql;
	}
	if (tb[0] == -1 && tb[1] == -1) tb[0] = tb[1] = ts;
	else if (tb[0] == -1 && tb[1] != -1) tb[0] = tb[1];
	else if (tb[1] == -1 && tb[0] != -1) tb[1] = tb[0];
	if (tb[0] != -1) { if (!ed_cut_(b, z->rev, z->y_id, qs, qe, tb[0], bs, e_rate, max_err, 0, &di[0], &al[0])) { di[0] = ql; al[0] = 0; } }
	if (tb[1] != -1) { if (!ed_cut_(b, z->rev, z->y_id, qs, qe, tb[1], bs, e_rate, max_err, 1, &di[1], &al[1])) { di[1] = ql; al[1] = 0; } }
	if (al[0] && al[1]) {
		if (al[0] + al[1] <= ql) tot_e += di[0] + di[1] + ql - (al[0] + al[1]);
		else { rr = (double)ql / (double)(al[0] + al[1]); tot_e = (int64_t)((double)tot_e + (double)(di[0] + di[1]) * rr); }
	} else if (!al[0] && !al[1]) tot_e += ql;
	else if (al[0]) tot_e += di[0] + (ql - al[0]);
	else if (al[1]) tot_e += di[1] + (ql - al[1]);
	return tot_e;
}

static double gen_extend_err_(ecb_t *b, ecz_t *z, int64_t bs, double e_rate, double e_max, int64_t max_err, int64_t *r_e)
{ /* gen_extend_err_exz, Correct.cpp:13400-13440 (sec_check = 0) */
	int64_t ovl = z->x_pos_e + 1 - z->x_pos_s, k, ce, an = (int64_t)z->wn, tot_l = 0, tot_e = 0, ws, we, ql;
	*r_e = INT64_MAX;
	for (k = an - 1, ce = z->x_pos_e; k >= 0; k--) {
		tot_l += z->w[k].x_end + 1 - z->w[k].x_start;
		tot_e += z->w[k].error;
		we = z->w[k].x_end;
		while (we < ce) {
			ws = we + 1; win_by_s(z, ws, bs, &we);
			ql = we + 1 - ws; tot_l += ql;
			tot_e += gen_extend_err_0_(b, z, bs, e_rate, max_err, ws, we, k);
			if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
		}
		ce = z->w[k].x_start - 1;
		if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
	}
	if (ce >= z->x_pos_s) {
		we = z->x_pos_s - 1;
		while (we < ce) {
			ws = we + 1; win_by_s(z, ws, bs, &we);
			ql = we + 1 - ws; tot_l += ql;
			tot_e += gen_extend_err_0_(b, z, bs, e_rate, max_err, ws, we, k);
			if (e_max > 0 && (double)tot_e > (double)ovl * e_max) return 1.7976931348623157e308;
		}
	}
	*r_e = tot_e;
	return (double)tot_e / (double)tot_l;
}

/* step A of gen_hc_r_alin for every chain of read rid (Correct.cpp:25636-25645).  st: 0 = window pass rejected
 * the overlap, 1 = aligned but rr > e_rate, 2 = accepted (re = estimated error count).  Window lists and their
 * cigar pools are returned concatenated (w_off/w_n, c_off/c_n per overlap). */
int hao_ec_align_A(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const uint64_t *fc, double e_rate, int64_t w_l,
                   hao_alnA_t **out, hao_wl_t **wl, uint64_t *n_wl, uint16_t **cig, uint64_t *n_cig)
{
	uint64_t ql = r->len[rid], nw = 0, mw = 0, nc = 0, mc = 0; uint32_t j; hao_wl_t *W = 0; uint16_t *C = 0;
	char *qs = MALLOC_N(char, ql + 1); ecb_t b; ecz_t z; hao_alnA_t *o = MALLOC_N(hao_alnA_t, n_ch);
	const double e_max = e_rate * 1.5;
	memset(&b, 0, sizeof(b)); memset(&z, 0, sizeof(z));
	hao_decode(r, rid, qs); b.r = r; b.qstr = qs;
	for (j = 0; j < n_ch; j++) {
		const hao_ovlp_t *c = &ch[j]; int64_t re = INT64_MAX; double rr = 1.7976931348623157e308; uint32_t ok;
		z.x_pos_s = c->x_pos_s; z.x_pos_e = c->x_pos_e; z.y_pos_s = c->y_pos_s; z.y_pos_e = c->y_pos_e; z.y_id = c->y_id; z.rev = c->y_pos_strand;
		z.fc = fc + c->fc_off; z.fc_n = c->fc_n;
		ok = align_hc_ed_post_(&b, &z, e_rate, w_l, EC_OVLP_CUT);
		if (ok) rr = gen_extend_err_(&b, &z, w_l, e_rate, e_max + 0.000001, EC_THRE_MAX, &re);
		o[j].st = !ok ? 0 : (rr > e_rate ? 1 : 2); o[j].align_length = (uint32_t)z.align_length; o[j].rr = rr; o[j].re = re;
		o[j].w_off = nw; o[j].w_n = z.wn; o[j].c_off = nc; o[j].c_n = z.cn;
		if (nw + z.wn > mw) { mw = (nw + z.wn) * 2 + 64; W = (hao_wl_t *)realloc(W, mw * sizeof(hao_wl_t)); }
		if (nc + z.cn > mc) { mc = (nc + z.cn) * 2 + 64; C = (uint16_t *)realloc(C, mc * 2); }
		if (z.wn) memcpy(W + nw, z.w, z.wn * sizeof(hao_wl_t));
		if (z.cn) memcpy(C + nc, z.c, z.cn * 2);
		nw += z.wn; nc += z.cn;
	}
	free(qs); free(b.tstr); free(b.ez.cig); free(b.ez.path); free(z.w); free(z.c);
	*out = o; *wl = W; *n_wl = nw; *cig = C; *n_cig = nc;
	return b.bad ? -1 : 0;
}

/* ================================================================================================== */
/* step B: base-level CIGAR of an accepted overlap (gen_hc_fast_cigar, Correct.cpp:25137 -> 17813)     */
/* ================================================================================================== */
#define EC_MAX_SIN_L 10000 /* Levenshtein_distance.h:757 */
#define EC_MAX_SIN_E 2047  /* Levenshtein_distance.h:756 */
#define EC_FORCE_SIN_L 512 /* Levenshtein_distance.h:758 */

static uint64_t lchain_refine_(const hao_hit_t *a, int64_t a_n, hao_hit_t *des, int64_t **t_io, int32_t **f_io, int64_t **p_io, int64_t *cap,
                               int64_t max_skip, int64_t max_iter, int64_t max_dis, int64_t long_gap)
{ /* lchain_refine, Hash_Table.cpp:2457-2541 */
	if (a_n <= 0) return 0;
	int64_t *p, *t, max_f, n_skip, st, max_j, sc, msc, msc_i, dq, dr, dd, i, j, cL = 0; int32_t *f;
	if (a_n > *cap) { *cap = a_n * 2 + 64; *t_io = (int64_t *)realloc(*t_io, *cap * 8); *p_io = (int64_t *)realloc(*p_io, *cap * 8); *f_io = (int32_t *)realloc(*f_io, *cap * 4); }
	t = *t_io; f = *f_io; p = *p_io; msc = msc_i = -1;
	for (i = 1, f[0] = 0, p[0] = -1, msc_i = a_n - 1; i < a_n; i++) {
		j = i - 1;
		dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
		dd = dr > dq ? dr - dq : dq - dr;
		if (dd <= long_gap || dq > max_dis) { p[i] = i - 1; f[i] = (int32_t)i; }
		else break;
	}
	if (i < a_n) {
		memset(t, 0, (size_t)a_n * sizeof(*t));
		f[0] = 0; p[0] = -1;
		for (i = 1, st = 0; i < a_n; ++i) {
			max_f = INT32_MIN; n_skip = 0; max_j = -1;
			if (i - st > max_iter) st = i - max_iter;
			j = i - 1;
			dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
			dd = dr > dq ? dr - dq : dq - dr;
			if (dd <= long_gap) dd = 0;
			sc = f[j] - dd;
			if (sc > max_f) { max_f = sc; max_j = j; }
			if (p[j] >= 0) t[p[j]] = i;
			for (--j; j >= st && (int64_t)a[i].self_offset <= max_dis + (int64_t)a[j].self_offset; --j) {
				dq = (int64_t)a[i].self_offset - (int64_t)a[j].self_offset; dr = (int64_t)a[i].offset - (int64_t)a[j].offset;
				dd = dr > dq ? dr - dq : dq - dr;
				if (dd <= long_gap) dd = 0;
				sc = f[j] - dd;
				if (sc > max_f) { max_f = sc; max_j = j; if (n_skip > 0) --n_skip; }
				else if (t[j] == (int32_t)i) { if (++n_skip > max_skip) break; }
				if (p[j] >= 0) t[p[j]] = i;
			}
			f[i] = (int32_t)max_f; p[i] = max_j;
		}
		i = a_n - 1; msc = f[i]; msc_i = i;
		for (j = i - 1; j >= 0 && (int64_t)a[i].self_offset <= max_dis + (int64_t)a[j].self_offset; --j)
			if (msc < f[j] && p[j] >= 0) { msc = f[j]; msc_i = j; }
	}
	i = msc_i; cL = 0;
	while (i >= 0) { t[cL++] = i; i = p[i]; }
	n_skip = cL >> 1;
	for (i = 0; i < n_skip; i++) { msc_i = t[i]; t[i] = t[cL - i - 1]; t[cL - i - 1] = msc_i; }
	if (des) { /* a may alias des: t[] is ascending, so des[i] = a[t[i]] with t[i] >= i never reads an overwritten slot */
		for (i = 0; i < cL; i++) des[i] = a[t[i]];
	}
	return (uint64_t)cL;
}

/* ---- multi-word banded Myers with traceback: ed_band_cal_{global,extension_0,extension_1,semi..absent_diag}_infi_w_trace
 * (Levenshtein_distance.h:2516, 2694, 2823, 3020; the _64 / _128 variants 1580-2133, 3370-3856 are the same algorithm for one
 * and two words).  Band = 2*thre+1 bits in nword 64-bit words. */
typedef struct { uint64_t *Peq[5], *VP, *VN, *X, *D0, *HN, *HP; int32_t nword, mword; } mwv_t;
typedef struct {
	int32_t ps, pe, pl, ts, te, tl, thre, err, nword;
	uint16_t *cig; size_t cn, cm;
	uint64_t *path; size_t pn, pm;
	mwv_t v;
} mez_t;

static void mw_reserve(mez_t *ez, int32_t nword)
{
	int k;
	if (nword > ez->v.mword) {
		ez->v.mword = nword + 4;
		for (k = 0; k < 5; k++) ez->v.Peq[k] = (uint64_t *)realloc(ez->v.Peq[k], (size_t)ez->v.mword * 8);
		ez->v.VP = (uint64_t *)realloc(ez->v.VP, (size_t)ez->v.mword * 8); ez->v.VN = (uint64_t *)realloc(ez->v.VN, (size_t)ez->v.mword * 8);
		ez->v.X = (uint64_t *)realloc(ez->v.X, (size_t)ez->v.mword * 8); ez->v.D0 = (uint64_t *)realloc(ez->v.D0, (size_t)ez->v.mword * 8);
		ez->v.HN = (uint64_t *)realloc(ez->v.HN, (size_t)ez->v.mword * 8); ez->v.HP = (uint64_t *)realloc(ez->v.HP, (size_t)ez->v.mword * 8);
	}
	ez->v.nword = nword;
}
static inline void mw_set_lsub(uint64_t *x, int32_t l, int32_t nw)
{ /* w_infi_set_bit_lsub, Levenshtein_distance.h:2136 */
	memset(x, 0, (size_t)nw * 8);
	if (l >> 6) memset(x, -1, (size_t)(l >> 6) * 8);
	if (l & 63) x[l >> 6] = (1ULL << (l & 63)) - 1;
}
static inline int mw_bit(const uint64_t *x, int32_t b) { return (int)((x[b >> 6] >> (b & 63)) & 1ULL); }
static inline void mw_core(mwv_t *v, int c)
{ /* ed_infi_core, Levenshtein_distance.h:2148-2174 */
	int32_t w, nw = v->nword; uint64_t ad = 0;
	for (w = 0; w < nw; w++) {
		v->X[w] = v->Peq[c][w] | v->VN[w];
		v->D0[w] = v->X[w] & v->VP[w];
		v->D0[w] += ad; ad = v->D0[w] < ad; v->D0[w] += v->VP[w]; ad |= v->D0[w] < v->VP[w];
		v->D0[w] ^= v->VP[w];
		v->D0[w] |= v->X[w];
		v->HN[w] = v->VP[w] & v->D0[w];
		v->HP[w] = ~(v->VP[w] | v->D0[w]);
		v->HP[w] |= v->VN[w];
	}
	for (w = nw - 1, ad = 0; w >= 0; w--) {
		v->X[w] = (v->D0[w] >> 1) | ad; ad = v->D0[w] << 63;
		v->VN[w] = v->X[w] & v->HP[w];
		v->VP[w] = ~(v->X[w] | v->HP[w]);
		v->VP[w] |= v->HN[w];
	}
}
static inline void mw_post_peq(mwv_t *v)
{ /* ed_infi_post_Peq, Levenshtein_distance.h:2176-2189 */
	int32_t w, nw = v->nword, k;
	for (k = 0; k < 4; k++) {
		for (w = 0; w + 1 < nw; w++) v->Peq[k][w] = (v->Peq[k][w] >> 1) | (v->Peq[k][w + 1] << 63);
		v->Peq[k][nw - 1] >>= 1;
	}
}
static inline void mw_path_push(mez_t *ez)
{
	size_t nw = (size_t)ez->v.nword;
	memcpy(ez->path + ez->pn, ez->v.D0, nw * 8); ez->pn += nw; memcpy(ez->path + ez->pn, ez->v.VP, nw * 8); ez->pn += nw;
	memcpy(ez->path + ez->pn, ez->v.VN, nw * 8); ez->pn += nw; memcpy(ez->path + ez->pn, ez->v.HP, nw * 8); ez->pn += nw;
	memcpy(ez->path + ez->pn, ez->v.HN, nw * 8); ez->pn += nw;
}
static void mez_push_trace(mez_t *ez, uint16_t c, uint32_t len)
{
	c <<= 14;
	while (len >= 0x3fff) { cig_push(&ez->cig, &ez->cn, &ez->cm, (uint16_t)(c + 0x3fff)); len -= 0x3fff; }
	if (len) cig_push(&ez->cig, &ez->cn, &ez->cm, (uint16_t)(c + len));
}

static void mw_gen_trace(mez_t *ez, int32_t ptrim, int reverse)
{ /* gen_trace, Levenshtein_distance.h:903-985 */
	if (ez->err > ez->thre) return;
	ez->cn = 0;
	int32_t V, H, D, min, cur, tn = ez->te + 1 - ez->ts, pn = tn + (ez->thre << 1), bd = (ez->thre << 1) + 1;
	int32_t bs = (int32_t)(ez->pn / (size_t)tn), bbs = bs / 5, poff = ez->pe, sft = bd - (pn - ez->pe - ptrim);
	int32_t i = tn, low = bd - 1, d = 0, pd = -1, pdn = 0; cur = ez->err;
	while (i > 0 && cur > 0) {
		const uint64_t *D0 = ez->path + (size_t)(i - 1) * bs, *VP = D0 + bbs, *VN = VP + bbs, *HP = VN + bbs, *HN = HP + bbs;
		D = cur - (1 - mw_bit(D0, sft)); d = 0; min = D;
		H = V = INT32_MAX;
		if (sft != low) { H = cur + mw_bit(HN, sft) - mw_bit(HP, sft); if (H + 1 == cur && H <= min) { min = H; d = 3; } }
		if (sft != 0) { V = cur + mw_bit(VN, sft - 1) - mw_bit(VP, sft - 1); if (V + 1 == cur && V <= min) { min = V; d = 2; } }
		if (d == 0) { if (D != cur) d = 1; i--; poff--; }
		else if (d == 2) { sft--; poff--; }
		else if (d == 3) { i--; sft++; }
		if (d == pd) pdn++;
		else { if (pdn > 0) mez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = 1; }
		cur = min;
	}
	if (i > 0) {
		d = 0; poff -= i;
		if (d == pd) pdn += i;
		else { if (pdn > 0) mez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	poff++;
	if (ez->ps < 0 || ez->ps >= ez->pl) ez->ps = poff;
	else if (poff > ez->ps) {
		d = 2; i = poff - ez->ps;
		if (d == pd) pdn += i;
		else { if (pdn > 0) mez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn); pd = d; pdn = i; }
	}
	if (pdn > 0) mez_push_trace(ez, (uint16_t)pd, (uint32_t)pdn);
	if (reverse) { size_t k, h = ez->cn >> 1; for (k = 0; k < h; k++) { uint16_t t = ez->cig[k]; ez->cig[k] = ez->cig[ez->cn - k - 1]; ez->cig[ez->cn - k - 1] = t; } }
}

/* mode: 0 global, 1 forward extension, 2 backward extension, 3 semi-global (absent diagonals = aux_beg).  ez->err must be
 * INT32_MAX on entry (cal_exz_infi_adv clears it, Correct.cpp:15621). */
static void mw_align(int mode, const char *pstr, int32_t pn, const char *tstr, int32_t tn, int32_t thre, int32_t abs_diag, mez_t *ez)
{
	int32_t bd = (thre << 1) + 1, nword = (bd >> 6) + !!(bd & 63), i, err, tn0, cut = thre + (thre << 1), i_bd, c, Peq_i; uint64_t Peq_m;
	int32_t pidx = 0, tidx = 0, pe, tmp_e = INT32_MAX, k, poff; mwv_t *v;
	ez->cn = 0;
	/* init_base_ed + per-mode start state */
	ez->thre = thre; ez->err = INT32_MAX; ez->pl = pn; ez->tl = tn;
	if (mode == 0) { ez->ps = ez->ts = 0; if (pn > tn + thre || tn > pn + thre) return; }
	else if (mode == 1) { ez->ps = ez->ts = 0; ez->pe = ez->te = -1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; }
	else if (mode == 2) { ez->ps = ez->ts = INT32_MAX; ez->pe = pn - 1; ez->te = tn - 1; if (pn > tn + thre) pn = tn + thre; else if (tn > pn + thre) tn = pn + thre; pidx = ez->pe; tidx = ez->te; }
	else { ez->ps = ez->pe = -1; ez->ts = 0; ez->te = tn - 1; if (pn > tn + cut || tn > pn + cut) return; }
	tn0 = tn - 1; pe = pn - 1;
	ez->nword = nword; mw_reserve(ez, nword); v = &ez->v;
	for (k = 0; k < 5; k++) memset(v->Peq[k], 0, (size_t)nword * 8);
#define PCH(j) nt4(mode == 2 ? pstr[pidx - (j)] : pstr[(j)])
#define TCH(j) nt4(mode == 2 ? tstr[tidx - (j)] : tstr[(j)])
	if (mode == 3) {
		memset(v->VP, 0, (size_t)nword * 8); mw_set_lsub(v->VN, abs_diag, nword);
		bd = ((thre << 1) + 1) - abs_diag; bd = bd <= pn ? bd : pn; i_bd = abs_diag;
		for (i = 0; i < bd; i++, i_bd++) { c = PCH(i); v->Peq[c][i_bd >> 6] |= 1ULL << (i_bd & 63); }
		i_bd = (thre << 1) - abs_diag; err = abs_diag;
	} else {
		bd = thre + 1; bd = bd <= pn ? bd : pn; i_bd = thre;
		for (i = 0; i < bd; i++, i_bd++) { c = PCH(i); v->Peq[c][i_bd >> 6] |= 1ULL << (i_bd & 63); }
		i_bd = thre; err = thre;
		mw_set_lsub(v->VN, thre, nword); mw_set_lsub(v->VP, (thre << 1) + 1, nword);
		for (k = 0; k < nword; k++) v->VP[k] ^= v->VN[k];
	}
	memset(v->Peq[4], 0, (size_t)nword * 8);
	if ((size_t)nword * tn * 5 > ez->pm) { ez->pm = (size_t)nword * tn * 5 + 256; ez->path = (uint64_t *)realloc(ez->path, ez->pm * 8); }
	ez->pn = 0;
	Peq_i = (thre << 1) >> 6; Peq_m = 1ULL << ((thre << 1) & 63);
	for (i = 0; i <= tn0; i++) {
		mw_core(v, TCH(i));
		if (!(v->D0[0] & 1ULL)) { ++err; if (err > cut) return; }
		if (i < tn0) {
			if (mode == 1 || mode == 2) { /* running best end point of an extension (2758-2779 / 2889-2910) */
				poff = i - thre; k = i + thre - pe;
				if (k >= 0) {
					if (tmp_e == INT32_MAX) {
						tmp_e = err;
						for (k = 0; poff < pe; poff++, k++) { tmp_e += mw_bit(v->VP, k); tmp_e -= mw_bit(v->VN, k); }
					} else {
						k = (thre << 1) - k;
						if (k >= 0) { tmp_e += mw_bit(v->HP, k); tmp_e -= mw_bit(v->HN, k); }
					}
					if (tmp_e <= ez->thre && tmp_e < ez->err) {
						ez->err = tmp_e;
						if (mode == 1) { ez->pe = pe; ez->te = i; } else { ez->ps = pidx - pe; ez->ts = tidx - i; }
					}
				}
			}
			mw_post_peq(v);
			++i_bd; c = 4;
			if (i_bd < pn) c = PCH(i_bd);
			if (c < 4) v->Peq[c][Peq_i] |= Peq_m;
		}
		mw_path_push(ez);
	}
#undef PCH
#undef TCH
	if (mode == 0) {
		int32_t site = tn - 1 - thre, ct = pn - 1;
		for (i = 0; site < ct; site++, i++) { err += mw_bit(v->VP, i); err -= mw_bit(v->VN, i); }
		if (site == ct && err <= thre) { ez->err = err; ez->pe = pn - 1; ez->te = tn - 1; }
		mw_gen_trace(ez, thre, 1);
	} else if (mode == 1 || mode == 2) {
		int32_t site = tn - 1 - thre, ct = pn - 1;
		for (i = 0; site < ct; i++) {
			err += mw_bit(v->VP, i); err -= mw_bit(v->VN, i); site++;
			if (err <= thre && err < ez->err) { ez->err = err; if (mode == 1) { ez->pe = site; ez->te = tn - 1; } else { ez->ps = pidx - site; ez->ts = tidx + 1 - tn; } }
		}
		if (err <= thre && err < ez->err) { ez->err = err; if (mode == 1) { ez->pe = site; ez->te = tn - 1; } else { ez->ps = pidx - site; ez->ts = tidx + 1 - tn; } }
		if (ez->te - ez->ts + 1 != tn) { ez->pn /= (size_t)tn; ez->pn *= (size_t)(ez->te + 1 - ez->ts); }
		if (mode == 1) mw_gen_trace(ez, thre, 1);
		else {
			poff = ez->ps; ez->ps = pidx - ez->pe; ez->pe = pidx - poff;
			mw_gen_trace(ez, thre, 0);
			poff = ez->ps; ez->ps = pidx - ez->pe; ez->pe = pidx - poff;
		}
	} else {
		int32_t site = tn - 1 - abs_diag, ai = pn - tn + abs_diag, uge = INT32_MAX;
		for (i = 0; site < 0 && i < ai; i++, site++) { err += mw_bit(v->VP, i); err -= mw_bit(v->VN, i); }
		if (err <= thre && err <= ez->err) { ez->err = err; ez->pe = site; }
		site -= i;
		while (i < ai) {
			err += mw_bit(v->VP, i); err -= mw_bit(v->VN, i); ++i;
			if (err <= thre && err <= ez->err) { ez->err = err; ez->pe = site + i; }
			if (i == thre) uge = err;
		}
		if (uge <= thre && uge == ez->err) ez->pe = site + thre;
		mw_gen_trace(ez, abs_diag, 1);
	}
}

/* ---- per-overlap driver of step B ------------------------------------------------------------------ */
typedef struct { hao_wl_t *w; size_t wn, wm; uint16_t *c; size_t cn, cm; } wlv_t; /* window_list_alloc */
typedef struct {
	const hao_reads_t *r; const char *qstr; int64_t ql; char *tstr; size_t tm; mez_t ez; int bad;
	int64_t *dp_t, *dp_p; int32_t *dp_f; int64_t dp_cap;
} ecB_t;

static hao_wl_t *wlv_pushp(wlv_t *v)
{
	if (v->wn == v->wm) { v->wm = v->wm ? v->wm << 1 : 16; v->w = (hao_wl_t *)realloc(v->w, v->wm * sizeof(hao_wl_t)); }
	return &v->w[v->wn++];
}
static void wlv_push_trace(wlv_t *v, uint16_t c, uint32_t len)
{
	c <<= 14;
	while (len >= 0x3fff) { cig_push(&v->c, &v->cn, &v->cm, (uint16_t)(c + 0x3fff)); len -= 0x3fff; }
	if (len) cig_push(&v->c, &v->cn, &v->cm, (uint16_t)(c + len));
}
static void push_alnw_(wlv_t *aux, const mez_t *ez)
{ /* push_alnw + append_wcigar, Correct.cpp:15988-16019, 15954-15986 */
	hao_wl_t *p; size_t k;
	if (aux->wn > 0) {
		p = &aux->w[aux->wn - 1];
		if (p->clen > 0) {
			int64_t t = (int64_t)p->error + (int64_t)ez->err;
			if (p->x_end + 1 == ez->ts && p->y_end + 1 == ez->ps && t < INT16_MAX) {
				p->x_end = ez->te; p->y_end = ez->pe; p->error = (int16_t)(p->error + ez->err);
				if (ez->cn > 0) { /* append_wcigar: merge the first run with the pool's last one when the op is the same */
					uint16_t c0 = aux->c[aux->cn - 1] >> 14, c = ez->cig[0] >> 14; uint32_t l0 = aux->c[aux->cn - 1] & 0x3fff, l = ez->cig[0] & 0x3fff; size_t ci = 1;
					for (; ci < ez->cn && (ez->cig[ci] >> 14) == c; ci++) l += ez->cig[ci] & 0x3fff; /* pop_trace */
					if (c0 == c) { l += l0; aux->cn--; }
					wlv_push_trace(aux, c, l);
					for (k = ci; k < ez->cn; k++) cig_push(&aux->c, &aux->cn, &aux->cm, ez->cig[k]);
					p->clen = (uint32_t)(aux->cn - p->cidx);
				}
				return;
			}
		}
	}
	p = wlv_pushp(aux);
	p->x_start = ez->ts; p->x_end = ez->te; p->y_start = ez->ps; p->y_end = ez->pe;
	p->extra_begin = p->extra_end = 0; /* not set by the reference (kv_pushp leaves the slot as it is); never read while clen > 0 */
	p->error_threshold = 0; p->error = (int16_t)ez->err;
	p->cidx = (uint32_t)aux->cn; p->clen = (uint32_t)ez->cn;
	for (k = 0; k < ez->cn; k++) cig_push(&aux->c, &aux->cn, &aux->cm, ez->cig[k]);
}
static void push_unmap_alnw_(wlv_t *aux, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode)
{ /* push_unmap_alnw, Correct.cpp:16021-16030 */
	hao_wl_t *p = wlv_pushp(aux);
	p->x_start = (int32_t)qs; p->x_end = (int32_t)qe; p->y_start = (int32_t)ts; p->y_end = (int32_t)te;
	p->error_threshold = (int16_t)mode; p->error = INT16_MAX; p->extra_begin = p->extra_end = -1; p->cidx = p->clen = 0;
}
static void set_exact_(mez_t *ez, int64_t qs, int64_t qe, int64_t ts, int64_t te)
{ /* set_exact_exz, Correct.cpp:16167-16175 */
	ez->thre = 0; ez->cn = 0; ez->err = 0; mez_push_trace(ez, 0, (uint32_t)(qe - qs));
	ez->pl = (int32_t)(te - ts); ez->ps = (int32_t)ts; ez->pe = (int32_t)(ts + ez->pl - 1);
	ez->tl = (int32_t)(qe - qs); ez->ts = (int32_t)qs; ez->te = (int32_t)(qs + ez->tl - 1);
}
static inline int64_t scale_ed_thre_(uint32_t err, uint32_t max_err)
{ /* scale_ed_thre, Correct.cpp:14354-14360 */
	uint64_t bd = ((uint64_t)err << 1) + 1, w = (bd >> 6) << 6; if (w < bd) w += 64;
	err = (uint32_t)((w - 1) >> 1); if (err > max_err) err = max_err;
	return err;
}
static void adjust_ext_offset_(int64_t *qs, int64_t *qe, int64_t *ts, int64_t *te, int64_t ql, int64_t tl, int64_t thre, int64_t mode)
{ /* adjust_ext_offset, Correct.cpp:14400-14422 */
	int64_t qoff, toff;
	if (mode == 1) { qoff = ql - *qs; toff = tl - *ts; if (qoff <= toff) { *qe = ql; *te = *ts + qoff + thre; } else { *te = tl; *qe = *qs + toff + thre; } }
	else if (mode == 2) { qoff = *qe; toff = *te; if (qoff <= toff) { *qs = 0; *ts = *te - qoff - thre; } else { *ts = 0; *qs = *qe - toff - thre; } }
	if (*qs < 0) *qs = 0;
	if (*ts < 0) *ts = 0;
	if (*qe > ql) *qe = ql;
	if (*te > tl) *te = tl;
}
static void ecB_tbuf(ecB_t *b, int64_t n) { if ((size_t)n + 8 > b->tm) { b->tm = (size_t)n + 64; b->tstr = (char *)realloc(b->tstr, b->tm); } }

static int64_t cal_estimate_err_hc_(const ecz_t *z, int64_t wl, int64_t qs, int64_t qe, int64_t ts, int64_t te, double e_rate, int64_t *exact)
{ /* cal_estimate_err_hc, Correct.cpp:15403-15455 */
	int64_t k, ws, we, wid, os, oe, ovlp, tot, cov_l, est = (int64_t)((qe - qs) * e_rate), wn = (int64_t)z->wn, exa = 1, ots, ote, q0, t0;
	*exact = 0;
	if (!wn) return est;
	if (qs < z->x_pos_s) qs = z->x_pos_s;
	if (qe > z->x_pos_e + 1) qe = z->x_pos_e + 1;
	ws = qs / wl; ws *= wl; wid = win_by_s(z, ws, wl, NULL);
	if (wid >= wn) wid = wn - 1;
	for (k = wid; k < wn && qs > z->w[k].x_end; k++);
	if (k == wn) return est;
	for (; k >= 0 && qs < z->w[k].x_start; k--);
	if (k < 0) k = 0;
	for (tot = cov_l = 0, ots = ote = -1; k < wn && z->w[k].x_start < qe; k++) {
		if (z->w[k].y_end == -1) continue;
		ws = z->w[k].x_start; we = (int64_t)z->w[k].x_end + 1;
		os = qs > ws ? qs : ws; oe = qe < we ? qe : we;
		ovlp = oe > os ? oe - os : 0;
		if (!ovlp) continue;
		cov_l += ovlp;
		if (ovlp == we - ws) tot += z->w[k].error;
		else tot = (int64_t)((double)tot + ((double)z->w[k].error) * ((double)ovlp) / ((double)(we - ws)));
		if (z->w[k].error > 0) exa = 0;
		if (exa) {
			q0 = os - ws;
			we = (int64_t)z->w[k].y_end + 1; ws = we - ((int64_t)z->w[k].x_end + 1 - z->w[k].x_start);
			os = ts > ws ? ts : ws; oe = te < we ? te : we;
			ovlp = oe > os ? oe - os : 0;
			t0 = os - ws;
			if (ovlp && q0 == t0) {
				if (ote == -1) { ots = os; ote = oe; }
				else if (ote == os) ote = oe;
				else exa = 0;
			} else exa = 0;
		}
	}
	tot = (int64_t)((double)tot + ((qe - qs) - cov_l) * e_rate);
	if (exa && (qe - qs) == cov_l) { if ((qe - qs) == (ote - ots) && ots == ts && ote == te) *exact = 1; }
	return tot;
}

static int64_t cal_exact_(ecB_t *b, const ecz_t *z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode)
{ /* cal_exact_exz, Correct.cpp:15725-15762 */
	mez_t *ez = &b->ez; int64_t ql = qe - qs, tl, t_tot_l = (int64_t)b->r->len[z->y_id];
	ez->err = INT32_MAX; ez->thre = 0; ez->cn = 0;
	if (mode == 3) { ts = (qs - z->x_pos_s) + z->y_pos_s; ts += y_start_off(qs, z->fc, z->fc_n, &b->bad); te = ts + ql; }
	else if (mode == 1) te = ts + ql;
	else if (mode == 2) ts = te - ql;
	if (ts < 0) ts = 0;
	if (ts > t_tot_l) ts = t_tot_l;
	if (te > t_tot_l) te = t_tot_l;
	ql = qe - qs; tl = te - ts;
	if (ql != tl) return 0;
	ecB_tbuf(b, tl); hao_decode_sub(b->r, z->y_id, ts, tl, (int)z->rev, b->tstr);
	if (memcmp(b->qstr + qs, b->tstr, (size_t)ql)) return 0;
	ez->err = 0; mez_push_trace(ez, 0, (uint32_t)ql);
	ez->pl = (int32_t)tl; ez->ps = (int32_t)ts; ez->pe = (int32_t)(ts + tl - 1);
	ez->tl = (int32_t)ql; ez->ts = (int32_t)qs; ez->te = (int32_t)(qs + ql - 1);
	return 1;
}

static int64_t cal_exz_adv_(ecB_t *b, const ecz_t *z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t thre, int64_t *pthre, int64_t mode)
{ /* cal_exz_infi_adv, Correct.cpp:15617-15666 */
	mez_t *ez = &b->ez; int64_t aux_beg = 0, ql = qe - qs, tl = te - ts, t_tot_l = (int64_t)b->r->len[z->y_id], dd = ql > tl ? ql : tl;
	ez->err = INT32_MAX;
	if (mode == 3) { /* update_semi_coord, Correct.cpp:14364-14381 */
		int64_t th = thre > dd ? dd : thre, aux_end, l, aln_l = (qe - qs) + (th << 1);
		ts = (qs - z->x_pos_s) + z->y_pos_s; ts += y_start_off(qs, z->fc, z->fc_n, &b->bad);
		if (!init_waln_(th, ts, t_tot_l, aln_l, &aux_beg, &aux_end, &ts, &l)) ts = te = aux_beg = -1;
		else te = ts + l;
	} else if (mode == 1 || mode == 2) adjust_ext_offset_(&qs, &qe, &ts, &te, b->ql, t_tot_l, thre > dd ? dd : thre, mode);
	if (qe > qs && te > ts && ts != -1 && te != -1) {
		ql = qe - qs; tl = te - ts; dd = ql > tl ? ql : tl;
		if (thre > dd) thre = dd;
		if (thre <= *pthre) return 0;
		*pthre = thre;
		ecB_tbuf(b, tl); hao_decode_sub(b->r, z->y_id, ts, tl, (int)z->rev, b->tstr);
		mw_align((int)mode, b->tstr, (int32_t)tl, b->qstr + qs, (int32_t)ql, (int32_t)thre, (int32_t)aux_beg, ez);
		if (ez->err <= ez->thre) { ez->ps += (int32_t)ts; ez->pe += (int32_t)ts; ez->ts += (int32_t)qs; ez->te += (int32_t)qs; return 1; }
		return 0;
	}
	return 0;
}

static int64_t hc_aln_adv_(ecB_t *b, const ecz_t *z, int64_t qs, int64_t qe, int64_t ts, int64_t te, int64_t mode, int64_t wl, double e_rate, wlv_t *aux)
{ /* hc_aln_exz_adv_hc, Correct.cpp:16178-16260 (maxl = MAX_SIN_L, maxe = MAX_SIN_E, force_l = FORCE_SIN_L, estimate_err = -1) */
	mez_t *ez = &b->ez; int64_t thre, ql = qe - qs, thre0, pthre = -1, full = 0, est;
	ez->err = INT32_MAX; ez->thre = 0;
	if (ts == -1 && te == -1) mode = 3;
	if (ql == 0 && te - ts == 0) return 1;
	if (ql <= 0 || te - ts <= 0) return 0;
	est = cal_estimate_err_hc_(z, wl, qs, qe, ts, te, e_rate, &full);
	if (est == 0) {
		if (full) { set_exact_(ez, qs, qe, ts, te); push_alnw_(aux, ez); return 1; }
		else if (cal_exact_(b, z, qs, qe, ts, te, mode)) { push_alnw_(aux, ez); return 1; }
	}
	if (ql <= EC_MAX_SIN_L && (est >> 1) <= EC_MAX_SIN_E) {
		thre = scale_ed_thre_((uint32_t)est, EC_MAX_SIN_E);
		if (cal_exz_adv_(b, z, qs, qe, ts, te, thre, &pthre, mode)) { push_alnw_(aux, ez); return 1; }
		thre0 = thre; thre = (int64_t)(ql * e_rate); thre = scale_ed_thre_((uint32_t)thre, EC_MAX_SIN_E);
		if (thre > thre0) { if (cal_exz_adv_(b, z, qs, qe, ts, te, thre, &pthre, mode)) { push_alnw_(aux, ez); return 1; } }
		thre0 = thre; thre <<= 1; thre = scale_ed_thre_((uint32_t)thre, EC_MAX_SIN_E);
		if (thre > thre0) { if (cal_exz_adv_(b, z, qs, qe, ts, te, thre, &pthre, mode)) { push_alnw_(aux, ez); return 1; } }
		thre0 = thre; thre = (int64_t)(ql * 0.51); thre = scale_ed_thre_((uint32_t)thre, EC_MAX_SIN_E);
		if (thre > thre0) { if (cal_exz_adv_(b, z, qs, qe, ts, te, thre, &pthre, mode)) { push_alnw_(aux, ez); return 1; } }
		if (ql <= EC_FORCE_SIN_L) { thre = EC_MAX_SIN_E; if (cal_exz_adv_(b, z, qs, qe, ts, te, thre, &pthre, mode)) { push_alnw_(aux, ez); return 1; } }
	}
	return 0;
}

static void hc_ovlp_base_direct_(ecB_t *b, const ecz_t *z, int64_t nh_err, const hao_hit_t *ch_a, int64_t ch_n, int64_t wl, double e_rate, int64_t tl, wlv_t *aux)
{ /* hc_ovlp_base_direct, Correct.cpp:17425-17509 (pre_mode = -1) */
	int64_t i, l, mode, q[2], t[2], qr, tr, zn, ql = b->ql;
	if (nh_err == 0 && z->wn) {
		zn = (int64_t)z->wn;
		for (i = 1; i < zn; i++) {
			if (z->w[i].error == 0 && z->w[i - 1].error == 0 && z->w[i].x_start == z->w[i - 1].x_end + 1 &&
			    z->w[i].y_end == z->w[i - 1].y_end + (z->w[i].x_end - z->w[i - 1].x_end)) continue;
			break;
		}
		if (i >= zn) {
			q[0] = z->w[0].x_start; q[1] = z->w[zn - 1].x_end;
			t[1] = z->w[zn - 1].y_end; t[0] = z->w[0].y_end - (z->w[0].x_end - z->w[0].x_start);
			if (q[0] <= t[0]) { t[0] -= q[0]; q[0] = 0; } else { q[0] -= t[0]; t[0] = 0; }
			qr = ql - q[1] - 1; tr = tl - t[1] - 1;
			if (qr <= tr) { q[1] = ql - 1; t[1] += qr; } else { t[1] = tl - 1; q[1] += tr; }
			if (q[0] == z->w[0].x_start && q[1] == z->w[zn - 1].x_end) { set_exact_(&b->ez, q[0], q[1] + 1, t[0], t[1] + 1); push_alnw_(aux, &b->ez); return; }
		}
	}
	for (l = -1, i = 0; i <= ch_n; i++) {
		q[0] = q[1] = t[0] = t[1] = mode = -1;
		if (l >= 0) { q[0] = ch_a[l].self_offset; t[0] = ch_a[l].offset; } else q[0] = 0;
		if (i < ch_n) { q[1] = ch_a[i].self_offset; t[1] = ch_a[i].offset; } else q[1] = ql;
		if (t[0] != -1 && t[1] != -1) mode = 0;
		else if (t[0] != -1 && t[1] == -1) mode = 1;
		else if (t[0] == -1 && t[1] != -1) mode = 2;
		else mode = 3;
		if (mode == 1 || mode == 2) adjust_ext_offset_(&q[0], &q[1], &t[0], &t[1], ql, tl, 0, mode);
		if (!hc_aln_adv_(b, z, q[0], q[1], t[0], t[1], mode, wl, e_rate, aux)) push_unmap_alnw_(aux, q[0], q[1] - 1, t[0], t[1] - 1, mode);
		l = i;
	}
}

typedef struct { int32_t st; int32_t need_rechain; int64_t re; uint32_t x_pos_s, x_pos_e, y_pos_s, y_pos_e; uint64_t w_off, w_n, c_off, c_n; } hao_alnB_t;

/* step B for the accepted overlaps of read rid.  a[] / wlA / (cigars unused) = step A's output; hits = the compacted chain anchors
 * (cl->list after h_ec_lchain; ch[j].non_homopolymer_errors = first anchor of chain j, modified in place by return_t_chain).
 * need_rechain = 1 marks overlaps whose base alignment leaves a window >= FORCE_SIN_L unaligned: the reference re-seeds and
 * re-chains those (rechain_aln_hc, Correct.cpp:17669) — not restated yet, their result is not comparable. */
int hao_ec_align_B(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const uint64_t *fc, hao_hit_t *hits, uint64_t n_hits,
                   const hao_alnA_t *a, const hao_wl_t *wlA, double e_rate, int64_t w_l,
                   hao_alnB_t **out, hao_wl_t **wl, uint64_t *n_wl, uint16_t **cig, uint64_t *n_cig)
{
	uint64_t ql = r->len[rid], nw = 0, mw = 0, nc = 0, mc = 0; uint32_t j; hao_wl_t *W = 0; uint16_t *C = 0;
	char *qs = MALLOC_N(char, ql + 1); ecB_t b; ecz_t z; wlv_t aux; hao_alnB_t *o = CALLOC_N(hao_alnB_t, n_ch);
	memset(&b, 0, sizeof(b)); memset(&z, 0, sizeof(z)); memset(&aux, 0, sizeof(aux));
	hao_decode(r, rid, qs); b.r = r; b.qstr = qs; b.ql = (int64_t)ql;
	for (j = 0; j < n_ch; j++) {
		const hao_ovlp_t *c = &ch[j]; int64_t i, scn, cn, tl = (int64_t)r->len[c->y_id], tot_e = 0; hao_hit_t *ca; uint32_t pid;
		o[j].st = a[j].st; o[j].w_off = nw; o[j].c_off = nc;
		if (a[j].st != 2) continue;
		z.x_pos_s = c->x_pos_s; z.x_pos_e = c->x_pos_e; z.y_pos_s = c->y_pos_s; z.y_pos_e = c->y_pos_e; z.y_id = c->y_id; z.rev = c->y_pos_strand;
		z.fc = fc + c->fc_off; z.fc_n = c->fc_n; z.w = (hao_wl_t *)(wlA + a[j].w_off); z.wn = a[j].w_n;
		/* return_t_chain, Correct.cpp:22997-23020 */
		i = c->non_homopolymer_errors; pid = hits[i].id_strand & 0x7fffffffu;
		for (; i < (int64_t)n_hits && (hits[i].id_strand & 0x7fffffffu) == pid && pid != 0x7fffffffu; i++);
		scn = i - c->non_homopolymer_errors; ca = hits + c->non_homopolymer_errors;
		cn = (int64_t)lchain_refine_(ca, scn, ca, &b.dp_t, &b.dp_f, &b.dp_p, &b.dp_cap, 50, 5000, 512, 16);
		for (i = cn; i < scn; i++) ca[i].id_strand = (ca[i].id_strand & 0x80000000u) | 0x7fffffffu;
		/* gen_hc_fast_cigar0, Correct.cpp:17813-17870 */
		aux.wn = aux.cn = 0;
		if (cn > 0) {
			hc_ovlp_base_direct_(&b, &z, a[j].re, ca, cn, w_l, e_rate, tl, &aux);
			for (i = 0; i < (int64_t)aux.wn; i++) {
				const hao_wl_t *u = &aux.w[i];
				if (!(u->error == INT16_MAX && u->clen == 0 && u->extra_end < 0)) continue;
				if ((int64_t)u->x_end + 1 - u->x_start >= EC_FORCE_SIN_L && (int64_t)u->y_end + 1 - u->y_start >= EC_FORCE_SIN_L) o[j].need_rechain = 1;
			}
			/* update_overlap_region, Correct.cpp:17249-17275 */
			if (aux.wn) { z.x_pos_s = aux.w[0].x_start; z.x_pos_e = aux.w[aux.wn - 1].x_end; z.y_pos_s = aux.w[0].y_start; z.y_pos_e = aux.w[aux.wn - 1].y_end; }
			{
				int64_t xr, yr;
				if (z.x_pos_s <= z.y_pos_s) { z.y_pos_s -= z.x_pos_s; z.x_pos_s = 0; } else { z.x_pos_s -= z.y_pos_s; z.y_pos_s = 0; }
				xr = (int64_t)ql - z.x_pos_e - 1; yr = tl - z.y_pos_e - 1;
				if (xr <= yr) { z.x_pos_e = (int64_t)ql - 1; z.y_pos_e += xr; } else { z.y_pos_e = tl - 1; z.x_pos_e += yr; }
			}
			for (i = 0; i < (int64_t)aux.wn; i++) {
				const hao_wl_t *u = &aux.w[i];
				if (u->error == INT16_MAX && u->clen == 0 && u->extra_end < 0) { int64_t xe = (int64_t)u->x_end + 1 - u->x_start, ye = (int64_t)u->y_end + 1 - u->y_start; tot_e += xe >= ye ? xe : ye; }
				else tot_e += u->error;
			}
		}
		o[j].re = tot_e; o[j].x_pos_s = (uint32_t)z.x_pos_s; o[j].x_pos_e = (uint32_t)z.x_pos_e; o[j].y_pos_s = (uint32_t)z.y_pos_s; o[j].y_pos_e = (uint32_t)z.y_pos_e;
		o[j].w_n = aux.wn; o[j].c_n = aux.cn;
		if (nw + aux.wn > mw) { mw = (nw + aux.wn) * 2 + 64; W = (hao_wl_t *)realloc(W, mw * sizeof(hao_wl_t)); }
		if (nc + aux.cn > mc) { mc = (nc + aux.cn) * 2 + 64; C = (uint16_t *)realloc(C, mc * 2); }
		if (aux.wn) memcpy(W + nw, aux.w, aux.wn * sizeof(hao_wl_t));
		if (aux.cn) memcpy(C + nc, aux.c, aux.cn * 2);
		nw += aux.wn; nc += aux.cn;
	}
	{ int k; for (k = 0; k < 5; k++) free(b.ez.v.Peq[k]); free(b.ez.v.VP); free(b.ez.v.VN); free(b.ez.v.X); free(b.ez.v.D0); free(b.ez.v.HN); free(b.ez.v.HP); }
	free(qs); free(b.tstr); free(b.ez.cig); free(b.ez.path); free(b.dp_t); free(b.dp_p); free(b.dp_f); free(aux.w); free(aux.c);
	*out = o; *wl = W; *n_wl = nw; *cig = C; *n_cig = nc;
	return b.bad ? -1 : 0;
}

/* ================================================================================================== */
/* step C: reassign_gaps (Correct.cpp:25409-25430, row a11) — left-normalise the indels of every window  */
/* ================================================================================================== */
static void append_cigar_(hao_wl_t *idx, wlv_t *res, uint16_t c, uint32_t l)
{ /* append_cigar, Correct.cpp:25151-25165 */
	if (l <= 0) return;
	uint16_t c0 = (uint16_t)-1; uint32_t l0 = 0;
	if (idx->clen > 0) { c0 = res->c[res->cn - 1] >> 14; l0 = res->c[res->cn - 1] & 0x3fff; }
	if (c0 == c) { l += l0; res->cn--; idx->clen--; }
	wlv_push_trace(res, c, l);
	idx->clen = (uint32_t)(res->cn - idx->cidx);
}

static uint16_t adjust_gap_(hao_wl_t *idx, wlv_t *res, const char *pstr, const char *tstr, int64_t pi, int64_t ti, uint16_t op0, uint16_t **buf, size_t *bm, int64_t *rd_err)
{ /* adjust_gap, Correct.cpp:25167-25250 */
	*rd_err = 0;
	if (idx->clen == 0) { append_cigar_(idx, res, op0, 1); return 0; }
	if (op0 != 2 && op0 != 3) return 0;
	if (op0 == 2) ti--; else pi--;
	uint16_t *ca = res->c + idx->cidx, p, ff; int64_t ci = idx->clen, op, cl, k, l[2]; size_t bn = 0;
	for (ci--, ff = 0; ci >= 0; ci--) {
		op = ca[ci] >> 14; cl = ca[ci] & 0x3fff;
		if (op == 2 || op == 3) {
			p = (uint16_t)((op0 << 14) + 1); cig_push(buf, &bn, bm, p);
			l[0] = cl;
			p = (uint16_t)((op << 14) + l[0]); cig_push(buf, &bn, bm, p);
			break;
		} else if (op == 0) {
			for (k = cl - 1, l[0] = l[1] = 0; k >= 0; k--, pi--, ti--) if (pstr[pi] != tstr[ti]) break;
			l[1] = cl - k - 1; l[0] = k + 1;
			if (l[1] > 0) { p = (uint16_t)((op << 14) + l[1]); cig_push(buf, &bn, bm, p); ff = 1; }
			if (l[0] > 0) { p = (uint16_t)((op0 << 14) + 1); cig_push(buf, &bn, bm, p); }
			if (l[0] > 0) { p = (uint16_t)((op << 14) + l[0]); cig_push(buf, &bn, bm, p); }
		} else {
			for (k = cl - 1, l[0] = cl, l[1] = 0; k >= 0; k--, pi--, ti--) {
				if (pstr[pi] == tstr[ti]) {
					l[1] = l[0] - k - 1; l[0] = k;
					if (l[1] > 0) { p = (uint16_t)((op << 14) + l[1]); cig_push(buf, &bn, bm, p); ff = 1; }
					p = 1; cig_push(buf, &bn, bm, p); /* one match */
					(*rd_err)++;
				}
			}
			if (l[0] > 0) { p = (uint16_t)((op << 14) + l[0]); cig_push(buf, &bn, bm, p); }
			l[0] = 0;
		}
		if (l[0] > 0) break;
	}
	if (!ff) { append_cigar_(idx, res, op0, 1); return 0; }
	else if (ci >= 0) {
		idx->clen = (uint32_t)ci; res->cn = idx->cidx + idx->clen;
		for (k = (int64_t)bn - 1; k >= 0; k--) append_cigar_(idx, res, (*buf)[k] >> 14, (*buf)[k] & 0x3fff);
	} else {
		idx->clen = 0; res->cn = idx->cidx;
		append_cigar_(idx, res, op0, 1);
		for (k = (int64_t)bn - 1; k >= 0; k--) append_cigar_(idx, res, (*buf)[k] >> 14, (*buf)[k] & 0x3fff);
	}
	return 1;
}

static uint16_t ajust_end_cigar_(hao_wl_t *idx, wlv_t *res)
{ /* ajust_end_cigar, Correct.cpp:25252-25272 */
	uint16_t *ca = res->c + idx->cidx, rr = 0; int64_t ci, cn = idx->clen, op, cl;
	if (cn <= 0) return rr;
	for (ci = 0; ci < cn; ci++) { op = ca[ci] >> 14; cl = ca[ci] & 0x3fff; if (op != 1) break; ca[ci] = (uint16_t)((3 << 14) + cl); idx->y_start += (int32_t)cl; rr = 1; }
	for (ci = cn - 1; ci >= 0; ci--) { op = ca[ci] >> 14; cl = ca[ci] & 0x3fff; if (op != 1) break; ca[ci] = (uint16_t)((3 << 14) + cl); idx->y_end -= (int32_t)cl; rr = 1; }
	return rr;
}

static uint16_t move_wins_(const hao_wl_t *zw, const uint16_t *zc, wlv_t *aux, const char *tseq, const char *pseq, uint16_t **buf, size_t *bm, int64_t *tot_re)
{ /* move_wins, Correct.cpp:25274-25365.  tseq = query, pseq = whole target read on the overlap's strand */
	if (zw->error == INT16_MAX && zw->clen == 0 && zw->extra_end < 0) { *wlv_pushp(aux) = *zw; return 0; }
	const uint16_t *cg = zc + zw->cidx; size_t cn = zw->clen, ci; hao_wl_t *p = wlv_pushp(aux); size_t pidx = aux->wn - 1, k2;
	p->x_start = zw->x_start; p->x_end = zw->x_end; p->y_start = zw->y_start; p->y_end = zw->y_end; p->extra_begin = p->extra_end = 0;
	p->error_threshold = 0; p->error = zw->error; p->cidx = (uint32_t)aux->cn; p->clen = 0;
	int64_t pi = zw->y_start, ti = zw->x_start, cl, op, k, rr = 0, re; int mm;
	if (zw->error == 0) { p->clen = (uint32_t)cn; for (k2 = 0; k2 < cn; k2++) cig_push(&aux->c, &aux->cn, &aux->cm, cg[k2]); return 0; }
	for (ci = 0, mm = 0; ci < cn; ci++) { op = cg[ci] >> 14; if (op < 2) mm = 1; else if (mm) break; }
	if (ci >= cn) {
		p->clen = (uint32_t)cn; for (k2 = 0; k2 < cn; k2++) cig_push(&aux->c, &aux->cn, &aux->cm, cg[k2]);
		p = &aux->w[pidx];
		if (ajust_end_cigar_(p, aux)) rr = 1;
		return (uint16_t)rr;
	}
	for (ci = 0, mm = 0; ci < cn; ci++) {
		op = cg[ci] >> 14; cl = cg[ci] & 0x3fff; p = &aux->w[pidx];
		if (op < 2) { append_cigar_(p, aux, (uint16_t)op, (uint32_t)cl); pi += cl; ti += cl; mm = 1; }
		else if (mm == 0) { append_cigar_(p, aux, (uint16_t)op, (uint32_t)cl); if (op == 2) pi += cl; else ti += cl; }
		else {
			for (k = 0; k < cl; k++) {
				if (adjust_gap_(p, aux, pseq, tseq, pi, ti, (uint16_t)op, buf, bm, &re)) { rr = 1; p->error = (int16_t)(p->error - re); *tot_re += re; }
				if (op == 2) pi++; else ti++;
			}
		}
	}
	p = &aux->w[pidx];
	if (ajust_end_cigar_(p, aux)) rr = 1;
	return (uint16_t)rr;
}

typedef struct { int64_t nh_err; uint32_t x_pos_s, x_pos_e, y_pos_s, y_pos_e; uint64_t w_off, w_n, c_off, c_n; } hao_alnC_t;

/* step C for the overlaps steps A + B accepted: a[] (re = the estimate reassign_gaps tests), bB / wlB / cigB = step B's output */
int hao_ec_align_C(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const hao_alnA_t *a, const hao_alnB_t *bB, const hao_wl_t *wlB, const uint16_t *cigB,
                   hao_alnC_t **out, hao_wl_t **wl, uint64_t *n_wl, uint16_t **cig, uint64_t *n_cig)
{
	uint64_t ql = r->len[rid], nw = 0, mw = 0, nc = 0, mc = 0, tm = 0; uint32_t j; hao_wl_t *W = 0; uint16_t *C = 0, *buf = 0; size_t bm = 0;
	char *qs = MALLOC_N(char, ql + 1), *ts = 0; wlv_t aux; hao_alnC_t *o = CALLOC_N(hao_alnC_t, n_ch);
	memset(&aux, 0, sizeof(aux));
	hao_decode(r, rid, qs);
	for (j = 0; j < n_ch; j++) {
		const hao_ovlp_t *c = &ch[j]; int64_t tl = (int64_t)r->len[c->y_id], k, rr = 0, re = 0; const hao_wl_t *zw = wlB + bB[j].w_off; const uint16_t *zc = cigB + bB[j].c_off;
		const hao_wl_t *src_w = zw; const uint16_t *src_c = zc; uint64_t src_wn = bB[j].w_n, src_cn = bB[j].c_n;
		o[j].w_off = nw; o[j].c_off = nc;
		if (a[j].st != 2) continue;
		o[j].nh_err = a[j].re; o[j].x_pos_s = bB[j].x_pos_s; o[j].x_pos_e = bB[j].x_pos_e; o[j].y_pos_s = bB[j].y_pos_s; o[j].y_pos_e = bB[j].y_pos_e;
		if (a[j].re != 0) {
			if ((uint64_t)tl + 1 > tm) { tm = (uint64_t)tl + 64; ts = (char *)realloc(ts, tm); }
			hao_decode_sub(r, c->y_id, 0, tl, (int)c->y_pos_strand, ts); /* recover_UC_Read[_RC] */
			aux.wn = aux.cn = 0;
			for (k = 0; k < (int64_t)bB[j].w_n; k++) if (move_wins_(&zw[k], zc, &aux, qs, ts, &buf, &bm, &re)) rr = 1;
			if (rr) { /* update_overlap_region */
				int64_t xs, xe, ys, ye, xr, yr;
				src_w = aux.w; src_c = aux.c; src_wn = aux.wn; src_cn = aux.cn;
				xs = o[j].x_pos_s; xe = o[j].x_pos_e; ys = o[j].y_pos_s; ye = o[j].y_pos_e;
				if (aux.wn) { xs = aux.w[0].x_start; xe = aux.w[aux.wn - 1].x_end; ys = aux.w[0].y_start; ye = aux.w[aux.wn - 1].y_end; }
				if (xs <= ys) { ys -= xs; xs = 0; } else { xs -= ys; ys = 0; }
				xr = (int64_t)ql - xe - 1; yr = tl - ye - 1;
				if (xr <= yr) { xe = (int64_t)ql - 1; ye += xr; } else { ye = tl - 1; xe += yr; }
				o[j].x_pos_s = (uint32_t)xs; o[j].x_pos_e = (uint32_t)xe; o[j].y_pos_s = (uint32_t)ys; o[j].y_pos_e = (uint32_t)ye;
			}
			o[j].nh_err = a[j].re - re;
		}
		o[j].w_n = src_wn; o[j].c_n = src_cn;
		if (nw + src_wn > mw) { mw = (nw + src_wn) * 2 + 64; W = (hao_wl_t *)realloc(W, mw * sizeof(hao_wl_t)); }
		if (nc + src_cn > mc) { mc = (nc + src_cn) * 2 + 64; C = (uint16_t *)realloc(C, mc * 2); }
		if (src_wn) memcpy(W + nw, src_w, src_wn * sizeof(hao_wl_t));
		if (src_cn) memcpy(C + nc, src_c, src_cn * 2);
		nw += src_wn; nc += src_cn;
	}
	free(qs); free(ts); free(buf); free(aux.w); free(aux.c);
	*out = o; *wl = W; *n_wl = nw; *cig = C; *n_cig = nc;
	return 0;
}

/* ================================================================================================== */
/* phasing: rphase_hc (Correct.cpp:20191-20320, row a13), HiFi path (bd = 0, occ_thres = 1, no HPC mask,   */
/* no quality values, no large-indel phasing) + dedup_chains (ecovlp.cpp:2984-3030)                       */
/* ================================================================================================== */
#define EC_PH_WIN 428 /* WINDOW_MAX_SIZE = WINDOW (375) + (int)(1.0 / HA_MIN_OV_DIFF 0.02) + 3, Correct.h:27 */
#define EC_RS_MIN 64  /* RS_MIN_SIZE, ksort.h */

typedef struct { uint32_t site, overlapID, overlapSite; uint8_t type; uint32_t cov; char misBase; } hev_t; /* haplotype_evdience, Correct.h:130-143 */
typedef struct { uint32_t id, overlap_num, occ_0, occ_1, occ_2; int score; uint32_t site; uint8_t is_homopolymer; } snp_t; /* SnpStats, Correct.h:152-167 */

/* klib's in-place MSD radix sort (KRADIX_SORT_INIT, ksort.h:172-221) on elements of `es` bytes with a `kb`-byte key:
 * unstable, restated move for move so that ties land where the reference puts them */
typedef uint64_t (*rs_key_f)(const void *);
static void grs_insertsort(char *beg, char *end, size_t es, rs_key_f key)
{
	char *i, *j, tmp[32];
	for (i = beg + es; i < end; i += es)
		if (key(i) < key(i - es)) {
			memcpy(tmp, i, es);
			for (j = i; j > beg && key(tmp) < key(j - es); j -= es) memcpy(j, j - es, es);
			memcpy(j, tmp, es);
		}
}
static void grs_sort(char *beg, char *end, size_t es, rs_key_f key, int n_bits, int s)
{
	int size = 1 << n_bits, m = size - 1, k; char *i; struct { char *b, *e; } b[256], *l;
	for (k = 0; k < size; k++) b[k].b = b[k].e = beg;
	for (i = beg; i != end; i += es) b[key(i) >> s & m].e += es;
	for (k = 1; k < size; k++) { b[k].e += b[k - 1].e - beg; b[k].b = b[k - 1].e; }
	for (k = 0; k < size;) {
		if (b[k].b != b[k].e) {
			if ((l = b + (key(b[k].b) >> s & m)) != b + k) {
				char tmp[32], swap[32];
				memcpy(tmp, b[k].b, es);
				do { memcpy(swap, tmp, es); memcpy(tmp, l->b, es); memcpy(l->b, swap, es); l->b += es; l = b + (key(tmp) >> s & m); } while (l != b + k);
				memcpy(b[k].b, tmp, es); b[k].b += es;
			} else b[k].b += es;
		} else ++k;
	}
	for (b[0].b = beg, k = 1; k < size; k++) b[k].b = b[k - 1].e;
	if (s) {
		s = s > n_bits ? s - n_bits : 0;
		for (k = 0; k < size; k++) {
			if ((b[k].e - b[k].b) / (ptrdiff_t)es > EC_RS_MIN) grs_sort(b[k].b, b[k].e, es, key, n_bits, s);
			else if ((b[k].e - b[k].b) / (ptrdiff_t)es > 1) grs_insertsort(b[k].b, b[k].e, es, key);
		}
	}
}
static void grs_radix_sort(void *beg_, size_t n, size_t es, rs_key_f key, int key_bytes)
{
	char *beg = (char *)beg_, *end = beg + n * es;
	if (n <= EC_RS_MIN) grs_insertsort(beg, end, es, key);
	else grs_sort(beg, end, es, key, 8, key_bytes * 8 - 8);
}
static uint64_t key_u64(const void *p) { return *(const uint64_t *)p; }
static uint64_t key_hev_site(const void *p) { return ((const hev_t *)p)->site; }
static uint64_t key_hev_id(const void *p) { return ((const hev_t *)p)->overlapID; }

typedef struct { hev_t *a; size_t n, m; } hevv_t;
static void hev_push(hevv_t *v, const hev_t *e) { if (v->n == v->m) { v->m = v->m ? v->m << 1 : 256; v->a = (hev_t *)realloc(v->a, v->m * sizeof(hev_t)); } v->a[v->n++] = *e; }
typedef struct { snp_t *a; size_t n, m; } snpv_t;
static snp_t *snp_pushp(snpv_t *v) { if (v->n == v->m) { v->m = v->m ? v->m << 1 : 64; v->a = (snp_t *)realloc(v->a, v->m * sizeof(snp_t)); } return &v->a[v->n++]; }

/* one accepted overlap as rphase_hc sees it */
typedef struct { uint32_t y_id, rev, x_pos_s, x_pos_e, align_length; int64_t nh_err; const hao_wl_t *w; uint64_t wn; const uint16_t *c; uint8_t is_match; int8_t strong; } phov_t;

/* extract_sub_cigar_hc (Correct.cpp:18541-18660) without its cursor bookkeeping: every query position of [s, e) the
 * window's cigar covers with a match / mismatch column.  set_f: count mismatch columns into f[]; else emit evidence. */
static void ph_walk(const hao_reads_t *r, const char *qstr, const phov_t *z, uint32_t ovid, const hao_wl_t *w, int64_t s, int64_t e, int set_f, uint8_t *f, hevv_t *hp)
{
	int64_t s0 = w->x_start, e0 = (int64_t)w->x_end + 1, xk = w->x_start, yk = w->y_start, ws, we, os, oe, t; uint32_t ci; int64_t s00 = s;
	if (s < s0) s = s0;
	if (e > e0) e = e0;
	if (s >= e || !w->clen) return;
	for (ci = 0; ci < w->clen && xk < e; ci++) {
		uint32_t op = z->c[w->cidx + ci] >> 14, cl = z->c[w->cidx + ci] & 0x3fff;
		ws = xk;
		if (op != 2) xk += cl;
		if (op != 3) yk += cl;
		we = xk;
		if (op != 0 && op != 1) continue;
		os = s > ws ? s : ws; oe = e < we ? e : we;
		if (oe <= os) continue;
		for (t = os; t < oe; t++) {
			uint8_t *ft = &f[t - s00];
			if (set_f) { if (op == 1) *ft = (uint8_t)(*ft <= 126 ? *ft + 1 : 127); }
			else if (*ft) {
				hev_t ev; char yb;
				ev.overlapID = ovid; ev.site = (uint32_t)t; ev.overlapSite = (uint32_t)(t - xk + yk); ev.cov = 1;
				if (op == 0) { ev.misBase = qstr[t]; ev.type = 0; }
				else { hao_decode_sub(r, z->y_id, t - xk + yk, 1, (int)z->rev, &yb); ev.misBase = yb; ev.type = 1; }
				hev_push(hp, &ev);
			}
		}
	}
}

static int nt6(char c) { switch (c) { case 'A': case 'a': return 0; case 'C': case 'c': return 1; case 'G': case 'g': return 2; case 'T': case 't': return 3; case 'N': case 'n': return 4; default: return 5; } } /* seq_nt6_table */

static size_t ph_push_info(snpv_t *st, hev_t *a, size_t a_n, hev_t *u_a)
{ /* push_info, Correct.cpp:10511-10600 (oa = NULL, v8 = NULL) */
	uint64_t i, k, m, occ_0 = 0, occ_1[6] = { 0, 0, 0, 0, 0, 0 }, occ_2 = 0, diff = 0, rev_n; uint8_t ihpc = 0; snp_t *p;
	grs_radix_sort(a, a_n, sizeof(hev_t), key_hev_id, 4);
	for (k = 1, m = i = 0; k <= a_n; k++) {
		if (k == a_n || a[k].overlapID != a[m].overlapID) {
			a[i] = a[m];
			if ((a[i].type & 1) == 0) { occ_0 += a[i].cov; if (!ihpc && ((a[i].type >> 1) & 1)) ihpc = 1; }
			else { occ_1[nt6(a[i].misBase)] += a[i].cov; diff += a[i].cov; }
			occ_2 += a[i].cov;
			i++; m = k;
		}
	}
	a_n = i;
	if (occ_0 == 0 || diff <= 1) return 0;
	rev_n = occ_2;
	for (i = m = 0; i < 4; i++) {
		if (occ_1[i] >= 2) {
			p = snp_pushp(st);
			p->id = (uint32_t)(st->n - 1); p->occ_0 = (uint32_t)(1 + occ_0); p->occ_1 = (uint32_t)occ_1[i]; p->occ_2 = (uint32_t)(occ_2 - p->occ_0 - p->occ_1);
			p->site = a[0].site; p->score = -1; p->overlap_num = (uint32_t)rev_n; p->is_homopolymer = ihpc;
			occ_1[i] = p->id; m++;
		} else occ_1[i] = (uint64_t)-1;
	}
	occ_1[4] = occ_1[5] = (uint64_t)-1;
	if (m == 0) return 0;
	for (i = m = 0; i < a_n; i++) {
		if ((a[i].type & 1) == 0) a[i].overlapSite = (uint32_t)(st->n - 1);
		else if (occ_1[nt6(a[i].misBase)] != (uint64_t)-1) { a[i].cov = a[i].overlapSite; a[i].overlapSite = (uint32_t)occ_1[nt6(a[i].misBase)]; }
		else continue;
		u_a[m++] = a[i];
	}
	return m;
}

static void ph_naive_hifi(hevv_t *hap, snpv_t *st, phov_t *ol, double up, int s_hap_cov, int infor_cov)
{ /* generate_haplotypes_naive_HiFi, Correct.cpp:8845-9110 (multi_check = 1, st_max = -1: is_st_bs is always false) */
	if (hap->n == 0) return;
	uint64_t k, l, i, o, *a, ii, m_snp = 0, m_list = 0, m_off, *srt = 0; size_t sn = 0, sm = 0; int64_t z; snp_t *s = 0, *t = 0;
#define SRT_PUSH(x) do { if (sn == sm) { sm = sm ? sm << 1 : 64; srt = (uint64_t *)realloc(srt, sm * 8); } srt[sn++] = (x); } while (0)
	for (k = 1, l = 0, i = 0; k <= st->n; ++k) { /* drop SNP sites adjacent to another SNP site */
		if (k == st->n || st->a[k].site != st->a[l].site) {
			if (l > 0 && st->a[l].site == st->a[l - 1].site + 1) { l = k; continue; }
			if (k < st->n && st->a[l].site + 1 == st->a[k].site) { l = k; continue; }
			for (; i < hap->n && hap->a[i].site != st->a[l].site; i++);
			m_off = l - m_snp;
			for (; i < hap->n && hap->a[i].site == st->a[l].site; i++) { hap->a[m_list] = hap->a[i]; hap->a[m_list++].overlapSite -= (uint32_t)m_off; }
			for (; l < k; l++) st->a[m_snp++] = st->a[l];
		}
	}
	st->n = m_snp; hap->n = m_list;
	if (st->n == 0 || hap->n == 0) { free(srt); return; }
	grs_radix_sort(hap->a, hap->n, sizeof(hev_t), key_hev_id, 4);
#define REAL(s_) (!((s_)->occ_0 < 2 || (s_)->occ_1 < 2))
	for (k = 1, l = 0; k <= hap->n; ++k) {
		if (k == hap->n || hap->a[k].overlapID != hap->a[l].overlapID) {
			for (i = l, o = 0; i < k; i++) {
				if ((hap->a[i].type & 1) != 1) continue;
				s = &st->a[hap->a[i].overlapSite];
				if (!REAL(s)) continue;
				if ((int)s->occ_0 >= s_hap_cov && (int)s->occ_1 >= infor_cov) o++;
			}
			if (o > 0) { o = ((uint32_t)-1) - o; o <<= 32; o += l; SRT_PUSH(o); }
			l = k;
		}
	}
	if (sn > 0) {
		grs_radix_sort(srt, sn, 8, key_u64, 8);
		for (k = 0; k < sn; k++) {
			o = 0; l = (uint32_t)srt[k];
			for (i = l; i < hap->n && hap->a[i].overlapID == hap->a[l].overlapID; i++) {
				if ((hap->a[i].type & 1) != 1) continue;
				s = &st->a[hap->a[i].overlapSite];
				if (!REAL(s)) continue;
				if ((int)s->occ_0 >= s_hap_cov && (int)s->occ_1 >= infor_cov) o++;
			}
			if (o == 0) continue;
			ii = hap->a[l].overlapID;
			if (ol[ii].is_match == 1) ol[ii].is_match = 2;
			for (i = l; i < hap->n && hap->a[i].overlapID == hap->a[l].overlapID; i++) {
				if ((hap->a[i].type & 1) == 1) { s = &st->a[hap->a[i].overlapSite]; s->score = 1; }
				else {
					s = &st->a[hap->a[i].overlapSite];
					for (z = hap->a[i].overlapSite; z >= 0; z--) { t = &st->a[z]; if (s->site != t->site) break; t->occ_0 -= hap->a[i].cov; }
				}
			}
		}
		for (k = 0; k < sn; k++) {
			o = 0; l = (uint32_t)srt[k];
			for (i = l; i < hap->n && hap->a[i].overlapID == hap->a[l].overlapID; i++) {
				if ((hap->a[i].type & 1) != 1) continue;
				s = &st->a[hap->a[i].overlapSite];
				if (!REAL(s)) continue;
				if (s->score == 1) o++;
			}
			ii = hap->a[l].overlapID;
			if (ol[ii].is_match == 1 && o > 0) ol[ii].is_match = 2;
		}
		for (k = 1, l = 0; k <= hap->n; ++k) {
			if (k == hap->n || hap->a[k].overlapID != hap->a[l].overlapID) {
				ii = hap->a[l].overlapID;
				if (ol[ii].is_match == 1) for (i = l; i < k; i++) if ((hap->a[i].type & 1) == 1) st->a[hap->a[i].overlapSite].score = -1;
				l = k;
			}
		}
	}
	/* multi_check */
	sn = 0;
	for (k = 1, l = 0; k <= hap->n; ++k) {
		if (k == hap->n || hap->a[k].overlapID != hap->a[l].overlapID) {
			if (ol[hap->a[l].overlapID].is_match == 2) { l = k; continue; }
			for (i = l, o = 0; i < k; i++) {
				if ((hap->a[i].type & 1) != 1) continue;
				s = &st->a[hap->a[i].overlapSite];
				if (!REAL(s)) continue;
				if ((int)s->occ_0 >= s_hap_cov && (int)s->occ_1 >= infor_cov) continue;
				if (s->score == 1) continue;
				o++; SRT_PUSH(hap->a[i].overlapSite);
			}
			sn -= o;
			if ((double)o >= (double)ol[hap->a[l].overlapID].align_length * up) {
				grs_radix_sort(srt + sn, o, 8, key_u64, 8);
				a = srt + sn;
				for (i = z = 0; i < o; i++) { /* s / t keep their last values across overlaps, like the reference's locals */
					if (i > 0) s = &st->a[a[i - 1]];
					if (i + 1 < o) t = &st->a[a[i + 1]];
					if (s && s->site + 32 > st->a[a[i]].site) continue;
					if (t && st->a[a[i]].site + 32 > t->site) continue;
					a[z] = a[i]; z++;
				}
				if (z >= 2) sn += (size_t)z;
			}
			l = k;
		}
	}
	if (sn > 0) {
		grs_radix_sort(srt, sn, 8, key_u64, 8);
		for (k = 1, l = 0; k <= sn; ++k) {
			if (k == sn || srt[k] != srt[l]) { if (k - l >= 2) st->a[srt[l]].score = 1; }
			l = k;
		}
	}
	for (k = 1, l = 0; k <= hap->n; ++k) {
		if (k == hap->n || hap->a[k].overlapID != hap->a[l].overlapID) {
			ii = hap->a[l].overlapID;
			if (ol[ii].is_match == 2) ol[ii].strong = 1;
			else if (ol[ii].is_match == 1) {
				for (i = l; i < k; i++) {
					s = &st->a[hap->a[i].overlapSite];
					if (s->score == 1 && REAL(s)) { ol[ii].strong = 1; if ((hap->a[i].type & 1) == 1) { ol[ii].is_match = 2; break; } }
				}
			}
			l = k;
		}
	}
#undef REAL
#undef SRT_PUSH
	free(srt);
}

/* rphase_hc for one read.  ov[] = the accepted overlaps in list order (overlapID = index); on return is_match / strong are set. */
static void ph_rphase(const hao_reads_t *r, uint32_t rid, phov_t *ov, uint32_t n_ov)
{
	int64_t ql = (int64_t)r->len[rid], s, e, k, i, t, srt_n, rr = 0, wl = EC_PH_WIN; uint32_t j, w; uint8_t flag[EC_PH_WIN];
	char *qstr = MALLOC_N(char, ql + 1); hevv_t hp; snpv_t st; uint64_t *idx = 0, ni = 0, mi = 0, m, rm_n; uint32_t nn_snp = 0;
	struct cw { uint32_t ov, w; } *cidx = 0; uint64_t nc = 0, mc = 0;
	memset(&hp, 0, sizeof(hp)); memset(&st, 0, sizeof(st)); memset(flag, 0, sizeof(flag));
	hao_decode(r, rid, qstr);
#define IDX_PUSH(x) do { if (ni == mi) { mi = mi ? mi << 1 : 256; idx = (uint64_t *)realloc(idx, mi * 8); } idx[ni++] = (x); } while (0)
	for (j = 0; j < n_ov; j++)
		for (w = 0; w < ov[j].wn; w++) {
			const hao_wl_t *u = &ov[j].w[w];
			if (u->error == INT16_MAX && u->clen == 0 && u->extra_end < 0) continue;
			if (u->x_end >= u->x_start) {
				if (nc == mc) { mc = mc ? mc << 1 : 256; cidx = (struct cw *)realloc(cidx, mc * sizeof(*cidx)); }
				IDX_PUSH(((uint64_t)u->x_start << 32) + nc);
				cidx[nc].ov = j; cidx[nc].w = w; nc++;
			}
		}
	srt_n = (int64_t)ni;
	grs_radix_sort(idx, ni, 8, key_u64, 8);
	for (k = 1, i = 0; k <= srt_n; k++) {
		if (k == srt_n || (idx[k] >> 32) != (idx[i] >> 32)) {
			if (k - i > 1) {
				for (t = i; t < k; t++) { const struct cw *cp = &cidx[(uint32_t)idx[t]]; m = (uint64_t)(ov[cp->ov].w[cp->w].x_end + 1); m <<= 32; m += (uint32_t)idx[t]; idx[t] = m; }
				grs_radix_sort(idx + i, (size_t)(k - i), 8, key_u64, 8);
			}
			i = k;
		}
	}
	i = 0; s = 0; e = wl; e = e <= ql ? e : ql;
	for (; s < ql;) {
		size_t l0; int64_t wi, wl0; uint64_t si = (uint64_t)-1, ei = 0; int fi = 0;
		if (rr) {
			for (m = rm_n = (uint64_t)srt_n; m < ni; m++) {
				const struct cw *cp = &cidx[idx[m]]; int64_t q0 = ov[cp->ov].w[cp->w].x_start, q1 = (int64_t)ov[cp->ov].w[cp->w].x_end + 1, os = q0 > s ? q0 : s, oe = q1 < e ? q1 : e;
				if (oe > os) idx[rm_n++] = idx[m];
			}
			ni = rm_n;
		}
		for (; i < srt_n; ++i) {
			const struct cw *cp = &cidx[(uint32_t)idx[i]]; int64_t q0 = ov[cp->ov].w[cp->w].x_start, q1 = (int64_t)ov[cp->ov].w[cp->w].x_end + 1, os, oe;
			if (q0 >= e) break;
			os = q0 > s ? q0 : s; oe = q1 < e ? q1 : e;
			if (oe > os) IDX_PUSH((uint32_t)idx[i]);
		}
		l0 = hp.n; rr = 0;
		for (m = (uint64_t)srt_n; m < ni; m++) { /* hc_phase_robust_rr(set_f = 1), Correct.cpp:19065-19079 */
			const struct cw *cp = &cidx[idx[m]]; const hao_wl_t *u = &ov[cp->ov].w[cp->w];
			if ((int64_t)u->x_end + 1 <= e) rr = 1;
			ph_walk(r, qstr, &ov[cp->ov], cp->ov, u, s, e, 1, flag, &hp);
		}
		for (wi = 0, wl0 = e - s; wi < wl0; wi++) {
			if (flag[wi] > 0) {
				if (flag[wi] > 1) { fi = 1; nn_snp++; flag[wi] = 1; ei = (uint64_t)wi + 1; if (si == (uint64_t)-1) si = (uint64_t)wi; }
				else flag[wi] = 0;
			}
		}
		if (fi) {
			rr = 0;
			for (m = (uint64_t)srt_n; m < ni; m++) {
				const struct cw *cp = &cidx[idx[m]]; const hao_wl_t *u = &ov[cp->ov].w[cp->w];
				if ((int64_t)u->x_end + 1 <= e) rr = 1;
				ph_walk(r, qstr, &ov[cp->ov], cp->ov, u, s, e, 0, flag, &hp);
			}
			if (hp.n > l0) grs_radix_sort(hp.a + l0, hp.n - l0, sizeof(hev_t), key_hev_site, 4);
		}
		if (ei > si) memset(flag + si, 0, (size_t)(ei - si));
		s += wl; e += wl; e = e <= ql ? e : ql;
	}
	(void)nn_snp;
	{
		size_t n = hp.n, tt = 0, ii = 0, kk;
		for (kk = 1; kk <= n; ++kk) if (kk == n || hp.a[kk].site != hp.a[ii].site) { tt += ph_push_info(&st, hp.a + ii, kk - ii, hp.a + tt); ii = kk; }
		hp.n = tt;
	}
	ph_naive_hifi(&hp, &st, ov, 0.04, 3, 3); /* asm_opt.s_hap_cov = infor_cov = 3, CommandLines.cpp:333-334 */
#undef IDX_PUSH
	free(qstr); free(hp.a); free(st.a); free(idx); free(cidx);
}

typedef struct { uint32_t y_id, rev, x_pos_s, x_pos_e, y_pos_s, y_pos_e, nh_err, is_match; int32_t strong; } hao_phase_t;

/* worker_hap_ec steps 5-6 (ecovlp.cpp:3299-3306) for one read: rphase_hc over the accepted overlaps (steps A-C output), then
 * dedup_chains.  out = accepted overlaps after rphase_hc in list order; dd = the list dedup_chains leaves. */
int hao_ec_phase(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const hao_alnA_t *a, const hao_alnC_t *c, const hao_wl_t *wl, const uint16_t *cig,
                 hao_phase_t **out, uint32_t *n_out, hao_phase_t **dd, uint32_t *n_dd)
{
	uint32_t j, n = 0, k, l, mm; phov_t *ov = MALLOC_N(phov_t, n_ch); hao_phase_t *o = MALLOC_N(hao_phase_t, n_ch), *d = MALLOC_N(hao_phase_t, n_ch);
	for (j = 0; j < n_ch; j++) {
		if (a[j].st != 2) continue;
		ov[n].y_id = ch[j].y_id; ov[n].rev = ch[j].y_pos_strand; ov[n].x_pos_s = c[j].x_pos_s; ov[n].x_pos_e = c[j].x_pos_e; ov[n].align_length = a[j].align_length; ov[n].nh_err = c[j].nh_err;
		ov[n].w = wl + c[j].w_off; ov[n].wn = c[j].w_n; ov[n].c = cig + c[j].c_off; ov[n].is_match = 1; ov[n].strong = 0;
		o[n].y_id = ch[j].y_id; o[n].rev = ch[j].y_pos_strand; o[n].x_pos_s = c[j].x_pos_s; o[n].x_pos_e = c[j].x_pos_e; o[n].y_pos_s = c[j].y_pos_s; o[n].y_pos_e = c[j].y_pos_e; o[n].nh_err = (uint32_t)c[j].nh_err;
		n++;
	}
	ph_rphase(r, rid, ov, n);
	for (j = 0; j < n; j++) { o[j].is_match = ov[j].is_match; o[j].strong = ov[j].strong; d[j] = o[j]; }
	*out = o; *n_out = n;
	/* dedup_chains, ecovlp.cpp:2984-3030: sort by y_id (klib radix sort on the records), keep the best chain of every target */
	mm = n;
	if (n > 1) {
		hao_ovlp_t *tmp = CALLOC_N(hao_ovlp_t, n); hao_phase_t *d2 = MALLOC_N(hao_phase_t, n);
		for (j = 0; j < n; j++) { tmp[j].y_id = d[j].y_id; tmp[j].x_pos_s = j; } /* carry the index through the reference's unstable sort */
		ov_sort_y_id(tmp, n);
		for (j = 0; j < n; j++) d2[j] = d[tmp[j].x_pos_s];
		memcpy(d, d2, n * sizeof(hao_phase_t)); free(tmp); free(d2);
		for (k = 1, l = mm = 0; k <= n; k++) {
			if (k == n || d[k].y_id != d[l].y_id) {
				uint32_t mm_k = l, s, mm_m;
				if (k - l > 1) {
					int64_t mm_sc = INT32_MIN, sc;
					for (s = l, mm_k = (uint32_t)-1, mm_m = 3; s < k; s++) {
						int sf = 0;
						sc = ((int64_t)d[s].x_pos_e + 1 - d[s].x_pos_s) - (int64_t)d[s].nh_err * 12;
						if (d[s].is_match < mm_m) sf = 1;
						else if (d[s].is_match == mm_m) {
							if (sc > mm_sc) sf = 1;
							else if (sc == mm_sc && (d[s].x_pos_e + 1 - d[s].x_pos_s) > (d[mm_k].x_pos_e + 1 - d[mm_k].x_pos_s)) sf = 1;
						}
						if (sf) { mm_sc = sc; mm_k = s; mm_m = d[s].is_match; }
					}
				}
				if (mm_k != (uint32_t)-1) { if (mm_k != mm) { hao_phase_t t = d[mm_k]; d[mm_k] = d[mm]; d[mm] = t; } mm++; }
				l = k;
			}
		}
	}
	*dd = d; *n_dd = mm;
	free(ov);
	return 0;
}

/* ================================================================================================== */
/* row a12: the exact shortcut of gen_hc_r_alin_ea (ecovlp.cpp:2810-2866)                               */
/* ================================================================================================== */
/* flag[j] = 1 when chain j is accepted without alignment: the read's overlap list of the previous round (in[], n_in) holds an
 * exact (el) record for the same target and strand — the first such record in list order — with the same coordinates, and the
 * two substrings are still identical. */
int hao_ec_ea_flags(const hao_reads_t *r, uint32_t rid, const hao_ovlp_t *ch, uint32_t n_ch, const hao_ma_t *in, uint32_t n_in, uint8_t *flag)
{
	uint64_t *ei = MALLOC_N(uint64_t, n_in + 1), *oi = MALLOC_N(uint64_t, n_ch + 1), en = 0, k, i, ql = r->len[rid], tm = 0; char *qs = 0, *ts = 0;
	memset(flag, 0, n_ch);
	for (k = 0; k < n_in; k++) if (in[k].el) ei[en++] = ((((uint64_t)in[k].tn << 1) | (in[k].rev & 1)) << 32) | k;
	if (!en) { free(ei); free(oi); return 0; }
	for (k = 0; k < n_ch; k++) oi[k] = ((((uint64_t)ch[k].y_id << 1) | ch[k].y_pos_strand) << 32) | k;
	qsort(ei, en, 8, u64_cmp); qsort(oi, n_ch, 8, u64_cmp); /* radix_sort_ec64 on unique keys */
	qs = MALLOC_N(char, ql + 1); hao_decode(r, rid, qs);
	for (k = i = 0; k < n_ch; k++) {
		const hao_ovlp_t *z = &ch[(uint32_t)oi[k]]; uint64_t key = ((uint64_t)z->y_id << 1) | z->y_pos_strand;
		for (; i < en && (ei[i] >> 32) < key; i++);
		if (i < en && (ei[i] >> 32) == key) {
			const hao_ma_t *p = &in[(uint32_t)ei[i]];
			if (z->x_pos_s == (uint32_t)p->qns && z->x_pos_e + 1 == p->qe && z->y_pos_s == p->ts && z->y_pos_e + 1 == p->te) {
				uint64_t tl = p->te - p->ts;
				if (tl + 1 > tm) { tm = tl + 64; ts = (char *)realloc(ts, tm); }
				hao_decode_sub(r, z->y_id, p->ts, (int64_t)tl, (int)z->y_pos_strand, ts);
				if ((uint64_t)(p->qe - (uint32_t)p->qns) == tl && memcmp(qs + (uint32_t)p->qns, ts, tl) == 0) flag[(uint32_t)oi[k]] = 1; /* exact_ec_check, ecovlp.cpp:2803 */
			}
		}
	}
	free(ei); free(oi); free(qs); free(ts);
	return 0;
}
