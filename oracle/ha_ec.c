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
