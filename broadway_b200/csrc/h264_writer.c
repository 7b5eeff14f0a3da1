/* h264_writer.c — synthetic H.264 Baseline (CAVLC) Annex-B bitstream writer.
 * See include/h264b200_writer.h for what it is for.  Syntax written here is the
 * syntax the reference parses in h264bsd_seq_param_set.c:98-330,
 * h264bsd_pic_param_set.c:88-330, h264bsd_slice_header.c:97-420,
 * h264bsd_slice_data.c:85-235, h264bsd_macroblock_layer.c:133-869 and
 * h264bsd_cavlc.c:748-915; predictor state follows ITU-T H.264 8.3.1.1 (intra
 * 4x4 mode), 8.4.1 (motion vectors) and 9.2.1 (nC). */
#include <stdlib.h>
#include <string.h>
#include "h264b200_writer.h"
#include "h264_consts.h"
#include "cavlc_tables.h"
#include "h264_fmo.h"

/* ------------------------------------------------------------------ PRNG */
typedef struct { uint64_t s; } rng_t;
static uint64_t rng_next(rng_t *r)
{
    uint64_t x = r->s;
    x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
    r->s = x;
    return x * 0x2545F4914F6CDD1DULL;
}
static uint32_t rng_u(rng_t *r, uint32_t n) { return n ? (uint32_t)((rng_next(r) >> 33) % n) : 0; }
static int rng_range(rng_t *r, int lo, int hi) { return lo + (int)rng_u(r, (uint32_t)(hi - lo + 1)); }
static int rng_permille(rng_t *r, uint32_t p) { return rng_u(r, 1000) < p; }

/* ------------------------------------------------------------ bit writer */
typedef struct { uint8_t *buf; size_t cap, pos; uint32_t acc; int n; int ovf; } bitw_t;
static void bw_init(bitw_t *b, uint8_t *buf, size_t cap) { b->buf = buf; b->cap = cap; b->pos = 0; b->acc = 0; b->n = 0; b->ovf = 0; }
static void bw_put(bitw_t *b, int nbits, uint32_t v)
{
    while (nbits > 0) {
        int take = nbits > 8 ? 8 : nbits;
        uint32_t chunk = (v >> (nbits - take)) & ((1u << take) - 1);
        int i;
        nbits -= take;
        for (i = take - 1; i >= 0; i--) {
            b->acc = (b->acc << 1) | ((chunk >> i) & 1);
            if (++b->n == 8) {
                if (b->pos < b->cap) b->buf[b->pos++] = (uint8_t)b->acc; else b->ovf = 1;
                b->acc = 0; b->n = 0;
            }
        }
    }
}
static void bw_ue(bitw_t *b, uint32_t v)
{
    uint32_t x = v + 1; int len = 0;
    while ((x >> len) > 1) len++;
    bw_put(b, len, 0);
    bw_put(b, len + 1, x);
}
static void bw_se(bitw_t *b, int v) { bw_ue(b, v > 0 ? (uint32_t)(2 * v - 1) : (uint32_t)(-2 * v)); }
static void bw_trailing(bitw_t *b) { bw_put(b, 1, 1); while (b->n) bw_put(b, 1, 0); }
static void bw_align_zero(bitw_t *b) { while (b->n) bw_put(b, 1, 0); }

/* Annex-B NAL: 00 00 00 01, header, payload with emulation prevention (7.4.1) */
static size_t emit_nal(uint8_t *out, size_t cap, int ref_idc, int type, const uint8_t *rbsp, size_t n)
{
    size_t o = 0, i; int zeros = 0;
    if (cap < n + n / 2 + 8) return 0;
    out[o++] = 0; out[o++] = 0; out[o++] = 0; out[o++] = 1;
    out[o++] = (uint8_t)((ref_idc << 5) | type);
    for (i = 0; i < n; i++) {
        if (zeros >= 2 && rbsp[i] <= 3) { out[o++] = 3; zeros = 0; }
        out[o++] = rbsp[i];
        zeros = rbsp[i] == 0 ? zeros + 1 : 0;
    }
    return o;
}

/* ------------------------------------------------------ macroblock state */
enum { K_NONE = 0, K_INTER, K_I4, K_I16, K_PCM };
typedef struct {
    uint8_t kind, skip; uint16_t slice;
    uint8_t tc[24];          /* TotalCoeff: luma by luma4x4BlkIdx, then Cb 16..19, Cr 20..23 */
    uint8_t i4mode[16];      /* by luma4x4BlkIdx */
    int8_t  ref[4];          /* refIdxL0 per 8x8 */
    int16_t mv[16][2];       /* by luma4x4BlkIdx */
} wmb_t;

typedef struct {
    const h264w_params_t *p;
    rng_t rng;
    uint32_t W, H, nmb;
    wmb_t *mb;
    uint32_t cur_slice, cur_addr;
    int qp;                  /* running QP_Y */
    int num_ref_active;
    int is_p;
} wr_t;

static wmb_t *nb_mb(wr_t *w, int mbx, int mby)
{
    uint32_t addr;
    if (mbx < 0 || mby < 0 || mbx >= (int)w->W || mby >= (int)w->H) return NULL;
    addr = (uint32_t)mby * w->W + (uint32_t)mbx;
    if (addr >= w->cur_addr) return NULL;                 /* not yet decoded */
    if (w->mb[addr].slice != w->cur_slice) return NULL;   /* other slice: unavailable */
    return &w->mb[addr];
}

/* ----------------------------------------------------------- CAVLC (9.2) */
static void put_code(bitw_t *b, vlc_code_t c) { bw_put(b, c.len, c.code); }

/* c[0..n) in scan order; nC < 0 selects the chroma DC tables. Returns TotalCoeff. */
static int write_block(bitw_t *b, const int *c, int n, int nC)
{
    int lev[16], run[16], tc = 0, t1 = 0, tz = 0, i, k, last = -1, sl, zl;
    for (i = n - 1; i >= 0; i--) if (c[i]) { if (last < 0) last = i; lev[tc++] = c[i]; }
    if (tc) {
        int idx = 0;
        tz = last + 1 - tc;
        for (i = last, k = -1; i >= 0; i--) {
            if (c[i]) { k++; run[k] = 0; idx++; }
            else run[k]++;
        }
        (void)idx;
        for (k = 0; k < tc && k < 3 && (lev[k] == 1 || lev[k] == -1); k++) t1++;
    }
    if (nC < 0) put_code(b, H264_COEFF_TOKEN_CHROMA_DC[tc][t1]);
    else put_code(b, H264_COEFF_TOKEN[nC < 2 ? 0 : nC < 4 ? 1 : nC < 8 ? 2 : 3][tc][t1]);
    if (!tc) return 0;
    sl = (tc > 10 && t1 < 3) ? 1 : 0;
    for (k = 0; k < tc; k++) {
        int L = lev[k], code, a;
        if (k < t1) { bw_put(b, 1, L < 0); continue; }
        code = L > 0 ? 2 * L - 2 : -2 * L - 1;
        if (k == t1 && t1 < 3) code -= 2;
        if (sl == 0) {
            if (code < 14) { bw_put(b, code, 0); bw_put(b, 1, 1); }
            else if (code < 30) { bw_put(b, 14, 0); bw_put(b, 1, 1); bw_put(b, 4, (uint32_t)(code - 14)); }
            else { bw_put(b, 15, 0); bw_put(b, 1, 1); bw_put(b, 12, (uint32_t)(code - 30)); }
        } else {
            if (code < (15 << sl)) { bw_put(b, code >> sl, 0); bw_put(b, 1, 1); bw_put(b, sl, (uint32_t)(code & ((1 << sl) - 1))); }
            else { bw_put(b, 15, 0); bw_put(b, 1, 1); bw_put(b, 12, (uint32_t)(code - (15 << sl))); }
        }
        if (sl == 0) sl = 1;
        a = L < 0 ? -L : L;
        if (a > (3 << (sl - 1)) && sl < 6) sl++;
    }
    if (tc < n) {
        if (nC < 0) put_code(b, H264_TOTAL_ZEROS_CHROMA_DC[tc - 1][tz]);
        else put_code(b, H264_TOTAL_ZEROS[tc - 1][tz]);
    }
    zl = tz;
    for (k = 0; k < tc - 1 && zl > 0; k++) {
        put_code(b, H264_RUN_BEFORE[(zl > 7 ? 7 : zl) - 1][run[k]]);
        zl -= run[k];
    }
    return tc;
}

/* largest |level| the escape code (level_prefix 15) can carry at any suffixLength */
#define LEVEL_LIMIT 2000

/* nC for luma block blk of the current macroblock (9.2.1) */
static int luma_nc(wr_t *w, const wmb_t *cur, int mbx, int mby, int blk)
{
    int x4 = H264_BLK_X[blk] >> 2, y4 = H264_BLK_Y[blk] >> 2, na = 0, nb = 0, aa = 0, ab = 0;
    if (x4 > 0) { aa = 1; na = cur->tc[H264_RASTER_TO_BLK[y4 * 4 + x4 - 1]]; }
    else { wmb_t *m = nb_mb(w, mbx - 1, mby); if (m) { aa = 1; na = m->tc[H264_RASTER_TO_BLK[y4 * 4 + 3]]; } }
    if (y4 > 0) { ab = 1; nb = cur->tc[H264_RASTER_TO_BLK[(y4 - 1) * 4 + x4]]; }
    else { wmb_t *m = nb_mb(w, mbx, mby - 1); if (m) { ab = 1; nb = m->tc[H264_RASTER_TO_BLK[12 + x4]]; } }
    if (aa && ab) return (na + nb + 1) >> 1;
    return aa ? na : ab ? nb : 0;
}
/* nC for chroma AC block c (0..3) of plane pl (0 Cb, 1 Cr) */
static int chroma_nc(wr_t *w, const wmb_t *cur, int mbx, int mby, int pl, int c)
{
    int x = c & 1, y = c >> 1, base = 16 + 4 * pl, na = 0, nb = 0, aa = 0, ab = 0;
    if (x > 0) { aa = 1; na = cur->tc[base + y * 2]; }
    else { wmb_t *m = nb_mb(w, mbx - 1, mby); if (m) { aa = 1; na = m->tc[base + y * 2 + 1]; } }
    if (y > 0) { ab = 1; nb = cur->tc[base + x]; }
    else { wmb_t *m = nb_mb(w, mbx, mby - 1); if (m) { ab = 1; nb = m->tc[base + 2 + x]; } }
    if (aa && ab) return (na + nb + 1) >> 1;
    return aa ? na : ab ? nb : 0;
}

/* ------------------------------------------- residual range check (8.5) */
static void idct4x4(const int *d, int *r)   /* d, r raster 4x4 */
{
    int f[16], i;
    for (i = 0; i < 4; i++) {
        const int *s = d + 4 * i;
        int e0 = s[0] + s[2], e1 = s[0] - s[2], e2 = (s[1] >> 1) - s[3], e3 = s[1] + (s[3] >> 1);
        f[4*i] = e0 + e3; f[4*i+1] = e1 + e2; f[4*i+2] = e1 - e2; f[4*i+3] = e0 - e3;
    }
    for (i = 0; i < 4; i++) {
        int e0 = f[i] + f[8+i], e1 = f[i] - f[8+i], e2 = (f[4+i] >> 1) - f[12+i], e3 = f[4+i] + (f[12+i] >> 1);
        r[i] = (e0 + e3 + 32) >> 6; r[4+i] = (e1 + e2 + 32) >> 6; r[8+i] = (e1 - e2 + 32) >> 6; r[12+i] = (e0 - e3 + 32) >> 6;
    }
}
static int in_range16(const int *r) { int i; for (i = 0; i < 16; i++) if (r[i] < -512 || r[i] > 511) return 0; return 1; }
/* scan-order levels (start..15) + optional pre-scaled DC -> residual ok? */
static int block_ok(const int *scan, int start, int qp, int have_dc, int dc)
{
    int d[16], r[16], i;
    memset(d, 0, sizeof d);
    for (i = start; i < 16; i++) {
        int pos = H264_ZIGZAG4x4[i];
        d[pos] = (scan[i] * H264_LEVEL_SCALE[qp % 6][H264_POS_CLASS[pos]]) << (qp / 6);
    }
    if (have_dc) d[0] = dc;
    idct4x4(d, r);
    return in_range16(r);
}
static void luma_dc_scaled(const int *scan, int qp, int *dc /* raster 4x4 of block positions */)
{
    int c[16], f[16], i, ls = H264_LEVEL_SCALE[qp % 6][0];
    memset(c, 0, sizeof c);
    for (i = 0; i < 16; i++) c[H264_ZIGZAG4x4[i]] = scan[i];
    /* f = H c H with H = [1 1 1 1; 1 1 -1 -1; 1 -1 -1 1; 1 -1 1 -1] (8.5.10) */
    for (i = 0; i < 4; i++) {
        int *s = c + 4 * i;
        f[4*i]   = s[0] + s[1] + s[2] + s[3];
        f[4*i+1] = s[0] + s[1] - s[2] - s[3];
        f[4*i+2] = s[0] - s[1] - s[2] + s[3];
        f[4*i+3] = s[0] - s[1] + s[2] - s[3];
    }
    for (i = 0; i < 4; i++) {
        int g0 = f[i] + f[4+i] + f[8+i] + f[12+i], g1 = f[i] + f[4+i] - f[8+i] - f[12+i];
        int g2 = f[i] - f[4+i] - f[8+i] + f[12+i], g3 = f[i] - f[4+i] + f[8+i] - f[12+i];
        int g[4], k; g[0] = g0; g[1] = g1; g[2] = g2; g[3] = g3;
        for (k = 0; k < 4; k++) {
            int v = g[k] * ls;
            dc[4*k + i] = qp >= 12 ? v << (qp / 6 - 2) : (v + (1 << (1 - qp / 6))) >> (2 - qp / 6);
        }
    }
}
static void chroma_dc_scaled(const int *c4, int qpc, int *dc)
{
    int ls = H264_LEVEL_SCALE[qpc % 6][0], f[4], k;
    f[0] = c4[0] + c4[1] + c4[2] + c4[3]; f[1] = c4[0] - c4[1] + c4[2] - c4[3];
    f[2] = c4[0] + c4[1] - c4[2] - c4[3]; f[3] = c4[0] - c4[1] - c4[2] + c4[3];
    for (k = 0; k < 4; k++) { int v = f[k] * ls; dc[k] = qpc >= 6 ? v << (qpc / 6 - 1) : v >> 1; }
}

/* draw coefficients for one block into scan[start..15]; returns count */
static int draw_block(wr_t *w, int *scan, int start)
{
    const h264w_params_t *p = w->p;
    int n = 1 + (int)rng_u(&w->rng, p->max_coeffs ? p->max_coeffs : 1), placed = 0, tries = 0, i;
    int span = 16 - start;
    for (i = 0; i < 16; i++) scan[i] = 0;
    if (n > span) n = span;
    while (placed < n && tries++ < 64) {
        /* low scan positions are favoured (min of two uniforms) */
        int a = (int)rng_u(&w->rng, (uint32_t)span), bq = (int)rng_u(&w->rng, (uint32_t)span);
        int pos = start + (a < bq ? a : bq);
        int mag;
        if (scan[pos]) continue;
        mag = 1 + (int)rng_u(&w->rng, (uint32_t)(p->max_level > 0 ? p->max_level : 1));
        if (rng_u(&w->rng, 3) == 0) mag = 1;               /* plenty of trailing ones */
        scan[pos] = rng_u(&w->rng, 2) ? mag : -mag;
        placed++;
    }
    return placed;
}
static void shrink(int *scan) { int i; for (i = 0; i < 16; i++) scan[i] /= 2; }
static int count_nz(const int *scan, int start) { int i, n = 0; for (i = start; i < 16; i++) n += scan[i] != 0; return n; }

/* ------------------------------------------------ intra mode bookkeeping */
static int mb_is_intra(const wmb_t *m) { return m->kind == K_I4 || m->kind == K_I16 || m->kind == K_PCM; }
/* neighbour usable for intra *sample* prediction */
static int intra_avail(wr_t *w, int mbx, int mby)
{
    wmb_t *m = nb_mb(w, mbx, mby);
    if (!m) return 0;
    if (w->p->constrained_intra_pred && !mb_is_intra(m)) return 0;
    return 1;
}
static int pred_i4mode(wr_t *w, const wmb_t *cur, int mbx, int mby, int blk)
{
    int x4 = H264_BLK_X[blk] >> 2, y4 = H264_BLK_Y[blk] >> 2, ma, mbm, dc = 0;
    if (x4 > 0) ma = cur->i4mode[H264_RASTER_TO_BLK[y4 * 4 + x4 - 1]];
    else {
        wmb_t *m = nb_mb(w, mbx - 1, mby);
        if (!m || (w->p->constrained_intra_pred && !mb_is_intra(m))) { dc = 1; ma = 2; }
        else ma = m->kind == K_I4 ? m->i4mode[H264_RASTER_TO_BLK[y4 * 4 + 3]] : 2;
    }
    if (y4 > 0) mbm = cur->i4mode[H264_RASTER_TO_BLK[(y4 - 1) * 4 + x4]];
    else {
        wmb_t *m = nb_mb(w, mbx, mby - 1);
        if (!m || (w->p->constrained_intra_pred && !mb_is_intra(m))) { dc = 1; mbm = 2; }
        else mbm = m->kind == K_I4 ? m->i4mode[H264_RASTER_TO_BLK[12 + x4]] : 2;
    }
    if (dc) return 2;
    return ma < mbm ? ma : mbm;
}

/* --------------------------------------------------- motion vector state */
typedef struct { int avail, ref, x, y; } mvn_t;
static mvn_t mv_nb(wr_t *w, const wmb_t *cur, int mbx, int mby, int x4, int y4, unsigned done)
{
    mvn_t n; int nx = mbx, ny = mby; const wmb_t *m; int blk;
    n.avail = 0; n.ref = -1; n.x = n.y = 0;
    if (x4 < 0) { nx--; x4 += 4; } else if (x4 > 3) { nx++; x4 -= 4; }
    if (y4 < 0) { ny--; y4 += 4; } else if (y4 > 3) return n;
    blk = H264_RASTER_TO_BLK[y4 * 4 + x4];
    if (nx == mbx && ny == mby) {
        if (!((done >> blk) & 1)) return n;
        m = cur;
    } else {
        m = nb_mb(w, nx, ny);
        if (!m) return n;
    }
    n.avail = 1;
    if (m->kind == K_INTER) { n.ref = m->ref[blk >> 2]; n.x = m->mv[blk][0]; n.y = m->mv[blk][1]; }
    return n;
}
static int median3(int a, int b, int c) { int mx = a > b ? a : b, mn = a < b ? a : b; return c > mx ? mx : c < mn ? mn : c; }
/* dir: 0 median, 1 prefer A, 2 prefer B, 3 prefer C (8.4.1.3) */
static void mv_pred(wr_t *w, const wmb_t *cur, int mbx, int mby, int x4, int y4, int w4, int ref, unsigned done, int dir, int *px, int *py)
{
    mvn_t a = mv_nb(w, cur, mbx, mby, x4 - 1, y4, done);
    mvn_t b = mv_nb(w, cur, mbx, mby, x4, y4 - 1, done);
    mvn_t c = mv_nb(w, cur, mbx, mby, x4 + w4, y4 - 1, done);
    if (!c.avail) c = mv_nb(w, cur, mbx, mby, x4 - 1, y4 - 1, done);
    if (dir == 1 && a.ref == ref) { *px = a.x; *py = a.y; return; }
    if (dir == 2 && b.ref == ref) { *px = b.x; *py = b.y; return; }
    if (dir == 3 && c.ref == ref) { *px = c.x; *py = c.y; return; }
    if (b.avail || c.avail || !a.avail) {
        int ia = a.ref == ref, ib = b.ref == ref, ic = c.ref == ref;
        if (ia + ib + ic != 1) { *px = median3(a.x, b.x, c.x); *py = median3(a.y, b.y, c.y); }
        else if (ia) { *px = a.x; *py = a.y; }
        else if (ib) { *px = b.x; *py = b.y; }
        else { *px = c.x; *py = c.y; }
    } else { *px = a.x; *py = a.y; }
}
static void skip_mv(wr_t *w, const wmb_t *cur, int mbx, int mby, int *px, int *py)
{
    mvn_t a = mv_nb(w, cur, mbx, mby, -1, 0, 0), b = mv_nb(w, cur, mbx, mby, 0, -1, 0);
    if (!a.avail || !b.avail || (a.ref == 0 && a.x == 0 && a.y == 0) || (b.ref == 0 && b.x == 0 && b.y == 0)) { *px = *py = 0; return; }
    mv_pred(w, cur, mbx, mby, 0, 0, 4, 0, 0, 0, px, py);
}
static void set_mv(wmb_t *m, int x4, int y4, int w4, int h4, int mx, int my, unsigned *done)
{
    int i, j;
    for (j = y4; j < y4 + h4; j++) for (i = x4; i < x4 + w4; i++) {
        int blk = H264_RASTER_TO_BLK[j * 4 + i];
        m->mv[blk][0] = (int16_t)mx; m->mv[blk][1] = (int16_t)my; *done |= 1u << blk;
    }
}
static void draw_mv(wr_t *w, int mbx, int mby, int *mx, int *my)
{
    const h264w_params_t *p = w->p;
    int r = p->mv_range_qpel;
    if (rng_permille(&w->rng, p->far_mv_permille)) {
        /* far outside the picture: exercises the coordinate clamp (h264bsd_reconstruct.c:2222-2314) */
        int fx = (int)(w->W * 16 + 64) * 4, fy = (int)(w->H * 16 + 64) * 4;
        if (fx > 8000) fx = 8000;
        if (fy > 2000) fy = 2000;
        *mx = rng_range(&w->rng, -fx, fx); *my = rng_range(&w->rng, -fy, fy);
        return;
    }
    *mx = rng_range(&w->rng, -r, r); *my = rng_range(&w->rng, -r, r);
    (void)mbx; (void)mby;
}

/* ---------------------------------------------------- macroblock writers */
typedef struct {
    int luma[16][16];     /* scan order per luma4x4BlkIdx */
    int luma_dc[16];      /* I16x16 DC, scan order */
    int cdc[2][4];
    int cac[2][4][16];    /* scan order, index 0 unused */
    int cbp_luma, cbp_chroma;
    int has_luma_dc_block;
} resid_t;

static int qpc_of(wr_t *w, int qp)
{
    int q = qp + w->p->chroma_qp_index_offset;
    if (q < 0) q = 0;
    if (q > 51) q = 51;
    return H264_QPC[q];
}

/* Draw residual for a macroblock at QP qp. i16: Intra16x16 structure. */
static void draw_residual(wr_t *w, resid_t *r, int qp, int i16)
{
    const h264w_params_t *p = w->p;
    int b, pl, qpc = qpc_of(w, qp), dcs[16], k;
    memset(r, 0, sizeof *r);
    if (i16) {
        int ac = rng_permille(&w->rng, 500);
        if (rng_permille(&w->rng, 700)) {
            for (;;) {
                draw_block(w, r->luma_dc, 0);
                luma_dc_scaled(r->luma_dc, qp, dcs);
                for (k = 0; k < 16; k++) { int z[16]; memset(z, 0, sizeof z); if (!block_ok(z, 1, qp, 1, dcs[k])) break; }
                if (k == 16) break;
                shrink(r->luma_dc);
            }
        }
        luma_dc_scaled(r->luma_dc, qp, dcs);
        if (ac) {
            int any = 0;
            for (b = 0; b < 16; b++) {
                int dc = dcs[(H264_BLK_Y[b] >> 2) * 4 + (H264_BLK_X[b] >> 2)];
                if (!rng_permille(&w->rng, p->coded_blk_permille)) continue;
                draw_block(w, r->luma[b], 1);
                while (!block_ok(r->luma[b], 1, qp, 1, dc)) shrink(r->luma[b]);
                any |= count_nz(r->luma[b], 1) > 0;
            }
            /* cbp luma 15 may be signalled with all AC blocks empty; keep it random */
            r->cbp_luma = (any || rng_permille(&w->rng, 200)) ? 15 : 0;
            if (!r->cbp_luma) for (b = 0; b < 16; b++) memset(r->luma[b], 0, sizeof r->luma[b]);
        }
    } else {
        for (b = 0; b < 16; b++) {
            if (!rng_permille(&w->rng, p->coded_blk_permille)) continue;
            draw_block(w, r->luma[b], 0);
            while (!block_ok(r->luma[b], 0, qp, 0, 0)) shrink(r->luma[b]);
            if (count_nz(r->luma[b], 0)) r->cbp_luma |= 1 << (b >> 2);
        }
        /* sometimes signal an 8x8 as coded although all four blocks are empty */
        if (rng_permille(&w->rng, 30)) r->cbp_luma |= 1 << rng_u(&w->rng, 4);
    }
    /* chroma */
    {
        int mode = rng_permille(&w->rng, p->coded_blk_permille) ? (rng_permille(&w->rng, 500) ? 2 : 1) : 0;
        int dcv[2][4];
        r->cbp_chroma = mode;
        for (pl = 0; pl < 2 && mode; pl++) {
            if (rng_permille(&w->rng, 700)) {
                int s[16], z[16];
                for (;;) {
                    draw_block(w, s, 12);   /* 4 entries at 12..15 */
                    for (k = 0; k < 4; k++) r->cdc[pl][k] = s[12 + k];
                    chroma_dc_scaled(r->cdc[pl], qpc, dcv[pl]);
                    memset(z, 0, sizeof z);
                    for (k = 0; k < 4; k++) if (!block_ok(z, 1, qpc, 1, dcv[pl][k])) break;
                    if (k == 4) break;
                    for (k = 0; k < 4; k++) s[12 + k] /= 2;
                    if (!(s[12] | s[13] | s[14] | s[15])) { memset(r->cdc[pl], 0, sizeof r->cdc[pl]); break; }
                }
            }
            chroma_dc_scaled(r->cdc[pl], qpc, dcv[pl]);
            if (mode == 2) for (b = 0; b < 4; b++) {
                if (!rng_permille(&w->rng, p->coded_blk_permille)) continue;
                draw_block(w, r->cac[pl][b], 1);
                while (!block_ok(r->cac[pl][b], 1, qpc, 1, dcv[pl][b])) shrink(r->cac[pl][b]);
            }
        }
    }
}

/* write residual() syntax (7.3.5.3) and record TotalCoeff; assumes cur->tc zeroed */
static void write_residual(wr_t *w, bitw_t *b, wmb_t *cur, int mbx, int mby, const resid_t *r, int i16)
{
    int i8, i4, pl, c;
    if (i16) write_block(b, r->luma_dc, 16, luma_nc(w, cur, mbx, mby, 0));
    for (i8 = 0; i8 < 4; i8++) for (i4 = 0; i4 < 4; i4++) {
        int blk = i8 * 4 + i4;
        if (!((r->cbp_luma >> i8) & 1)) { cur->tc[blk] = 0; continue; }
        if (i16) cur->tc[blk] = (uint8_t)write_block(b, r->luma[blk] + 1, 15, luma_nc(w, cur, mbx, mby, blk));
        else     cur->tc[blk] = (uint8_t)write_block(b, r->luma[blk], 16, luma_nc(w, cur, mbx, mby, blk));
    }
    if (r->cbp_chroma) for (pl = 0; pl < 2; pl++) write_block(b, r->cdc[pl], 4, -1);
    if (r->cbp_chroma == 2) for (pl = 0; pl < 2; pl++) for (c = 0; c < 4; c++)
        cur->tc[16 + 4 * pl + c] = (uint8_t)write_block(b, r->cac[pl][c] + 1, 15, chroma_nc(w, cur, mbx, mby, pl, c));
}

/* choose mb_qp_delta: wander inside [qp-jitter, qp+jitter] */
static int draw_qp_delta(wr_t *w)
{
    const h264w_params_t *p = w->p;
    int lo, hi, t;
    if (p->qp_jitter <= 0 || !rng_permille(&w->rng, 150)) return 0;
    lo = p->qp - p->qp_jitter; hi = p->qp + p->qp_jitter;
    if (lo < 0) lo = 0;
    if (hi > 51) hi = 51;
    t = rng_range(&w->rng, lo, hi);
    return t - w->qp;
}

static void write_intra_mb(wr_t *w, bitw_t *b, wmb_t *cur, int mbx, int mby, int mb_type_offset, int force_pcm)
{
    const h264w_params_t *p = w->p;
    int aA = intra_avail(w, mbx - 1, mby), aB = intra_avail(w, mbx, mby - 1), aD = intra_avail(w, mbx - 1, mby - 1);
    int chroma_mode, blk, k, dqp;
    resid_t r;

    if (force_pcm || rng_permille(&w->rng, p->ipcm_permille)) {
        cur->kind = K_PCM;
        bw_ue(b, (uint32_t)(mb_type_offset + 25));
        bw_align_zero(b);
        for (k = 0; k < 384; k++) bw_put(b, 8, rng_u(&w->rng, 256));
        memset(cur->tc, 16, sizeof cur->tc);
        return;      /* QP_Y unchanged for later macroblocks; deblocking sees qp 0 */
    }
    /* chroma mode: 0 DC, 1 horizontal (left), 2 vertical (up), 3 plane (all) */
    {
        int legal[4], n = 0;
        legal[n++] = 0;
        if (aA) legal[n++] = 1;
        if (aB) legal[n++] = 2;
        if (aA && aB && aD) legal[n++] = 3;
        chroma_mode = legal[rng_u(&w->rng, (uint32_t)n)];
    }
    if (rng_permille(&w->rng, p->i16_permille)) {
        int legal[4], n = 0, mode;
        cur->kind = K_I16;
        if (aB) legal[n++] = 0;
        if (aA) legal[n++] = 1;
        legal[n++] = 2;
        if (aA && aB && aD) legal[n++] = 3;
        mode = legal[rng_u(&w->rng, (uint32_t)n)];
        dqp = draw_qp_delta(w);
        w->qp += dqp;
        draw_residual(w, &r, w->qp, 1);
        bw_ue(b, (uint32_t)(mb_type_offset + 1 + mode + 4 * r.cbp_chroma + (r.cbp_luma ? 12 : 0)));
        bw_ue(b, (uint32_t)chroma_mode);
        bw_se(b, dqp);
        write_residual(w, b, cur, mbx, mby, &r, 1);
        return;
    }
    cur->kind = K_I4;
    bw_ue(b, (uint32_t)mb_type_offset);
    for (blk = 0; blk < 16; blk++) {
        int x4 = H264_BLK_X[blk] >> 2, y4 = H264_BLK_Y[blk] >> 2;
        int left = x4 > 0 ? 1 : aA, up = y4 > 0 ? 1 : aB;
        int ul = (x4 > 0 && y4 > 0) ? 1 : x4 > 0 ? aB : y4 > 0 ? aA : aD;
        int legal[9], n = 0, mode, pred;
        if (up) legal[n++] = 0;
        if (left) legal[n++] = 1;
        legal[n++] = 2;
        if (up) { legal[n++] = 3; legal[n++] = 7; }
        if (up && left && ul) { legal[n++] = 4; legal[n++] = 5; legal[n++] = 6; }
        if (left) legal[n++] = 8;
        mode = legal[rng_u(&w->rng, (uint32_t)n)];
        pred = pred_i4mode(w, cur, mbx, mby, blk);
        /* make prev_intra4x4_pred_mode_flag=1 common enough to be exercised */
        if (rng_permille(&w->rng, 300)) { for (k = 0; k < n; k++) if (legal[k] == pred) mode = pred; }
        cur->i4mode[blk] = (uint8_t)mode;
        if (mode == pred) bw_put(b, 1, 1);
        else { bw_put(b, 1, 0); bw_put(b, 3, (uint32_t)(mode < pred ? mode : mode - 1)); }
    }
    bw_ue(b, (uint32_t)chroma_mode);
    {
        int qp_try = w->qp, cbp, code;
        dqp = draw_qp_delta(w);
        draw_residual(w, &r, w->qp + dqp, 0);
        cbp = r.cbp_luma | (r.cbp_chroma << 4);
        for (code = 0; code < 48; code++) if (H264_CBP_MAP[code][0] == cbp) break;
        bw_ue(b, (uint32_t)code);
        if (cbp) { bw_se(b, dqp); w->qp = qp_try + dqp; write_residual(w, b, cur, mbx, mby, &r, 0); }
        else {
            /* no residual => no mb_qp_delta; the draw above used qp+dqp only for range checks of nothing */
        }
    }
}

static void write_inter_mb(wr_t *w, bitw_t *b, wmb_t *cur, int mbx, int mby)
{
    const h264w_params_t *p = w->p;
    unsigned done = 0;
    int shape, k, nref = w->num_ref_active, dqp, cbp, code;
    resid_t r;
    cur->kind = K_INTER;
    shape = p->part_mix ? (int)rng_u(&w->rng, 100) : 0;
    /* 40% 16x16, 17% 16x8, 17% 8x16, 26% 8x8 (of the non-skipped inter macroblocks) */
    if (shape < 40) {
        int ref = (int)rng_u(&w->rng, (uint32_t)nref), mx, my, px, py;
        bw_ue(b, 0);
        if (nref > 1) { if (nref == 2) bw_put(b, 1, (uint32_t)!ref); else bw_ue(b, (uint32_t)ref); }
        for (k = 0; k < 4; k++) cur->ref[k] = (int8_t)ref;
        mv_pred(w, cur, mbx, mby, 0, 0, 4, ref, done, 0, &px, &py);
        draw_mv(w, mbx, mby, &mx, &my);
        bw_se(b, mx - px); bw_se(b, my - py);
        set_mv(cur, 0, 0, 4, 4, mx, my, &done);
    } else if (shape < 74) {
        int hor = shape < 57;   /* 16x8 : 8x16 */
        int ref[2], mx[2], my[2], px, py, i;
        bw_ue(b, hor ? 1u : 2u);
        for (i = 0; i < 2; i++) {
            ref[i] = (int)rng_u(&w->rng, (uint32_t)nref);
            if (nref > 1) { if (nref == 2) bw_put(b, 1, (uint32_t)!ref[i]); else bw_ue(b, (uint32_t)ref[i]); }
        }
        if (hor) { cur->ref[0] = cur->ref[1] = (int8_t)ref[0]; cur->ref[2] = cur->ref[3] = (int8_t)ref[1]; }
        else     { cur->ref[0] = cur->ref[2] = (int8_t)ref[0]; cur->ref[1] = cur->ref[3] = (int8_t)ref[1]; }
        for (i = 0; i < 2; i++) {
            if (hor) mv_pred(w, cur, mbx, mby, 0, 2 * i, 4, ref[i], done, i == 0 ? 2 : 1, &px, &py);
            else     mv_pred(w, cur, mbx, mby, 2 * i, 0, 2, ref[i], done, i == 0 ? 1 : 3, &px, &py);
            draw_mv(w, mbx, mby, &mx[i], &my[i]);
            bw_se(b, mx[i] - px); bw_se(b, my[i] - py);
            if (hor) set_mv(cur, 0, 2 * i, 4, 2, mx[i], my[i], &done);
            else     set_mv(cur, 2 * i, 0, 2, 4, mx[i], my[i], &done);
        }
    } else {
        int ref0 = nref > 1 ? rng_permille(&w->rng, 300) : rng_permille(&w->rng, 500);   /* P_8x8ref0 */
        int sub[4], ref[4], i, s;
        /* all mvds are written after all sub types and ref indices: keep them in a side buffer */
        int mvd[16][2], nmvd = 0;
        bw_ue(b, ref0 ? 4u : 3u);
        for (i = 0; i < 4; i++) { sub[i] = (int)rng_u(&w->rng, 4); bw_ue(b, (uint32_t)sub[i]); }
        for (i = 0; i < 4; i++) {
            ref[i] = ref0 ? 0 : (int)rng_u(&w->rng, (uint32_t)nref);
            if (!ref0 && nref > 1) { if (nref == 2) bw_put(b, 1, (uint32_t)!ref[i]); else bw_ue(b, (uint32_t)ref[i]); }
            cur->ref[i] = (int8_t)ref[i];
        }
        for (i = 0; i < 4; i++) {
            int ox = (i & 1) * 2, oy = (i >> 1) * 2;
            int nsub = sub[i] == 0 ? 1 : sub[i] == 3 ? 4 : 2;
            for (s = 0; s < nsub; s++) {
                int x4, y4, w4, h4, mx, my, px, py;
                switch (sub[i]) {
                case 0: x4 = ox; y4 = oy; w4 = 2; h4 = 2; break;
                case 1: x4 = ox; y4 = oy + s; w4 = 2; h4 = 1; break;          /* 8x4 */
                case 2: x4 = ox + s; y4 = oy; w4 = 1; h4 = 2; break;          /* 4x8 */
                default: x4 = ox + (s & 1); y4 = oy + (s >> 1); w4 = 1; h4 = 1; break;
                }
                mv_pred(w, cur, mbx, mby, x4, y4, w4, ref[i], done, 0, &px, &py);
                draw_mv(w, mbx, mby, &mx, &my);
                /* neighbouring sub-blocks often share a vector in real streams */
                if (s > 0 && rng_permille(&w->rng, 300)) { mx = px; my = py; }
                mvd[nmvd][0] = mx - px; mvd[nmvd][1] = my - py; nmvd++;
                set_mv(cur, x4, y4, w4, h4, mx, my, &done);
            }
        }
        for (i = 0; i < nmvd; i++) { bw_se(b, mvd[i][0]); bw_se(b, mvd[i][1]); }
    }
    dqp = draw_qp_delta(w);
    draw_residual(w, &r, w->qp + dqp, 0);
    cbp = r.cbp_luma | (r.cbp_chroma << 4);
    for (code = 0; code < 48; code++) if (H264_CBP_MAP[code][1] == cbp) break;
    bw_ue(b, (uint32_t)code);
    if (cbp) { bw_se(b, dqp); w->qp += dqp; write_residual(w, b, cur, mbx, mby, &r, 0); }
}

/* ------------------------------------------------------- stream assembly */
typedef struct { int idc, a, bq, qp; } slice_par_t;
/* per-picture reference handling (dpb_stress): nal_ref_idc, list reordering, adaptive marking */
typedef struct { int ref_idc; int reorder_diff_minus1; int mmco1_diff_minus1; int make_long; int drop_long; } pic_par_t;

/* mbs[0..n_mbs): macroblock addresses of the slice in decoding order (ascending; contiguous without FMO) */
static size_t write_slice(wr_t *w, uint8_t *out, size_t cap, uint8_t *scratch, size_t scratch_cap,
                          const uint32_t *mbs, uint32_t n_mbs, int idr, int is_p, uint32_t frame_num,
                          uint32_t idr_pic_id, uint32_t poc_lsb, const slice_par_t *sp, const pic_par_t *pp,
                          int cycle_bits, uint32_t change_cycle)
{
    const h264w_params_t *p = w->p;
    bitw_t b; uint32_t i, skip_run = 0;
    bw_init(&b, scratch, scratch_cap);
    bw_ue(&b, mbs[0]);
    bw_ue(&b, is_p ? 0u : 2u);
    bw_ue(&b, 0);                                    /* pic_parameter_set_id */
    bw_put(&b, 4, frame_num & 15);                   /* log2_max_frame_num = 4 */
    if (idr) bw_ue(&b, idr_pic_id);
    if (p->poc_type == 0) bw_put(&b, 8, poc_lsb & 255);
    if (is_p) {
        int def = (int)p->num_ref_frames;            /* PPS default = num_ref_frames */
        if (w->num_ref_active != def) { bw_put(&b, 1, 1); bw_ue(&b, (uint32_t)(w->num_ref_active - 1)); }
        else bw_put(&b, 1, 0);
        if (pp->reorder_diff_minus1 >= 0) {          /* move an older short-term picture to index 0 (8.2.4.3.1) */
            bw_put(&b, 1, 1);                        /* ref_pic_list_reordering_flag_l0 */
            bw_ue(&b, 0); bw_ue(&b, (uint32_t)pp->reorder_diff_minus1);   /* reordering_of_pic_nums_idc 0, abs_diff_pic_num_minus1 */
            bw_ue(&b, 3);                            /* end */
        } else bw_put(&b, 1, 0);
    }
    /* dec_ref_pic_marking: only in reference pictures (7.3.3) */
    if (pp->ref_idc) {
        if (idr) { bw_put(&b, 1, 0); bw_put(&b, 1, 0); }
        else if (pp->mmco1_diff_minus1 >= 0 || pp->make_long || pp->drop_long) {   /* adaptive marking (8.2.5.4) */
            bw_put(&b, 1, 1);
            if (pp->mmco1_diff_minus1 >= 0) { bw_ue(&b, 1); bw_ue(&b, (uint32_t)pp->mmco1_diff_minus1); }   /* short-term -> unused */
            if (pp->drop_long) { bw_ue(&b, 2); bw_ue(&b, 0); }                                               /* LongTermPicNum 0 -> unused */
            if (pp->make_long) { bw_ue(&b, 4); bw_ue(&b, 1); bw_ue(&b, 6); bw_ue(&b, 0); }                   /* max idx + 1 = 1; current -> long-term idx 0 */
            bw_ue(&b, 0);
        } else bw_put(&b, 1, 0);
    }
    bw_se(&b, sp->qp - p->qp);                       /* slice_qp_delta vs pic_init_qp */
    /* deblocking_filter_control_present_flag = 1 */
    bw_ue(&b, (uint32_t)sp->idc);
    if (sp->idc != 1) { bw_se(&b, sp->a); bw_se(&b, sp->bq); }
    if (cycle_bits >= 0) bw_put(&b, cycle_bits, change_cycle);   /* slice_group_change_cycle (map types 3..5) */

    w->qp = sp->qp;
    w->is_p = is_p;
    for (i = 0; i < n_mbs; i++) {
        uint32_t addr = mbs[i];
        int mbx = (int)(addr % w->W), mby = (int)(addr / w->W);
        wmb_t *cur = &w->mb[addr];
        memset(cur, 0, sizeof *cur);
        cur->slice = (uint16_t)w->cur_slice;
        w->cur_addr = addr;
        if (is_p && rng_permille(&w->rng, p->p_skip_permille)) {
            int mx, my, k;
            cur->kind = K_INTER; cur->skip = 1;
            skip_mv(w, cur, mbx, mby, &mx, &my);
            for (k = 0; k < 16; k++) { cur->mv[k][0] = (int16_t)mx; cur->mv[k][1] = (int16_t)my; }
            skip_run++;
            continue;
        }
        if (is_p) { bw_ue(&b, skip_run); skip_run = 0; }
        if (!is_p || rng_permille(&w->rng, p->p_intra_permille))
            write_intra_mb(w, &b, cur, mbx, mby, is_p ? 5 : 0, p->first_idr_ipcm && idr);
        else
            write_inter_mb(w, &b, cur, mbx, mby);
    }
    if (is_p && skip_run) bw_ue(&b, skip_run);
    bw_trailing(&b);
    if (b.ovf) return 0;
    return emit_nal(out, cap, pp->ref_idc, idr ? 5 : 1, scratch, b.pos);
}

void h264w_default_params(h264w_params_t *p, uint32_t width_mbs, uint32_t height_mbs, uint32_t n_frames)
{
    memset(p, 0, sizeof *p);
    p->width_mbs = width_mbs; p->height_mbs = height_mbs; p->n_frames = n_frames;
    p->idr_period = 0; p->seed = 1234; p->qp = 30; p->qp_jitter = 4;
    p->coded_blk_permille = 300; p->max_coeffs = 4; p->max_level = 8;
    p->num_ref_frames = 1; p->slices_per_pic = 1; p->poc_type = 2;
    p->deblock_idc = 0; p->p_intra_permille = 30; p->p_skip_permille = 100;
    p->ipcm_permille = 2; p->i16_permille = 500; p->mv_range_qpel = 128; p->far_mv_permille = 2;
    p->level_idc = (width_mbs * height_mbs > 8704) ? 51 : 40; p->part_mix = 1;
}

size_t h264w_bound(const h264w_params_t *p)
{
    /* worst case is all I_PCM (384 B/MB plus emulation prevention); coded MBs stay well below */
    return (size_t)p->n_frames * p->width_mbs * p->height_mbs * 1536 + 4096;
}

size_t h264w_generate(const h264w_params_t *p, uint8_t *out, size_t cap)
{
    wr_t w; bitw_t b; uint8_t hdr[64]; uint8_t *scratch; size_t scratch_cap, o = 0, n;
    uint32_t f, since_idr = 0, idr_id = 0, s, n_short = 0, n_long = 0, prev_ref_fn = 0, ref_ord = 0;
    uint32_t short_ord[17];                          /* ordinal (count of reference pictures since the IDR) of each short-term picture, most recent first */
    h264_fmo_t fmo; uint8_t *group_ids = NULL, *map = NULL; uint32_t *order = NULL;
    if (!p || !out || !p->width_mbs || !p->height_mbs || !p->n_frames) return 0;
    if (p->num_ref_frames < 1 || p->num_ref_frames > 16 || (p->poc_type != 0 && p->poc_type != 2)) return 0;
    if (p->qp < 0 || p->qp > 51 || p->deblock_idc > 2) return 0;
    memset(&w, 0, sizeof w);
    w.p = p; w.W = p->width_mbs; w.H = p->height_mbs; w.nmb = w.W * w.H;
    w.rng.s = p->seed * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    if (!w.rng.s) w.rng.s = 1;
    w.mb = (wmb_t *)calloc(w.nmb, sizeof(wmb_t));
    scratch_cap = (size_t)w.nmb * 1536 + 4096;
    scratch = (uint8_t *)malloc(scratch_cap);
    if (!w.mb || !scratch) { free(w.mb); free(scratch); return 0; }

    /* SPS (7.3.2.1) */
    bw_init(&b, hdr, sizeof hdr);
    bw_put(&b, 8, 66);                 /* profile_idc: Baseline */
    bw_put(&b, 8, 0xC0);               /* constraint_set0/1 flags, reserved zero */
    bw_put(&b, 8, p->level_idc);
    bw_ue(&b, 0);                      /* seq_parameter_set_id */
    bw_ue(&b, 0);                      /* log2_max_frame_num_minus4 */
    bw_ue(&b, p->poc_type);
    if (p->poc_type == 0) bw_ue(&b, 4);/* log2_max_pic_order_cnt_lsb_minus4 -> 8 bits */
    bw_ue(&b, p->num_ref_frames);
    bw_put(&b, 1, 0);                  /* gaps_in_frame_num_value_allowed_flag */
    bw_ue(&b, w.W - 1);
    bw_ue(&b, w.H - 1);
    bw_put(&b, 1, 1);                  /* frame_mbs_only_flag */
    bw_put(&b, 1, 1);                  /* direct_8x8_inference_flag */
    if (p->crop) { bw_put(&b, 1, 1); bw_ue(&b, 0); bw_ue(&b, 0); bw_ue(&b, 0); bw_ue(&b, 4); }
    else bw_put(&b, 1, 0);
    bw_put(&b, 1, 0);                  /* vui_parameters_present_flag */
    bw_trailing(&b);
    n = emit_nal(out + o, cap - o, 1, 7, hdr, b.pos); if (!n) goto fail; o += n;
    /* PPS (7.3.2.2) */
    bw_init(&b, scratch, scratch_cap);
    bw_ue(&b, 0); bw_ue(&b, 0);
    bw_put(&b, 1, 0);                  /* entropy_coding_mode_flag: CAVLC */
    bw_put(&b, 1, 0);                  /* pic_order_present_flag */
    if (p->fmo_type) {                 /* flexible macroblock ordering: slice_group_map_type = fmo_type - 1 */
        uint32_t i, bits = 0;
        memset(&fmo, 0, sizeof fmo);
        fmo.n_groups = p->fmo_groups < 2 ? 2 : p->fmo_groups > 8 ? 8 : p->fmo_groups;
        fmo.type = p->fmo_type - 1;
        if (fmo.type >= 3 && fmo.type <= 5) fmo.n_groups = 2;
        bw_ue(&b, fmo.n_groups - 1);
        bw_ue(&b, fmo.type);
        if (fmo.type == 0) {
            for (i = 0; i < fmo.n_groups; i++) { fmo.run_length[i] = 1 + rng_u(&w.rng, 2 * w.W); bw_ue(&b, fmo.run_length[i] - 1); }
        } else if (fmo.type == 2) {
            for (i = 0; i + 1 < fmo.n_groups; i++) {
                uint32_t x0 = rng_u(&w.rng, w.W), x1 = x0 + rng_u(&w.rng, w.W - x0), y0 = rng_u(&w.rng, w.H), y1 = y0 + rng_u(&w.rng, w.H - y0);
                fmo.top_left[i] = y0 * w.W + x0; fmo.bottom_right[i] = y1 * w.W + x1;
                bw_ue(&b, fmo.top_left[i]); bw_ue(&b, fmo.bottom_right[i]);
            }
        } else if (fmo.type >= 3 && fmo.type <= 5) {
            fmo.change_direction = rng_u(&w.rng, 2);
            fmo.change_rate = 1 + rng_u(&w.rng, w.W);
            bw_put(&b, 1, fmo.change_direction); bw_ue(&b, fmo.change_rate - 1);
        } else if (fmo.type == 6) {
            group_ids = (uint8_t *)malloc(w.nmb);
            if (!group_ids) goto fail;
            while ((1u << bits) < fmo.n_groups) bits++;
            bw_ue(&b, w.nmb - 1);
            for (i = 0; i < w.nmb; i++) { group_ids[i] = (uint8_t)rng_u(&w.rng, fmo.n_groups); bw_put(&b, (int)bits, group_ids[i]); }
            fmo.group_id = group_ids;
        }
    } else bw_ue(&b, 0);               /* num_slice_groups_minus1 */
    bw_ue(&b, p->num_ref_frames - 1);  /* num_ref_idx_l0_default_active_minus1 */
    bw_ue(&b, 0);
    bw_put(&b, 1, 0); bw_put(&b, 2, 0);/* weighted_pred_flag, weighted_bipred_idc */
    bw_se(&b, p->qp - 26);
    bw_se(&b, 0);
    bw_se(&b, p->chroma_qp_index_offset);
    bw_put(&b, 1, 1);                  /* deblocking_filter_control_present_flag */
    bw_put(&b, 1, p->constrained_intra_pred ? 1 : 0);
    bw_put(&b, 1, 0);                  /* redundant_pic_cnt_present_flag */
    bw_trailing(&b);
    if (b.ovf) goto fail;
    {   /* the RBSP sits in `scratch`, which emit_nal only reads */
        uint8_t *tmp = (uint8_t *)malloc(b.pos + 1);
        if (!tmp) goto fail;
        memcpy(tmp, scratch, b.pos);
        n = emit_nal(out + o, cap - o, 1, 8, tmp, b.pos);
        free(tmp);
        if (!n) goto fail;
        o += n;
    }
    order = (uint32_t *)malloc(w.nmb * sizeof(uint32_t));
    map = (uint8_t *)malloc(w.nmb);
    if (!order || !map) goto fail;

    for (f = 0; f < p->n_frames; f++) {
        int idr = f == 0 || (p->idr_period && f % p->idr_period == 0);
        int is_p = !idr && !p->intra_only;
        uint32_t nsl = p->slices_per_pic ? p->slices_per_pic : 1, first = 0;
        pic_par_t pp;
        uint32_t frame_num;
        if (idr) { since_idr = 0; idr_id++; n_short = 0; n_long = 0; ref_ord = 0; }
        if (nsl > w.nmb) nsl = w.nmb;
        /* frame_num: 0 at an IDR, else one more than the last REFERENCE picture's (7.4.3) */
        frame_num = idr ? 0 : prev_ref_fn + 1;
        pp.ref_idc = 1; pp.reorder_diff_minus1 = -1; pp.mmco1_diff_minus1 = -1; pp.make_long = 0; pp.drop_long = 0;
        if (p->dpb_stress && !idr) {
            /* PicNum of a short-term picture = CurrPicNum - (ref_ord - its ordinal) as long as fewer than 16 reference
             * pictures lie between (frame_num has 4 bits) */
            /* every third picture is a non-reference picture (never two in a row: POC type 2 forbids it) */
            if (since_idr % 3 == 2) pp.ref_idc = 0;
            if (is_p && n_short >= 2 && since_idr % 2 == 0) pp.reorder_diff_minus1 = (int)(ref_ord - short_ord[1]) - 1;
            if (pp.ref_idc) {
                if (p->dpb_stress > 1 && since_idr % 8 == 3 && !n_long) pp.make_long = 1;
                else if (p->dpb_stress > 1 && since_idr % 8 == 7 && n_long) pp.drop_long = 1;
                /* a reference picture sometimes removes the oldest short-term picture itself instead of the sliding window;
                 * it must when adaptive marking is on and the buffer is full */
                if (n_short >= 2 && (since_idr % 4 == 1 || ((pp.make_long || pp.drop_long) && n_short + n_long - (uint32_t)pp.drop_long >= p->num_ref_frames)))
                    pp.mmco1_diff_minus1 = (int)(ref_ord - short_ord[n_short - 1]) - 1;
                if ((pp.make_long || pp.drop_long) && pp.mmco1_diff_minus1 < 0 && n_short + n_long - (uint32_t)pp.drop_long >= p->num_ref_frames) { pp.make_long = 0; pp.drop_long = 0; }
            }
        }
        w.num_ref_active = (int)(n_short + n_long < p->num_ref_frames ? n_short + n_long : p->num_ref_frames);
        if (w.num_ref_active < 1) w.num_ref_active = 1;
        {   /* macroblocks of a slice are stamped with its id as they are written: forget the previous picture's */
            uint32_t a;
            for (a = 0; a < w.nmb; a++) w.mb[a].slice = 0;
        }
        if (p->fmo_type) {
            /* one or more slices per slice group, each listing the group's macroblocks in ascending order */
            uint32_t cycle = 0, units0 = 0, g, a;
            int cycle_bits = -1;
            if (fmo.type >= 3 && fmo.type <= 5) {
                const uint32_t max_cycle = (w.nmb + fmo.change_rate - 1) / fmo.change_rate;
                cycle_bits = (int)h264_fmo_cycle_bits(w.nmb, fmo.change_rate);
                cycle = rng_u(&w.rng, max_cycle + 1);
                units0 = cycle * fmo.change_rate; if (units0 > w.nmb) units0 = w.nmb;
            }
            h264_fmo_build_map(map, w.W, w.H, &fmo, units0);
            s = 0;
            for (g = 0; g < fmo.n_groups; g++) {
                uint32_t cnt = 0, parts, part;
                for (a = 0; a < w.nmb; a++) if (map[a] == g) order[cnt++] = a;
                if (!cnt) continue;
                parts = nsl < cnt ? nsl : cnt;
                for (part = 0; part < parts; part++) {
                    const uint32_t lo = (cnt * part) / parts, hi = (cnt * (part + 1)) / parts;
                    slice_par_t sp;
                    sp.idc = (int)p->deblock_idc; sp.a = p->alpha_c0_offset_div2; sp.bq = p->beta_offset_div2; sp.qp = p->qp;
                    if (p->multi_slice_params) {
                        sp.idc = (int)rng_u(&w.rng, 3); sp.a = rng_range(&w.rng, -3, 3); sp.bq = rng_range(&w.rng, -3, 3);
                        sp.qp = p->qp + rng_range(&w.rng, -2, 2);
                        if (sp.qp < 0) sp.qp = 0;
                        if (sp.qp > 51) sp.qp = 51;
                    }
                    w.cur_slice = ++s;
                    n = write_slice(&w, out + o, cap - o, scratch, scratch_cap, order + lo, hi - lo, idr, is_p,
                                    frame_num, idr_id & 0xffff, since_idr * 2, &sp, &pp, cycle_bits, cycle);
                    if (!n) goto fail;
                    o += n;
                }
            }
        } else
        for (s = 0; s < nsl; s++) {
            uint32_t cnt = (w.nmb * (s + 1)) / nsl - (w.nmb * s) / nsl, a;
            slice_par_t sp;
            sp.idc = (int)p->deblock_idc; sp.a = p->alpha_c0_offset_div2; sp.bq = p->beta_offset_div2; sp.qp = p->qp;
            if (p->multi_slice_params) {
                sp.idc = (int)rng_u(&w.rng, 3); sp.a = rng_range(&w.rng, -3, 3); sp.bq = rng_range(&w.rng, -3, 3);
                sp.qp = p->qp + rng_range(&w.rng, -2, 2);
                if (sp.qp < 0) sp.qp = 0;
                if (sp.qp > 51) sp.qp = 51;
            }
            for (a = 0; a < cnt; a++) order[a] = first + a;
            w.cur_slice = s + 1;
            n = write_slice(&w, out + o, cap - o, scratch, scratch_cap, order, cnt, idr, is_p,
                            frame_num, idr_id & 0xffff, since_idr * 2, &sp, &pp, -1, 0);
            if (!n) goto fail;
            o += n; first += cnt;
        }
        if (pp.ref_idc) {                            /* decoded reference picture marking as the decoder will do it */
            uint32_t k;
            if (idr) { n_short = 0; n_long = 0; }
            else if (pp.mmco1_diff_minus1 >= 0 || pp.make_long || pp.drop_long) {
                if (pp.mmco1_diff_minus1 >= 0) n_short--;                    /* the oldest one */
                if (pp.drop_long) n_long = 0;
            } else if (n_short + n_long >= p->num_ref_frames && n_short) n_short--;   /* sliding window */
            if (pp.make_long) n_long = 1;
            else { for (k = n_short; k > 0; k--) short_ord[k] = short_ord[k - 1]; short_ord[0] = ref_ord; n_short++; }
            ref_ord++;
            prev_ref_fn = frame_num;
        }
        since_idr++;
    }
    free(w.mb); free(scratch); free(group_ids); free(map); free(order);
    return o;
fail:
    free(w.mb); free(scratch); free(group_ids); free(map); free(order);
    return 0;
}
