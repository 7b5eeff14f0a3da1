/* h264_slice.c — slice_data() / macroblock_layer() parsing into GPU records.
 *
 * Host half of the reference's per-macroblock loop (h264bsd_slice_data.c:85-235
 * h264bsdDecodeSliceData; h264bsd_macroblock_layer.c:133-242 macroblock_layer,
 * :353-496 mb_pred/sub_mb_pred, :699-869 residual + nC; h264bsd_inter_prediction.c
 * :499-1031 motion vector prediction; h264bsd_intra_prediction.c:1885-1936
 * Intra4x4PredMode derivation; h264bsd_macroblock_layer.c:1043-1049 QP update).
 * Everything after the QP update (ProcessResidual, intra/inter reconstruction)
 * is NOT done here: this file only fills one h264b200_mb_t per macroblock plus
 * coefficient slots (include/h264b200_records.h) for the CUDA kernels.
 *
 * Design differences from the reference: no 2 KB macroblockLayer_t memset per
 * macroblock (:152), levels are written once, already de-zig-zagged, as int16
 * into the slot array; neighbour state is a 32-byte context per macroblock.
 */
#include <string.h>
#include "h264_internal.h"
#include "h264_consts.h"
#include "h264_cavlc_inl.h"

typedef struct {
    h264_decoder_t *d;
    br_t *b;
    const h264_slice_hdr_t *sh;
    h264_pic_input_t *pic;
    uint32_t W, H;
    uint16_t slice_id;
    int is_p, constrained_intra, chroma_qp_off;
    int qp;
    int ref_slot[H264_MAX_REFS + 1];
    /* current macroblock */
    uint32_t addr; int mbx, mby;
    h264_mbctx_t *cur; h264b200_mb_t *rec;
    const h264_mbctx_t *cA, *cB, *cC, *cD;     /* NULL when unavailable (other slice / outside / not decoded) */
    const h264b200_mb_t *rA, *rB, *rC, *rD;
} sl_t;

static void set_neighbours(sl_t *s)
{
    h264_mbctx_t *ctx = s->d->mbctx; h264b200_mb_t *recs = s->pic->mbs;
    uint32_t a = s->addr, W = s->W;
    s->cA = s->cB = s->cC = s->cD = NULL; s->rA = s->rB = s->rC = s->rD = NULL;
    if (s->mbx > 0 && ctx[a - 1].slice_id == s->slice_id) { s->cA = &ctx[a - 1]; s->rA = &recs[a - 1]; }
    if (s->mby > 0) {
        if (ctx[a - W].slice_id == s->slice_id) { s->cB = &ctx[a - W]; s->rB = &recs[a - W]; }
        if (s->mbx + 1 < (int)W && ctx[a - W + 1].slice_id == s->slice_id) { s->cC = &ctx[a - W + 1]; s->rC = &recs[a - W + 1]; }
        if (s->mbx > 0 && ctx[a - W - 1].slice_id == s->slice_id) { s->cD = &ctx[a - W - 1]; s->rD = &recs[a - W - 1]; }
    }
}

/* ------------------------------------------------ motion vector prediction */
typedef struct { int avail, ref; int x, y; } mvn_t;

static inline mvn_t mvn_from(const h264_mbctx_t *c, const h264b200_mb_t *r, int x4, int y4)
{
    mvn_t n; n.avail = 0; n.ref = -1; n.x = n.y = 0;
    if (!c) return n;
    n.avail = 1;
    if (c->kind == H264B200_MB_INTER) {
        n.ref = c->ref_idx[(y4 >> 1) * 2 + (x4 >> 1)];
        n.x = r->mv[y4 * 4 + x4][0]; n.y = r->mv[y4 * 4 + x4][1];
    }
    return n;
}
/* neighbour 4x4 block at (x4,y4) relative to the current macroblock; `done`: raster bit mask of
 * current-macroblock blocks whose vectors are already derived */
static inline mvn_t mvn_at(const sl_t *s, int x4, int y4, unsigned done)
{
    mvn_t n; n.avail = 0; n.ref = -1; n.x = n.y = 0;
    if (y4 < 0) {
        if (x4 < 0) return mvn_from(s->cD, s->rD, 3, 3);
        if (x4 > 3) return mvn_from(s->cC, s->rC, x4 - 4, 3);
        return mvn_from(s->cB, s->rB, x4, 3);
    }
    if (x4 < 0) return mvn_from(s->cA, s->rA, 3, y4);
    if (x4 > 3) return n;
    if (!((done >> (y4 * 4 + x4)) & 1)) return n;
    n.avail = 1;
    n.ref = s->cur->ref_idx[(y4 >> 1) * 2 + (x4 >> 1)];
    n.x = s->rec->mv[y4 * 4 + x4][0]; n.y = s->rec->mv[y4 * 4 + x4][1];
    return n;
}
static inline int median3(int a, int b, int c) { int mx = a > b ? a : b, mn = a < b ? a : b; return c > mx ? mx : c < mn ? mn : c; }

/* dir: 0 median, 1 A first (8x16 left / 16x8 bottom), 2 B first (16x8 top), 3 C first (8x16 right) */
static void predict_mv(const sl_t *s, int x4, int y4, int w4, int ref, unsigned done, int dir, int *px, int *py)
{
    mvn_t a = mvn_at(s, x4 - 1, y4, done), b = mvn_at(s, x4, y4 - 1, done), c = mvn_at(s, x4 + w4, y4 - 1, done);
    if (!c.avail) c = mvn_at(s, x4 - 1, y4 - 1, done);
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
static inline int mv_in_range(int x, int y) { return x >= -8192 && x <= 8191 && y >= -2048 && y <= 2047; }
/* the vector of a w4 x h4 partition goes to every 4x4 block it covers (raster order): whole 32-bit {hor, ver} pairs,
 * two at a time where the partition is at least 8 samples wide */
static inline void fill_mv(h264b200_mb_t *r, int x4, int y4, int w4, int h4, int mx, int my, unsigned *done)
{
    const uint32_t v = (uint32_t)(uint16_t)mx | ((uint32_t)(uint16_t)my << 16);
    const uint64_t vv = (uint64_t)v | ((uint64_t)v << 32);
    int j;
    for (j = y4; j < y4 + h4; j++) {
        int16_t *row = r->mv[j * 4 + x4];
        if (w4 == 4) { memcpy(row, &vv, 8); memcpy(row + 4, &vv, 8); }
        else if (w4 == 2) memcpy(row, &vv, 8);
        else memcpy(row, &v, 4);
        *done |= ((1u << w4) - 1u) << (j * 4 + x4);
    }
}

/* --------------------------------------------------------------- residual */
static inline int16_t *slot_ptr(sl_t *s, uint32_t slot) { return s->pic->coef + (size_t)slot * 16; }

static const uint8_t IDENT8[8] = {0, 1, 2, 3, 4, 5, 6, 7};

/* nC (9.2.1) from a TotalCoeff cache laid out with one guard row above and one guard column to the
 * left; 64 marks "not available", so that a+b >= 64 exactly when a neighbour is missing */
static inline int nc_of(const uint8_t *cache, int idx, int pitch)
{
    int n = cache[idx - 1] + cache[idx - pitch];
    if (n < 64) n = (n + 1) >> 1;
    return n & 31;
}

/* residual( ) of 7.3.5.3 for one macroblock; i16: Intra16x16 structure. Returns 0 / -1. */
static int parse_residual(sl_t *s, int cbp, int i16)
{
    h264b200_mb_t *r = s->rec; h264_mbctx_t *c = s->cur; br_t *b = s->b;
    uint32_t slot = s->pic->coef_used, mask = 0;
    int blk, pl, k, tc, dc_nz = 0;
    uint8_t lc[5 * 8], cc[2][3 * 4];
    /* guard cells from the left / upper macroblocks (luma4x4BlkIdx 5,7,13,15 = right column; 10,11,14,15 = bottom row) */
    memset(lc, 0, sizeof lc); memset(cc, 0, sizeof cc);
    if (s->cA) {
        const uint8_t *t = s->cA->tc;
        lc[8] = t[5]; lc[16] = t[7]; lc[24] = t[13]; lc[32] = t[15];
        cc[0][4] = t[17]; cc[0][8] = t[19]; cc[1][4] = t[21]; cc[1][8] = t[23];
    } else { lc[8] = lc[16] = lc[24] = lc[32] = 64; cc[0][4] = cc[0][8] = cc[1][4] = cc[1][8] = 64; }
    if (s->cB) {
        const uint8_t *t = s->cB->tc;
        lc[1] = t[10]; lc[2] = t[11]; lc[3] = t[14]; lc[4] = t[15];
        cc[0][1] = t[18]; cc[0][2] = t[19]; cc[1][1] = t[22]; cc[1][2] = t[23];
    } else { lc[1] = lc[2] = lc[3] = lc[4] = 64; cc[0][1] = cc[0][2] = cc[1][1] = cc[1][2] = 64; }

    r->coef_offset = slot;
    if (i16) {
        int16_t *p = slot_ptr(s, slot);
        tc = h264_cavlc_block(b, nc_of(lc, 9, 8), 16, p, H264_ZIGZAG4x4);
        if (tc < 0) return -1;
        if (tc) { dc_nz = 1; mask |= H264B200_RESID_LUMA_DC; slot++; }
    }
    for (blk = 0; blk < 16; blk++) {
        /* luma4x4BlkIdx -> cache index of (x4, y4) */
        const int idx = 9 + (((blk & 1) | ((blk >> 1) & 2))) + 8 * (((blk >> 1) & 1) | ((blk >> 2) & 2));
        int16_t *p;
        if (!((cbp >> (blk >> 2)) & 1)) {              /* 8x8 quadrant without AC: TotalCoeff stays 0 (context was cleared) */
            if (dc_nz) { int k4; for (k4 = 0; k4 < 4; k4++) { memset(slot_ptr(s, slot), 0, 32); mask |= 1u << (blk + k4); slot++; } }
            blk += 3;
            continue;
        }
        p = slot_ptr(s, slot);
        if (i16) tc = h264_cavlc_block(b, nc_of(lc, idx, 8), 15, p, H264_ZIGZAG4x4 + 1);
        else     tc = h264_cavlc_block(b, nc_of(lc, idx, 8), 16, p, H264_ZIGZAG4x4);
        if (tc < 0) return -1;
        c->tc[blk] = lc[idx] = (uint8_t)tc;
        if (tc) { r->nz_mask |= (uint16_t)(1u << blk); mask |= 1u << blk; slot++; }
        else if (dc_nz) { memset(p, 0, 32); mask |= 1u << blk; slot++; }
    }
    if (cbp & 0x30) {
        int16_t *p = slot_ptr(s, slot);
        int cdc[2];
        memset(p, 0, 32);                           /* both planes share this slot; a plane without coefficients leaves its half zero */
        for (pl = 0; pl < 2; pl++) {
            int16_t tmp[16];
            cdc[pl] = h264_cavlc_block(b, -1, 4, tmp, IDENT8);
            if (cdc[pl] < 0) return -1;
            if (cdc[pl]) memcpy(p + 4 * pl, tmp, 8);
        }
        if (cdc[0] || cdc[1]) { mask |= H264B200_RESID_CHROMA_DC; slot++; }
        for (pl = 0; pl < 2; pl++) for (k = 0; k < 4; k++) {
            const int idx = 5 + (k & 1) + 4 * (k >> 1);                 /* (y+1)*4 + (x+1) */
            p = slot_ptr(s, slot);
            tc = 0;
            if (cbp & 0x20) {
                tc = h264_cavlc_block(b, nc_of(cc[pl], idx, 4), 15, p, H264_ZIGZAG4x4 + 1);
                if (tc < 0) return -1;
                cc[pl][idx] = (uint8_t)tc;
            }
            c->tc[16 + 4 * pl + k] = (uint8_t)tc;
            if (tc) { mask |= 1u << (16 + 4 * pl + k); slot++; }
            else if (cdc[pl]) { memset(p, 0, 32); mask |= 1u << (16 + 4 * pl + k); slot++; }
        }
    }
    r->resid_mask = mask;
    s->pic->coef_used = slot;
    return 0;
}

/* ------------------------------------------------------------------ intra */
static inline int intra_usable(const sl_t *s, const h264_mbctx_t *c)
{
    return c && !(s->constrained_intra && c->kind == H264B200_MB_INTER);
}
static int pred_i4_mode(const sl_t *s, int blk)
{
    int r = H264_BLK_TO_RASTER[blk], x4 = r & 3, y4 = r >> 2, ma, mb;
    if (x4 > 0) ma = s->rec->i4_mode[H264_RASTER_TO_BLK[r - 1]];
    else {
        if (!intra_usable(s, s->cA)) return 2;
        ma = s->cA->kind == H264B200_MB_I4x4 ? s->rA->i4_mode[H264_RASTER_TO_BLK[r + 3]] : 2;
    }
    if (y4 > 0) mb = s->rec->i4_mode[H264_RASTER_TO_BLK[r - 4]];
    else {
        if (!intra_usable(s, s->cB)) return 2;
        mb = s->cB->kind == H264B200_MB_I4x4 ? s->rB->i4_mode[H264_RASTER_TO_BLK[12 + x4]] : 2;
    }
    return ma < mb ? ma : mb;
}

static inline int update_qp(sl_t *s, int delta)
{
    if (delta < -26 || delta > 25) return -1;
    if (delta) { s->qp += delta; if (s->qp < 0) s->qp += 52; else if (s->qp >= 52) s->qp -= 52; }
    return 0;
}
static inline void set_qp_fields(sl_t *s, h264b200_mb_t *r)
{
    int qc = s->qp + s->chroma_qp_off;
    qc = qc < 0 ? 0 : qc > 51 ? 51 : qc;
    r->qp_y = (uint8_t)s->qp; r->qp_dbk = (uint8_t)s->qp; r->qp_c = H264_QPC[qc];
}

static int parse_intra_mb(sl_t *s, uint32_t mb_type /* 0 I4x4, 1..24 I16x16, 25 I_PCM */)
{
    h264b200_mb_t *r = s->rec; h264_mbctx_t *c = s->cur; br_t *b = s->b;
    int aA = intra_usable(s, s->cA), aB = intra_usable(s, s->cB), aC = intra_usable(s, s->cC), aD = intra_usable(s, s->cD);
    uint32_t v; int blk, cbp;
    r->avail = (uint8_t)((aA ? H264B200_AVAIL_A : 0) | (aB ? H264B200_AVAIL_B : 0) | (aC ? H264B200_AVAIL_C : 0) | (aD ? H264B200_AVAIL_D : 0));
    c->ref_idx[0] = c->ref_idx[1] = c->ref_idx[2] = c->ref_idx[3] = -1;
    s->pic->n_intra++;
    if (mb_type == 25) {
        size_t byte_pos;
        uint8_t *dst;
        c->kind = r->mb_class = H264B200_MB_IPCM;
        /* pcm_alignment_zero_bit */
        while (br_pos(b) & 7) if (br_get1(b)) return -1;
        byte_pos = (size_t)(br_pos(b) >> 3);
        if (byte_pos + 384 > b->len) return -1;
        r->coef_offset = s->pic->coef_used;
        dst = (uint8_t *)slot_ptr(s, s->pic->coef_used);
        memcpy(dst, b->data + byte_pos, 384);
        s->pic->coef_used += 12;
        br_seek_bytes(b, byte_pos + 384);
        memset(c->tc, 16, sizeof c->tc);
        r->nz_mask = 0xffff;
        set_qp_fields(s, r);
        r->qp_dbk = 0;                          /* h264bsd_macroblock_layer.c:1003 */
        return 0;
    }
    if (mb_type == 0) {
        c->kind = r->mb_class = H264B200_MB_I4x4;
        for (blk = 0; blk < 16; blk++) {
            int pred = pred_i4_mode(s, blk), mode;
            if (br_get1(b)) mode = pred;
            else { int rem = (int)br_get(b, 3); mode = rem < pred ? rem : rem + 1; }
            r->i4_mode[blk] = (uint8_t)mode;
            /* legality given neighbour availability (h264bsd_intra_prediction.c:773-823) */
            {
                int rr = H264_BLK_TO_RASTER[blk], x4 = rr & 3, y4 = rr >> 2;
                int left = x4 > 0 ? 1 : aA, up = y4 > 0 ? 1 : aB;
                int ul = (x4 > 0 && y4 > 0) ? 1 : x4 > 0 ? aB : y4 > 0 ? aA : aD;
                switch (mode) {
                case 0: case 3: case 7: if (!up) return -1; break;
                case 1: case 8: if (!left) return -1; break;
                case 4: case 5: case 6: if (!up || !left || !ul) return -1; break;
                default: break;
                }
            }
        }
    } else {
        c->kind = r->mb_class = H264B200_MB_I16x16;
        r->i16_mode = (uint8_t)((mb_type - 1) & 3);
        switch (r->i16_mode) {
        case 0: if (!aB) return -1; break;
        case 1: if (!aA) return -1; break;
        case 3: if (!aA || !aB || !aD) return -1; break;
        default: break;
        }
    }
    v = br_ue(b); if (v > 3) return -1;
    r->chroma_mode = (uint8_t)v;
    switch (v) {
    case 1: if (!aA) return -1; break;
    case 2: if (!aB) return -1; break;
    case 3: if (!aA || !aB || !aD) return -1; break;
    default: break;
    }
    if (mb_type == 0) {
        v = br_ue(b); if (v > 47) return -1;
        cbp = H264_CBP_MAP[v][0];
    } else cbp = (((mb_type - 1) >> 2) % 3) << 4 | (mb_type >= 13 ? 15 : 0);
    if (cbp || mb_type != 0) {
        if (update_qp(s, br_se(b))) return -1;
        set_qp_fields(s, r);
        if (parse_residual(s, cbp, mb_type != 0)) return -1;
    } else set_qp_fields(s, r);
    return 0;
}

/* ------------------------------------------------------------------ inter */
static inline int read_ref_idx(sl_t *s, uint32_t n_active)
{
    uint32_t v;
    if (n_active <= 1) return 0;
    v = br_te(s->b, n_active - 1);
    if (v >= n_active) return -1;
    return (int)v;
}
static inline int set_ref(sl_t *s, int q, int ref)
{
    int slot = s->ref_slot[ref];
    if (slot < 0) return -1;                    /* missing / non-existing reference (h264bsd_dpb.c:846-860) */
    s->cur->ref_idx[q] = (int8_t)ref; s->rec->ref_slot[q] = (uint8_t)slot;
    s->pic->ref_slots_used[slot] = 1;
    return 0;
}

static int parse_inter_mb(sl_t *s, uint32_t mb_type /* 0..4 */)
{
    h264b200_mb_t *r = s->rec; h264_mbctx_t *c = s->cur; br_t *b = s->b;
    uint32_t n_active = s->sh->num_ref_idx_active, v;
    unsigned done = 0;
    int px, py, mx, my, i, cbp;
    c->kind = r->mb_class = H264B200_MB_INTER;
    s->pic->n_inter++;
    if (mb_type == 0) {
        int ref = read_ref_idx(s, n_active), dx, dy;
        if (ref < 0) return -1;
        for (i = 0; i < 4; i++) if (set_ref(s, i, ref)) return -1;
        dx = br_se(b); dy = br_se(b);
        predict_mv(s, 0, 0, 4, ref, 0, 0, &px, &py);
        mx = (int16_t)((unsigned)px + (unsigned)dx); my = (int16_t)((unsigned)py + (unsigned)dy);
        if (!mv_in_range(mx, my)) return -1;
        fill_mv(r, 0, 0, 4, 4, mx, my, &done);
        r->part_flags = 31;
    } else if (mb_type == 1 || mb_type == 2) {
        int ref[2], dx[2], dy[2];
        for (i = 0; i < 2; i++) { ref[i] = read_ref_idx(s, n_active); if (ref[i] < 0) return -1; }
        for (i = 0; i < 2; i++) { dx[i] = br_se(b); dy[i] = br_se(b); }
        if (mb_type == 1) { if (set_ref(s, 0, ref[0]) || set_ref(s, 1, ref[0]) || set_ref(s, 2, ref[1]) || set_ref(s, 3, ref[1])) return -1; }
        else              { if (set_ref(s, 0, ref[0]) || set_ref(s, 2, ref[0]) || set_ref(s, 1, ref[1]) || set_ref(s, 3, ref[1])) return -1; }
        for (i = 0; i < 2; i++) {
            if (mb_type == 1) predict_mv(s, 0, 2 * i, 4, ref[i], done, i == 0 ? 2 : 1, &px, &py);
            else              predict_mv(s, 2 * i, 0, 2, ref[i], done, i == 0 ? 1 : 3, &px, &py);
            mx = (int16_t)((unsigned)px + (unsigned)dx[i]); my = (int16_t)((unsigned)py + (unsigned)dy[i]);
            if (!mv_in_range(mx, my)) return -1;
            if (mb_type == 1) fill_mv(r, 0, 2 * i, 4, 2, mx, my, &done);
            else              fill_mv(r, 2 * i, 0, 2, 4, mx, my, &done);
        }
        r->part_flags = 15;
    } else {
        int sub[4], ref[4], q, k, mvd[16][2], n = 0, m = 0;
        for (q = 0; q < 4; q++) { v = br_ue(b); if (v > 3) return -1; sub[q] = (int)v; if (!v) r->part_flags |= (uint8_t)(1 << q); }
        for (q = 0; q < 4; q++) {
            ref[q] = mb_type == 4 ? 0 : read_ref_idx(s, n_active);
            if (ref[q] < 0 || set_ref(s, q, ref[q])) return -1;
        }
        for (q = 0; q < 4; q++) { int cnt = sub[q] == 0 ? 1 : sub[q] == 3 ? 4 : 2; for (k = 0; k < cnt; k++) { mvd[n][0] = br_se(b); mvd[n][1] = br_se(b); n++; } }
        for (q = 0; q < 4; q++) {
            int ox = (q & 1) * 2, oy = (q >> 1) * 2, cnt = sub[q] == 0 ? 1 : sub[q] == 3 ? 4 : 2;
            for (k = 0; k < cnt; k++, m++) {
                int x4, y4, w4, h4;
                switch (sub[q]) {
                case 0: x4 = ox; y4 = oy; w4 = 2; h4 = 2; break;
                case 1: x4 = ox; y4 = oy + k; w4 = 2; h4 = 1; break;
                case 2: x4 = ox + k; y4 = oy; w4 = 1; h4 = 2; break;
                default: x4 = ox + (k & 1); y4 = oy + (k >> 1); w4 = 1; h4 = 1; break;
                }
                predict_mv(s, x4, y4, w4, ref[q], done, 0, &px, &py);
                mx = (int16_t)((unsigned)px + (unsigned)mvd[m][0]); my = (int16_t)((unsigned)py + (unsigned)mvd[m][1]);
                if (!mv_in_range(mx, my)) return -1;
                fill_mv(r, x4, y4, w4, h4, mx, my, &done);
            }
        }
    }
    v = br_ue(b); if (v > 47) return -1;
    cbp = H264_CBP_MAP[v][1];
    if (cbp) {
        if (update_qp(s, br_se(b))) return -1;
        set_qp_fields(s, r);
        if (parse_residual(s, cbp, 0)) return -1;
    } else set_qp_fields(s, r);
    return 0;
}

static int do_skip_mb(sl_t *s)
{
    h264b200_mb_t *r = s->rec; h264_mbctx_t *c = s->cur;
    mvn_t a = mvn_at(s, -1, 0, 0), bq = mvn_at(s, 0, -1, 0);
    int mx = 0, my = 0, i;
    unsigned done = 0;
    c->kind = r->mb_class = H264B200_MB_INTER;
    s->pic->n_inter++;
    for (i = 0; i < 4; i++) if (set_ref(s, i, 0)) return -1;
    if (a.avail && bq.avail && !(a.ref == 0 && a.x == 0 && a.y == 0) && !(bq.ref == 0 && bq.x == 0 && bq.y == 0)) {
        predict_mv(s, 0, 0, 4, 0, 0, 0, &mx, &my);
        if (!mv_in_range(mx, my)) return -1;
    }
    fill_mv(r, 0, 0, 4, 4, mx, my, &done);
    r->part_flags = 31;
    set_qp_fields(s, r);
    return 0;
}

/* --------------------------------------------------------------- the loop */
int h264_decode_slice_data(h264_decoder_t *d, br_t *b, const h264_slice_hdr_t *sh)
{
    sl_t s;
    uint32_t skip_run = 0, mb_count = 0, addr = sh->first_mb, next_linear = 0xffffffffu, i;
    int prev_skipped = 0, more;
    memset(&s, 0, sizeof s);
    s.d = d; s.b = b; s.sh = sh; s.pic = d->pic; s.W = d->width_mbs; s.H = d->height_mbs;
    s.is_p = sh->slice_type == 0;
    s.constrained_intra = d->active_pps->constrained_intra_pred;
    s.chroma_qp_off = d->active_pps->chroma_qp_index_offset;
    s.qp = sh->slice_qp;
    s.slice_id = (uint16_t)(++d->slice_id);
    d->slice_last_mb = 0;
    for (i = 0; i <= H264_MAX_REFS; i++) s.ref_slot[i] = -1;
    if (s.is_p) for (i = 0; i < sh->num_ref_idx_active && i <= H264_MAX_REFS; i++) s.ref_slot[i] = h264_dpb_ref_slot(&d->dpb, i);

    do {
        h264_mbctx_t *c = &d->mbctx[addr];
        h264b200_mb_t *r;
        int rc = 0;
        if (c->decoded) return -1;              /* redundant pictures are not decoded; a primary MB twice is an error */
        /* growing moves the picture's records AND slots (one block): take pointers only afterwards */
        if (s.pic->coef_used + 32 > s.pic->coef_cap && d->be->coef_grow(d->be, d->be_inst, s.pic, s.pic->coef_used + 4096)) return -1;
        r = &s.pic->mbs[addr];
        s.addr = addr;
        if (addr != next_linear) { s.mbx = (int)(addr % s.W); s.mby = (int)(addr / s.W); }    /* first macroblock, or a jump in the slice group map */
        else if (++s.mbx == (int)s.W) { s.mbx = 0; s.mby++; }
        next_linear = addr + 1;
        s.cur = c; s.rec = r;
        memset(r, 0, 64);                       /* mv[] is always written for inter MBs and never read for intra */
        memset(c, 0, sizeof *c);
        c->slice_id = s.slice_id; r->slice_id = s.slice_id;
        r->chroma_qp_off = (int8_t)s.chroma_qp_off;
        r->dbk_off_a = sh->alpha_off; r->dbk_off_b = sh->beta_off;
        r->dbk_idc = sh->disable_deblocking_idc;
        set_neighbours(&s);
        if (s.is_p && !prev_skipped) {
            skip_run = br_ue(b);
            if (skip_run == 0xffffffffu || skip_run > d->pic_size_mbs - addr) return -1;
            if (skip_run) prev_skipped = 1;
        }
        if (skip_run) { skip_run--; rc = do_skip_mb(&s); }
        else {
            uint32_t mb_type = br_ue(b);
            prev_skipped = 0;
            if (s.is_p) {
                if (mb_type > 30) return -1;
                rc = mb_type < 5 ? parse_inter_mb(&s, mb_type) : parse_intra_mb(&s, mb_type - 5);
            } else {
                if (mb_type > 25) return -1;
                rc = parse_intra_mb(&s, mb_type);
            }
        }
        if (rc || br_overrun(b)) { c->slice_id = 0; return -1; }
        /* deblocking edge flags (h264bsd_deblocking.c:288-319); idc 2 compares slice ids */
        if (sh->disable_deblocking_idc != 1) {
            int fl = H264B200_DBK_INNER;
            if (s.mbx > 0 && (sh->disable_deblocking_idc != 2 || d->mbctx[addr - 1].slice_id == s.slice_id)) fl |= H264B200_DBK_LEFT;
            if (s.mby > 0 && (sh->disable_deblocking_idc != 2 || d->mbctx[addr - s.W].slice_id == s.slice_id)) fl |= H264B200_DBK_TOP;
            r->dbk_flags = (uint8_t)fl;
            s.pic->any_deblock = 1;
        }
        c->decoded = 1;
        mb_count++;
        if (!s.is_p) d->slice_last_mb = addr;       /* h264bsd_slice_data.c:208-211 */
        more = br_more_data(b) || skip_run;
        if (d->active_pps->num_slice_groups > 1) {   /* next macroblock of the same slice group (h264bsd_util.c:219-245) */
            const uint8_t *map = d->slice_group_map, grp = map[addr];
            do addr++; while (addr < d->pic_size_mbs && map[addr] != grp);
        } else addr++;
        if (more && addr >= d->pic_size_mbs) return -1;
    } while (more);
    if (d->num_decoded_mbs + mb_count > d->pic_size_mbs) return -1;
    d->num_decoded_mbs += mb_count;
    return 0;
}
