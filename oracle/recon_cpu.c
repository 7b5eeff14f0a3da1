/* oracle/recon_cpu.c — TEST INFRASTRUCTURE ONLY.  Not linked into libh264b200.so.
 *
 * A plain-C, single-threaded restatement of the reference's pixel
 * reconstruction, driven by the macroblock RECORDS the product's host parser
 * emits (include/h264b200_records.h) instead of by the bitstream.  It exists to
 *   (1) prove, on a CPU-only box, that the host parser + DPB + record format are
 *       sufficient and correct: records -> pixels here must be MD5-identical to
 *       the unmodified reference (oracle/_ref) on every test stream, and
 *   (2) give the GPU parity tests a stage-by-stage checker (residual slots after
 *       K1, the picture before deblocking, the picture after deblocking).
 * PARITY PINNING: the reference ships no golden vectors; this
 * restatement is pinned by tests/test_oracle_cpu.py: against the per-frame MD5
 * fixtures under tests/golden/ that tools/make_golden.py generated with oracle/_ref
 * (the unmodified reference), against the live reference on a seeded parameter
 * sweep where oracle/_ref is built, and by tests/test_ffmpeg_crosscheck_cpu.py
 * against FFmpeg's luma.
 *
 * What each part restates (reference file:line):
 *   residual_mb      h264bsd_macroblock_layer.c:1343-1424 ProcessResidual,
 *                    h264bsd_transform.c:94-231 / :252-335 / :356-398
 *   inter_mb         h264bsd_inter_prediction.c:364-487, h264bsd_reconstruct.c
 *                    :1819-1941 h264bsdPredictSamples (+ 9 luma interpolators
 *                    :491-1791, chroma :110-476, edge clamp :2222-2314),
 *                    h264bsd_image.c:171-343 h264bsdWriteOutputBlocks
 *   intra_mb         h264bsd_intra_prediction.c:475-532, :626-686 (+ :999-1148),
 *                    :700-832 (+ :1387-1829), :844-914 (+ :1159-1376), :926-988
 *   deblock_picture  h264bsd_deblocking.c:574-1736 (non-OMXDL half)
 * It plugs in as an `h264_backend_t` (csrc/h264_internal.h), so the very same
 * host decoder sources run above it: oracle/libh264b200_cpuchk.so.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "recon_cpu.h"
#include "../broadway_b200/csrc/h264_consts.h"

static inline int clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static inline int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }
static inline int iabs(int v) { return v < 0 ? -v : v; }

/* ======================================================== K1: residual */
static void idct4x4(const int *d, int *r)
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

/* returns 1 if any residual sample left [-512,511] (h264bsd_transform.c:181-185) */
static int block_residual(int16_t *slot, int qp, int have_dc, int dc)
{
    int d[16], r[16], i, bad = 0;
    for (i = 0; i < 16; i++) d[i] = (slot[i] * H264_LEVEL_SCALE[qp % 6][H264_POS_CLASS[i]]) << (qp / 6);
    if (have_dc) d[0] = dc;
    idct4x4(d, r);
    for (i = 0; i < 16; i++) { if (r[i] < -512 || r[i] > 511) bad = 1; slot[i] = (int16_t)r[i]; }
    return bad;
}

int recon_cpu_residual_mb(const h264b200_mb_t *mb, int16_t *coef)
{
    int16_t *slot = coef + (size_t)mb->coef_offset * 16;
    int dcy[16], dcc[8], blk, k, bad = 0, qp = mb->qp_y, qpc = mb->qp_c;
    int has_ldc = (mb->resid_mask & H264B200_RESID_LUMA_DC) != 0, has_cdc = (mb->resid_mask & H264B200_RESID_CHROMA_DC) != 0;
    if (mb->mb_class == H264B200_MB_IPCM || mb->mb_class == H264B200_MB_MISSING || mb->mb_class == H264B200_MB_CONCEAL) return 0;
    if (has_ldc) {
        int f[16], i, ls = H264_LEVEL_SCALE[qp % 6][0];
        for (i = 0; i < 4; i++) {
            const int16_t *s = slot + 4 * i;
            f[4*i] = s[0] + s[1] + s[2] + s[3]; f[4*i+1] = s[0] + s[1] - s[2] - s[3];
            f[4*i+2] = s[0] - s[1] - s[2] + s[3]; f[4*i+3] = s[0] - s[1] + s[2] - s[3];
        }
        for (i = 0; i < 4; i++) {
            int g[4];
            g[0] = f[i] + f[4+i] + f[8+i] + f[12+i]; g[1] = f[i] + f[4+i] - f[8+i] - f[12+i];
            g[2] = f[i] - f[4+i] - f[8+i] + f[12+i]; g[3] = f[i] - f[4+i] + f[8+i] - f[12+i];
            for (k = 0; k < 4; k++) {
                int v = g[k] * ls;
                dcy[4*k + i] = qp >= 12 ? v << (qp / 6 - 2) : (v + (1 << (1 - qp / 6))) >> (2 - qp / 6);
            }
        }
        slot += 16;
    }
    for (blk = 0; blk < 16; blk++) if ((mb->resid_mask >> blk) & 1) {
        bad |= block_residual(slot, qp, has_ldc, has_ldc ? dcy[H264_BLK_TO_RASTER[blk]] : 0);
        slot += 16;
    }
    if (has_cdc) {
        int ls = H264_LEVEL_SCALE[qpc % 6][0], pl;
        for (pl = 0; pl < 2; pl++) {
            const int16_t *c = slot + 4 * pl;
            int f[4];
            f[0] = c[0] + c[1] + c[2] + c[3]; f[1] = c[0] - c[1] + c[2] - c[3];
            f[2] = c[0] + c[1] - c[2] - c[3]; f[3] = c[0] - c[1] - c[2] + c[3];
            for (k = 0; k < 4; k++) { int v = f[k] * ls; dcc[4*pl + k] = qpc >= 6 ? v << (qpc / 6 - 1) : v >> 1; }
        }
        slot += 16;
    }
    for (blk = 16; blk < 24; blk++) if ((mb->resid_mask >> blk) & 1) {
        bad |= block_residual(slot, qpc, has_cdc, has_cdc ? dcc[blk - 16] : 0);
        slot += 16;
    }
    return bad;
}

/* slot of block blk (0..23) after K1, or NULL when the block has no residual */
static const int16_t *resid_slot(const h264b200_mb_t *mb, const int16_t *coef, int blk)
{
    uint32_t m = mb->resid_mask, below;
    if (!((m >> blk) & 1)) return NULL;
    below = m & ((1u << blk) - 1);
    if (m & H264B200_RESID_LUMA_DC) below |= 1u << 30;
    if (blk >= 16 && (m & H264B200_RESID_CHROMA_DC)) below |= 1u << 31;
    return coef + ((size_t)mb->coef_offset + (size_t)__builtin_popcount(below)) * 16;
}

/* ==================================================== frame addressing */
typedef struct { uint8_t *y, *cb, *cr; int w, h; } planes_t;     /* w,h in luma pels */
static planes_t planes_of(uint8_t *frame, int wm, int hm)
{
    planes_t p; p.w = wm * 16; p.h = hm * 16; p.y = frame; p.cb = frame + (size_t)p.w * p.h; p.cr = p.cb + (size_t)(p.w / 2) * (p.h / 2);
    return p;
}
static inline int refpel(const uint8_t *pl, int w, int h, int x, int y)
{
    x = x < 0 ? 0 : x >= w ? w - 1 : x; y = y < 0 ? 0 : y >= h ? h - 1 : y;
    return pl[(size_t)y * w + x];
}

/* ==================================================== K2: inter prediction */
static inline int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }
static int hb1(const uint8_t *p, int w, int h, int x, int y)
{ return tap6(refpel(p,w,h,x-2,y), refpel(p,w,h,x-1,y), refpel(p,w,h,x,y), refpel(p,w,h,x+1,y), refpel(p,w,h,x+2,y), refpel(p,w,h,x+3,y)); }
static int vh1(const uint8_t *p, int w, int h, int x, int y)
{ return tap6(refpel(p,w,h,x,y-2), refpel(p,w,h,x,y-1), refpel(p,w,h,x,y), refpel(p,w,h,x,y+1), refpel(p,w,h,x,y+2), refpel(p,w,h,x,y+3)); }
static int half_b(const uint8_t *p, int w, int h, int x, int y) { return clip255((hb1(p,w,h,x,y) + 16) >> 5); }
static int half_h(const uint8_t *p, int w, int h, int x, int y) { return clip255((vh1(p,w,h,x,y) + 16) >> 5); }
static int half_j(const uint8_t *p, int w, int h, int x, int y)
{
    int j1 = tap6(hb1(p,w,h,x,y-2), hb1(p,w,h,x,y-1), hb1(p,w,h,x,y), hb1(p,w,h,x,y+1), hb1(p,w,h,x,y+2), hb1(p,w,h,x,y+3));
    return clip255((j1 + 512) >> 10);
}
/* luma sample at integer position (x,y) + fraction (fx,fy) quarter pels: 8.4.2.2.1 */
int recon_cpu_luma_sample(const uint8_t *p, int w, int h, int x, int y, int fx, int fy)
{
    int G = refpel(p,w,h,x,y);
    switch (fy * 4 + fx) {
    case 0:  return G;
    case 1:  return (G + half_b(p,w,h,x,y) + 1) >> 1;                               /* a */
    case 2:  return half_b(p,w,h,x,y);                                              /* b */
    case 3:  return (refpel(p,w,h,x+1,y) + half_b(p,w,h,x,y) + 1) >> 1;             /* c */
    case 4:  return (G + half_h(p,w,h,x,y) + 1) >> 1;                               /* d */
    case 5:  return (half_b(p,w,h,x,y) + half_h(p,w,h,x,y) + 1) >> 1;               /* e */
    case 6:  return (half_b(p,w,h,x,y) + half_j(p,w,h,x,y) + 1) >> 1;               /* f */
    case 7:  return (half_b(p,w,h,x,y) + half_h(p,w,h,x+1,y) + 1) >> 1;             /* g */
    case 8:  return half_h(p,w,h,x,y);                                              /* h */
    case 9:  return (half_h(p,w,h,x,y) + half_j(p,w,h,x,y) + 1) >> 1;               /* i */
    case 10: return half_j(p,w,h,x,y);                                              /* j */
    case 11: return (half_j(p,w,h,x,y) + half_h(p,w,h,x+1,y) + 1) >> 1;             /* k */
    case 12: return (refpel(p,w,h,x,y+1) + half_h(p,w,h,x,y) + 1) >> 1;             /* n */
    case 13: return (half_h(p,w,h,x,y) + half_b(p,w,h,x,y+1) + 1) >> 1;             /* p */
    case 14: return (half_j(p,w,h,x,y) + half_b(p,w,h,x,y+1) + 1) >> 1;             /* q */
    default: return (half_h(p,w,h,x+1,y) + half_b(p,w,h,x,y+1) + 1) >> 1;           /* r */
    }
}
int recon_cpu_chroma_sample(const uint8_t *p, int w, int h, int x, int y, int fx, int fy)
{
    int A = refpel(p,w,h,x,y), B = refpel(p,w,h,x+1,y), C = refpel(p,w,h,x,y+1), D = refpel(p,w,h,x+1,y+1);
    return ((8 - fx) * (8 - fy) * A + fx * (8 - fy) * B + (8 - fx) * fy * C + fx * fy * D + 32) >> 6;
}

static void inter_mb(const h264b200_mb_t *mb, const int16_t *coef, int mbx, int mby, planes_t cur, uint8_t *const *frames, int wm, int hm)
{
    int b4, x, y, pl;
    for (b4 = 0; b4 < 16; b4++) {                   /* raster 4x4 blocks, each with its own vector */
        int bx = b4 & 3, by = b4 >> 2, mvx = mb->mv[b4][0], mvy = mb->mv[b4][1];
        planes_t ref = planes_of(frames[mb->ref_slot[(by >> 1) * 2 + (bx >> 1)]], wm, hm);
        const int16_t *rs = resid_slot(mb, coef, H264_RASTER_TO_BLK[b4]);
        int x0 = mbx * 16 + bx * 4, y0 = mby * 16 + by * 4;
        for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) {
            int v = recon_cpu_luma_sample(ref.y, ref.w, ref.h, x0 + x + (mvx >> 2), y0 + y + (mvy >> 2), mvx & 3, mvy & 3);
            if (rs) v = clip255(v + rs[y * 4 + x]);
            cur.y[(size_t)(y0 + y) * cur.w + x0 + x] = (uint8_t)v;
        }
        for (pl = 0; pl < 2; pl++) {                /* the 2x2 chroma pels under this luma block */
            const uint8_t *rp = pl ? ref.cr : ref.cb; uint8_t *cp = pl ? cur.cr : cur.cb;
            int cx0 = mbx * 8 + bx * 2, cy0 = mby * 8 + by * 2, cblk = 16 + 4 * pl + (by >> 1) * 2 + (bx >> 1);
            const int16_t *cs = resid_slot(mb, coef, cblk);
            for (y = 0; y < 2; y++) for (x = 0; x < 2; x++) {
                int v = recon_cpu_chroma_sample(rp, ref.w / 2, ref.h / 2, cx0 + x + (mvx >> 3), cy0 + y + (mvy >> 3), mvx & 7, mvy & 7);
                if (cs) v = clip255(v + cs[((by & 1) * 2 + y) * 4 + (bx & 1) * 2 + x]);
                cp[(size_t)(cy0 + y) * (cur.w / 2) + cx0 + x] = (uint8_t)v;
            }
        }
    }
}

/* ==================================================== K3: intra prediction */
static void add_block(uint8_t *dst, int stride, const uint8_t *pred /* 4x4 */, const int16_t *rs)
{
    int x, y;
    for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) dst[y * stride + x] = (uint8_t)(rs ? clip255(pred[y * 4 + x] + rs[y * 4 + x]) : pred[y * 4 + x]);
}

static void intra4x4_pred(uint8_t *pred, int mode, const int *top /* [-1..7] via top[i+1] */, const int *left /* [0..3] */, int has_top, int has_left)
{
#define T(i) top[(i) + 1]
#define L(i) ((i) < 0 ? top[0] : left[i])
    int x, y;
    for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) {
        int v;
        switch (mode) {
        case 0: v = T(x); break;
        case 1: v = L(y); break;
        case 2:
            if (has_top && has_left) v = (T(0) + T(1) + T(2) + T(3) + L(0) + L(1) + L(2) + L(3) + 4) >> 3;
            else if (has_left) v = (L(0) + L(1) + L(2) + L(3) + 2) >> 2;
            else if (has_top) v = (T(0) + T(1) + T(2) + T(3) + 2) >> 2;
            else v = 128;
            break;
        case 3: v = (x == 3 && y == 3) ? (T(6) + 3 * T(7) + 2) >> 2 : (T(x + y) + 2 * T(x + y + 1) + T(x + y + 2) + 2) >> 2; break;
        case 4:
            if (x > y) v = (T(x - y - 2) + 2 * T(x - y - 1) + T(x - y) + 2) >> 2;
            else if (x < y) v = (L(y - x - 2) + 2 * L(y - x - 1) + L(y - x) + 2) >> 2;
            else v = (T(0) + 2 * T(-1) + L(0) + 2) >> 2;
            break;
        case 5: {
            int z = 2 * x - y;
            if (z >= 0 && !(z & 1)) v = (T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 1) >> 1;
            else if (z >= 0) v = (T(x - (y >> 1) - 2) + 2 * T(x - (y >> 1) - 1) + T(x - (y >> 1)) + 2) >> 2;
            else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
            else v = (L(y - 1) + 2 * L(y - 2) + L(y - 3) + 2) >> 2;
            break; }
        case 6: {
            int z = 2 * y - x;
            if (z >= 0 && !(z & 1)) v = (L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 1) >> 1;
            else if (z >= 0) v = (L(y - (x >> 1) - 2) + 2 * L(y - (x >> 1) - 1) + L(y - (x >> 1)) + 2) >> 2;
            else if (z == -1) v = (L(0) + 2 * T(-1) + T(0) + 2) >> 2;
            else v = (T(x - 1) + 2 * T(x - 2) + T(x - 3) + 2) >> 2;
            break; }
        case 7:
            if (!(y & 1)) v = (T(x + (y >> 1)) + T(x + (y >> 1) + 1) + 1) >> 1;
            else v = (T(x + (y >> 1)) + 2 * T(x + (y >> 1) + 1) + T(x + (y >> 1) + 2) + 2) >> 2;
            break;
        default: {
            int z = x + 2 * y;
            if (z > 5) v = L(3);
            else if (z == 5) v = (L(2) + 3 * L(3) + 2) >> 2;
            else if (!(z & 1)) v = (L(y + (x >> 1)) + L(y + (x >> 1) + 1) + 1) >> 1;
            else v = (L(y + (x >> 1)) + 2 * L(y + (x >> 1) + 1) + L(y + (x >> 1) + 2) + 2) >> 2;
            break; }
        }
        pred[y * 4 + x] = (uint8_t)v;
    }
#undef T
#undef L
}

static void plane_pred(uint8_t *dst, int stride, int n /* 16 or 8 */, const int *top /* [-1..n-1] via +1 */, const int *left /* [-1..n-1] via +1 */)
{
    int half = n / 2, hh = 0, vv = 0, i, a, b, c, x, y;
    for (i = 0; i < half; i++) {
        hh += (i + 1) * (top[half + i + 1] - top[half - 2 - i + 1]);
        vv += (i + 1) * (left[half + i + 1] - left[half - 2 - i + 1]);
    }
    a = 16 * (left[n] + top[n]);
    if (n == 16) { b = (5 * hh + 32) >> 6; c = (5 * vv + 32) >> 6; }
    else { b = (34 * hh + 32) >> 6; c = (34 * vv + 32) >> 6; }
    for (y = 0; y < n; y++) for (x = 0; x < n; x++)
        dst[y * stride + x] = (uint8_t)clip255((a + b * (x - (half - 1)) + c * (y - (half - 1)) + 16) >> 5);
}

static void intra_mb(const h264b200_mb_t *mb, const int16_t *coef, int mbx, int mby, planes_t cur)
{
    int aA = mb->avail & H264B200_AVAIL_A, aB = mb->avail & H264B200_AVAIL_B, aC = mb->avail & H264B200_AVAIL_C, aD = mb->avail & H264B200_AVAIL_D;
    uint8_t *Y = cur.y + (size_t)mby * 16 * cur.w + mbx * 16;
    int x, y, blk, pl, cw = cur.w / 2;
    if (mb->mb_class == H264B200_MB_IPCM) {
        const uint8_t *s = (const uint8_t *)(coef + (size_t)mb->coef_offset * 16);
        for (y = 0; y < 16; y++) memcpy(Y + (size_t)y * cur.w, s + y * 16, 16);
        for (pl = 0; pl < 2; pl++) {
            uint8_t *C = (pl ? cur.cr : cur.cb) + (size_t)mby * 8 * cw + mbx * 8;
            for (y = 0; y < 8; y++) memcpy(C + (size_t)y * cw, s + 256 + 64 * pl + y * 8, 8);
        }
        return;
    }
    if (mb->mb_class == H264B200_MB_I16x16) {
        uint8_t pred[256];
        int top[17], left[17];
        for (x = -1; x < 16; x++) top[x + 1] = (aB && (x >= 0 || aD)) ? Y[-cur.w + x] : 0;
        for (y = -1; y < 16; y++) left[y + 1] = (y < 0) ? top[0] : (aA ? Y[(size_t)y * cur.w - 1] : 0);
        if (!aB && aD) top[0] = left[0] = Y[-cur.w - 1];
        switch (mb->i16_mode) {
        case 0: for (y = 0; y < 16; y++) for (x = 0; x < 16; x++) pred[y * 16 + x] = (uint8_t)top[x + 1]; break;
        case 1: for (y = 0; y < 16; y++) for (x = 0; x < 16; x++) pred[y * 16 + x] = (uint8_t)left[y + 1]; break;
        case 2: {
            int s = 0, v;
            if (aA && aB) { for (x = 0; x < 16; x++) s += top[x + 1] + left[x + 1]; v = (s + 16) >> 5; }
            else if (aA) { for (x = 0; x < 16; x++) s += left[x + 1]; v = (s + 8) >> 4; }
            else if (aB) { for (x = 0; x < 16; x++) s += top[x + 1]; v = (s + 8) >> 4; }
            else v = 128;
            memset(pred, v, 256);
            break; }
        default: plane_pred(pred, 16, 16, top, left); break;
        }
        for (blk = 0; blk < 16; blk++) {
            int bx = H264_BLK_X[blk], by = H264_BLK_Y[blk];
            uint8_t p4[16];
            for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) p4[y * 4 + x] = pred[(by + y) * 16 + bx + x];
            add_block(Y + (size_t)by * cur.w + bx, cur.w, p4, resid_slot(mb, coef, blk));
        }
    } else {
        for (blk = 0; blk < 16; blk++) {             /* decoding order: each block sees its reconstructed neighbours */
            int bx = H264_BLK_X[blk], by = H264_BLK_Y[blk], x4 = bx >> 2, y4 = by >> 2;
            uint8_t *B = Y + (size_t)by * cur.w + bx, p4[16];
            int has_left = x4 > 0 ? 1 : aA != 0, has_top = y4 > 0 ? 1 : aB != 0;
            int has_ul = (x4 > 0 && y4 > 0) ? 1 : x4 > 0 ? aB != 0 : y4 > 0 ? aA != 0 : aD != 0;
            int has_ur, top[9], left[4], i;
            if (y4 == 0) has_ur = x4 < 3 ? aB != 0 : aC != 0;
            else if (x4 == 3) has_ur = 0;
            else has_ur = H264_RASTER_TO_BLK[(y4 - 1) * 4 + x4 + 1] < blk;
            top[0] = has_ul ? B[-cur.w - 1] : 0;
            for (i = 0; i < 4; i++) top[i + 1] = has_top ? B[-cur.w + i] : 0;
            for (i = 4; i < 8; i++) top[i + 1] = has_ur ? B[-cur.w + i] : top[4];      /* replicate p[3,-1] (:789-792) */
            for (i = 0; i < 4; i++) left[i] = has_left ? B[(size_t)i * cur.w - 1] : 0;
            intra4x4_pred(p4, mb->i4_mode[blk], top, left, has_top, has_left);
            add_block(B, cur.w, p4, resid_slot(mb, coef, blk));
        }
    }
    for (pl = 0; pl < 2; pl++) {                     /* chroma 8x8 */
        uint8_t *C = (pl ? cur.cr : cur.cb) + (size_t)mby * 8 * cw + mbx * 8, pred[64];
        int top[9], left[9];
        for (x = -1; x < 8; x++) top[x + 1] = (aB && (x >= 0 || aD)) ? C[-cw + x] : 0;
        for (y = -1; y < 8; y++) left[y + 1] = (y < 0) ? top[0] : (aA ? C[(size_t)y * cw - 1] : 0);
        if (!aB && aD) top[0] = left[0] = C[-cw - 1];
        switch (mb->chroma_mode) {
        case 0:
            for (blk = 0; blk < 4; blk++) {
                int xo = (blk & 1) * 4, yo = (blk >> 1) * 4, st = 0, sl = 0, v, i;
                for (i = 0; i < 4; i++) { st += top[xo + i + 1]; sl += left[yo + i + 1]; }
                if (blk == 0 || blk == 3) {
                    if (aA && aB) v = (st + sl + 4) >> 3; else if (aB) v = (st + 2) >> 2; else if (aA) v = (sl + 2) >> 2; else v = 128;
                } else if (blk == 1) {
                    if (aB) v = (st + 2) >> 2; else if (aA) v = (sl + 2) >> 2; else v = 128;
                } else {
                    if (aA) v = (sl + 2) >> 2; else if (aB) v = (st + 2) >> 2; else v = 128;
                }
                for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) pred[(yo + y) * 8 + xo + x] = (uint8_t)v;
            }
            break;
        case 1: for (y = 0; y < 8; y++) for (x = 0; x < 8; x++) pred[y * 8 + x] = (uint8_t)left[y + 1]; break;
        case 2: for (y = 0; y < 8; y++) for (x = 0; x < 8; x++) pred[y * 8 + x] = (uint8_t)top[x + 1]; break;
        default: plane_pred(pred, 8, 8, top, left); break;
        }
        for (blk = 0; blk < 4; blk++) {
            int bx = (blk & 1) * 4, by = (blk >> 1) * 4;
            uint8_t p4[16];
            for (y = 0; y < 4; y++) for (x = 0; x < 4; x++) p4[y * 4 + x] = pred[(by + y) * 8 + bx + x];
            add_block(C + (size_t)by * cw + bx, cw, p4, resid_slot(mb, coef, 16 + 4 * pl + blk));
        }
    }
}

void recon_cpu_predict_picture(const h264b200_mb_t *mbs, const int16_t *coef, int wm, int hm, uint8_t *cur_frame, uint8_t *const *frames)
{
    planes_t cur = planes_of(cur_frame, wm, hm);
    int mbx, mby;
    /* raster order satisfies every intra dependency (left, up-left, up, up-right) */
    for (mby = 0; mby < hm; mby++) for (mbx = 0; mbx < wm; mbx++) {
        const h264b200_mb_t *mb = &mbs[mby * wm + mbx];
        if (mb->mb_class == H264B200_MB_MISSING || mb->mb_class == H264B200_MB_CONCEAL) continue;
        if (mb->mb_class == H264B200_MB_INTER) inter_mb(mb, coef, mbx, mby, cur, frames, wm, hm);
        else intra_mb(mb, coef, mbx, mby, cur);
    }
}

/* ==================================================== K3c: spatial concealment */
/* ConcealMb's interpolation path + Transform (h264bsd_conceal.c:330-631), driven by the order list the host parser
 * prepared (H264B200_MB_CONCEAL records; avail = usable neighbours at the macroblock's turn) */
static void conceal_plane(uint8_t *P, int st, int x0, int y0, int S, int fl, int luma)
{
    const int g = S / 4, A = fl & H264B200_CN_ABOVE, B = fl & H264B200_CN_BELOW, L = fl & H264B200_CN_LEFT, R = fl & H264B200_CN_RIGHT;
    int a[4] = {0, 0, 0, 0}, b[4] = {0, 0, 0, 0}, l[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0}, fp[16], k, t, j = 0, hor = 0, ver = 0, x, y;
    memset(fp, 0, sizeof fp);
    for (k = 0; k < 4; k++) for (t = 0; t < g; t++) {
        if (A) a[k] += P[(size_t)(y0 - 1) * st + x0 + k * g + t];
        if (B) b[k] += P[(size_t)(y0 + S) * st + x0 + k * g + t];
        if (L) l[k] += P[(size_t)(y0 + k * g + t) * st + x0 - 1];
        if (R) r[k] += P[(size_t)(y0 + k * g + t) * st + x0 + S];
    }
    if (A) { j++; hor++; fp[0] += a[0] + a[1] + a[2] + a[3]; fp[1] += a[0] + a[1] - a[2] - a[3]; }
    if (B) { j++; hor++; fp[0] += b[0] + b[1] + b[2] + b[3]; fp[1] += b[0] + b[1] - b[2] - b[3]; }
    if (L) { j++; ver++; fp[0] += l[0] + l[1] + l[2] + l[3]; fp[4] += l[0] + l[1] - l[2] - l[3]; }
    if (R) { j++; ver++; fp[0] += r[0] + r[1] + r[2] + r[3]; fp[4] += r[0] + r[1] - r[2] - r[3]; }
    if (!hor && L && R) fp[1] = (l[0] + l[1] + l[2] + l[3] - r[0] - r[1] - r[2] - r[3]) >> (luma ? 5 : 4);
    else if (hor) fp[1] >>= ((luma ? 3 : 2) + hor);
    if (!ver && A && B) fp[4] = (a[0] + a[1] + a[2] + a[3] - b[0] - b[1] - b[2] - b[3]) >> (luma ? 5 : 4);
    else if (ver) fp[4] >>= ((luma ? 3 : 2) + ver);
    switch (j) {
    case 1: fp[0] >>= luma ? 4 : 3; break;
    case 2: fp[0] >>= luma ? 5 : 4; break;
    case 3: fp[0] = (21 * fp[0]) >> (luma ? 10 : 9); break;
    default: fp[0] >>= luma ? 6 : 5; break;
    }
    if (!fp[1] && !fp[4]) { for (k = 1; k < 16; k++) fp[k] = fp[0]; }
    else {
        int t0 = fp[0], t1 = fp[1], v = fp[4], c;
        fp[0] = t0 + t1; fp[1] = t0 + (t1 >> 1); fp[2] = t0 - (t1 >> 1); fp[3] = t0 - t1;
        fp[5] = fp[6] = fp[7] = v;
        for (c = 0; c < 4; c++) {
            int u0 = fp[c], u1 = fp[4 + c];
            fp[c] = u0 + u1; fp[4 + c] = u0 + (u1 >> 1); fp[8 + c] = u0 - (u1 >> 1); fp[12 + c] = u0 - u1;
        }
    }
    for (y = 0; y < S; y++) for (x = 0; x < S; x++) P[(size_t)(y0 + y) * st + x0 + x] = (uint8_t)clip255(fp[4 * (y / g) + x / g]);
}

void recon_cpu_conceal_picture(const h264b200_mb_t *mbs, const uint32_t *list, uint32_t n, int wm, int hm, uint8_t *frame)
{
    planes_t f = planes_of(frame, wm, hm);
    uint32_t e;
    for (e = 0; e < n; e++) {
        const int mbx = (int)(list[e] % (uint32_t)wm), mby = (int)(list[e] / (uint32_t)wm), fl = mbs[list[e]].avail;
        conceal_plane(f.y, f.w, mbx * 16, mby * 16, 16, fl, 1);
        conceal_plane(f.cb, f.w / 2, mbx * 8, mby * 8, 8, fl, 0);
        conceal_plane(f.cr, f.w / 2, mbx * 8, mby * 8, 8, fl, 0);
    }
}

/* ==================================================== K4: deblocking */
static inline int mb_intra(const h264b200_mb_t *m) { return m->mb_class != H264B200_MB_INTER || (m->flags & H264B200_MBF_DBK_AS_INTRA); }

/* bS between 4x4 blocks p (in mbp, raster index rp) and q (in mbq, raster rq); mb_edge: edge lies on a MB boundary */
static int boundary_strength(const h264b200_mb_t *mbp, int rp, const h264b200_mb_t *mbq, int rq, int mb_edge)
{
    if (mb_intra(mbp) || mb_intra(mbq)) return mb_edge ? 4 : 3;
    if (((mbp->nz_mask >> H264_RASTER_TO_BLK[rp]) & 1) || ((mbq->nz_mask >> H264_RASTER_TO_BLK[rq]) & 1)) return 2;
    if (mbp->ref_slot[(rp >> 3) * 2 + ((rp & 3) >> 1)] != mbq->ref_slot[(rq >> 3) * 2 + ((rq & 3) >> 1)]) return 1;
    if (iabs(mbp->mv[rp][0] - mbq->mv[rq][0]) >= 4 || iabs(mbp->mv[rp][1] - mbq->mv[rq][1]) >= 4) return 1;
    return 0;
}

/* filter one line of samples across an edge: pix points at q0, step = distance between samples */
static void filter_line(uint8_t *pix, int step, int bs, int alpha, int beta, int tc0, int luma)
{
    int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (iabs(p0 - q0) >= alpha || iabs(p1 - p0) >= beta || iabs(q1 - q0) >= beta) return;
    if (bs < 4) {
        int tc, d;
        if (luma) {
            int p2 = pix[-3 * step], q2 = pix[2 * step], ap = iabs(p2 - p0), aq = iabs(q2 - q0);
            tc = tc0 + (ap < beta) + (aq < beta);
            if (ap < beta) pix[-2 * step] = (uint8_t)(p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
            if (aq < beta) pix[step] = (uint8_t)(q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
        } else tc = tc0 + 1;
        d = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        pix[-step] = (uint8_t)clip255(p0 + d);
        pix[0] = (uint8_t)clip255(q0 - d);
    } else if (luma) {
        int p2 = pix[-3 * step], q2 = pix[2 * step], p3 = pix[-4 * step], q3 = pix[3 * step];
        int small = iabs(p0 - q0) < ((alpha >> 2) + 2);
        if (iabs(p2 - p0) < beta && small) {
            pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (iabs(q2 - q0) < beta && small) {
            pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
        pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

static inline int qpc_for(const h264b200_mb_t *cur, int qp)
{ return H264_QPC[clip3(0, 51, qp + cur->chroma_qp_off)]; }

void recon_cpu_deblock_picture(const h264b200_mb_t *mbs, int wm, int hm, uint8_t *frame)
{
    planes_t f = planes_of(frame, wm, hm);
    int mbx, mby, dir, e, k, i, cw = f.w / 2;
    for (mby = 0; mby < hm; mby++) for (mbx = 0; mbx < wm; mbx++) {
        const h264b200_mb_t *q = &mbs[mby * wm + mbx];
        uint8_t *Y = f.y + (size_t)mby * 16 * f.w + mbx * 16;
        if (q->mb_class == H264B200_MB_MISSING || !q->dbk_flags) continue;
        for (dir = 0; dir < 2; dir++) {              /* 0: vertical edges (left to right), 1: horizontal edges */
            for (e = 0; e < 4; e++) {
                const h264b200_mb_t *p = q;
                int mb_edge = e == 0, bs[4], qp_p, qpav, ia, ib, alpha, beta, any = 0;
                if (mb_edge) {
                    if (!(q->dbk_flags & (dir ? H264B200_DBK_TOP : H264B200_DBK_LEFT))) continue;
                    p = dir ? q - wm : q - 1;
                    if (p->mb_class == H264B200_MB_MISSING) continue;
                } else if (!(q->dbk_flags & H264B200_DBK_INNER)) continue;
                for (k = 0; k < 4; k++) {
                    int rq = dir ? e * 4 + k : k * 4 + e;
                    int rp = mb_edge ? (dir ? 12 + k : k * 4 + 3) : (dir ? rq - 4 : rq - 1);
                    bs[k] = boundary_strength(p, rp, q, rq, mb_edge);
                    any |= bs[k];
                }
                if (!any) continue;
                qp_p = p->qp_dbk;
                /* luma */
                qpav = (qp_p + q->qp_dbk + 1) >> 1;
                ia = clip3(0, 51, qpav + q->dbk_off_a); ib = clip3(0, 51, qpav + q->dbk_off_b);
                alpha = H264_ALPHA[ia]; beta = H264_BETA[ib];
                for (i = 0; i < 16; i++) {
                    int b = bs[i >> 2];
                    uint8_t *pix = dir ? Y + (size_t)(e * 4) * f.w + i : Y + (size_t)i * f.w + e * 4;
                    if (b) filter_line(pix, dir ? f.w : 1, b, alpha, beta, b < 4 ? H264_TC0[ia][b - 1] : 0, 1);
                }
                /* chroma: edges 0 and 2 of the luma grid */
                if (e & 1) continue;
                {
                    int pl;
                    for (pl = 0; pl < 2; pl++) {
                        uint8_t *C = (pl ? f.cr : f.cb) + (size_t)mby * 8 * cw + mbx * 8;
                        int qc_q = qpc_for(q, q->qp_dbk), qc_p = qpc_for(q, qp_p);
                        qpav = (qc_p + qc_q + 1) >> 1;
                        ia = clip3(0, 51, qpav + q->dbk_off_a); ib = clip3(0, 51, qpav + q->dbk_off_b);
                        alpha = H264_ALPHA[ia]; beta = H264_BETA[ib];
                        for (i = 0; i < 8; i++) {
                            int b = bs[i >> 1];
                            uint8_t *pix = dir ? C + (size_t)(e * 2) * cw + i : C + (size_t)i * cw + e * 2;
                            if (b) filter_line(pix, dir ? cw : 1, b, alpha, beta, b < 4 ? H264_TC0[ia][b - 1] : 0, 0);
                        }
                    }
                }
            }
        }
    }
}

/* ==================================================== backend glue */
typedef struct {
    uint32_t wm, hm, n_slots;
    uint8_t *frames[H264_MAX_SLOTS];
    uint8_t *predeblock;         /* copy of the last picture before deblocking (tests) */
    h264_pic_input_t pic;
    uint32_t errors;
    /* device-parse emulation (kp_cpu.cpp): contexts for kp_core, status words per frame slot */
    int dev_parse;
    KpMbCtx *kctx;
    h264b200_picstat_t stat[H264_MAX_SLOTS];
    /* deferred launches, like the CUDA engine's per-instance FIFO (device-parse + batched): pictures wait as blocks until
     * recon_cpu_advance launches the oldest one, held back while an output of its frame slot is unreleased */
    int deferred;
    void *ctx;
    struct cpu_qpic { uint8_t *block; uint32_t used; int cur_slot; uint32_t gate_gen; } *q;
    uint32_t q_cap, q_head, q_n;
    pthread_mutex_t q_mu;           /* the queue is filled by the stream's parser thread and drained by whichever thread drives the engine */
    uint32_t qgen[H264_MAX_SLOTS], lgen[H264_MAX_SLOTS], popped[H264_MAX_SLOTS], released[H264_MAX_SLOTS];
} cpu_inst_t;

static pthread_mutex_t g_reg_mu = PTHREAD_MUTEX_INITIALIZER;
static pthread_mutex_t g_adv_mu = PTHREAD_MUTEX_INITIALIZER;   /* held while recon_cpu_advance works on registered instances: an instance is not destroyed under it */
static cpu_inst_t *g_reg[4096];
static uint32_t g_reg_n;

void kp_cpu_parse_picture(const KpPic *pic, const KpTables *tables);   /* kp_cpu.cpp */
static KpTables *g_kp_tables;

static recon_cpu_tap_t g_tap;
void recon_cpu_set_tap(const recon_cpu_tap_t *t) { if (t) g_tap = *t; else memset(&g_tap, 0, sizeof g_tap); }

static void *cpu_inst_create_ex(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n_slots, int host_parse)
{
    cpu_inst_t *in = (cpu_inst_t *)calloc(1, sizeof *in);
    uint32_t i;
    if (!in || n_slots > H264_MAX_SLOTS) return NULL;
    in->wm = wm; in->hm = hm; in->n_slots = n_slots;
    for (i = 0; i < n_slots; i++) in->frames[i] = (uint8_t *)calloc((size_t)wm * hm, 384);
    in->predeblock = (uint8_t *)calloc((size_t)wm * hm, 384);
    in->pic.mbs = (h264b200_mb_t *)calloc((size_t)wm * hm, sizeof(h264b200_mb_t));
    in->dev_parse = be->parse_mode && !host_parse;    /* h264b200SetHostParse: a host-parsed instance on a device-parse engine */
    in->deferred = in->dev_parse && be->ctx != NULL;      /* engine_shim.c sets ctx on batched engines */
    in->ctx = be->ctx;
    pthread_mutex_init(&in->q_mu, NULL);
    if (in->deferred) {
        pthread_mutex_lock(&g_reg_mu);
        if (g_reg_n < 4096) g_reg[g_reg_n++] = in;
        pthread_mutex_unlock(&g_reg_mu);
    }
    in->pic.coef_cap = in->dev_parse ? KP_COEF_CAP(wm * hm) : wm * hm * 8 + 64;
    in->pic.coef = (int16_t *)malloc((size_t)in->pic.coef_cap * 32);
    if (in->dev_parse) {
        in->kctx = (KpMbCtx *)calloc((size_t)wm * hm, sizeof(KpMbCtx));
        in->pic.block_cap = 1 << 16; in->pic.block = (uint8_t *)malloc(in->pic.block_cap);
        if (!g_kp_tables) { g_kp_tables = (KpTables *)malloc(sizeof(KpTables)); h264_kp_fill_tables(g_kp_tables); }
    }
    return in;
}
static void *cpu_inst_create(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n_slots) { return cpu_inst_create_ex(be, wm, hm, n_slots, 0); }
static void cpu_inst_destroy(h264_backend_t *be, void *inst)
{
    cpu_inst_t *in = (cpu_inst_t *)inst; uint32_t i;
    (void)be;
    if (in->deferred) {
        pthread_mutex_lock(&g_adv_mu);
        pthread_mutex_lock(&g_reg_mu);
        for (i = 0; i < g_reg_n; i++) if (g_reg[i] == in) { g_reg[i] = g_reg[--g_reg_n]; break; }
        pthread_mutex_unlock(&g_reg_mu);
        pthread_mutex_unlock(&g_adv_mu);
        for (i = 0; i < in->q_n; i++) free(in->q[(in->q_head + i) % in->q_cap].block);
        free(in->q);
    }
    for (i = 0; i < in->n_slots; i++) free(in->frames[i]);
    free(in->predeblock); free(in->pic.mbs); free(in->pic.coef); free(in->pic.block); free(in->kctx); free(in);
}
static h264_pic_input_t *cpu_pic_begin(h264_backend_t *be, void *inst) { (void)be; return &((cpu_inst_t *)inst)->pic; }
static int cpu_coef_grow(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_slots)
{
    int16_t *n; uint32_t cap = pic->coef_cap * 2 > min_slots ? pic->coef_cap * 2 : min_slots;
    (void)be; (void)inst;
    n = (int16_t *)realloc(pic->coef, (size_t)cap * 32);
    if (!n) return -1;
    pic->coef = n; pic->coef_cap = cap;
    return 0;
}
static int cpu_reconstruct(cpu_inst_t *in, h264_pic_input_t *pic);
static int cpu_pic_submit(h264_backend_t *be, void *inst, h264_pic_input_t *pic)
{
    cpu_inst_t *in = (cpu_inst_t *)inst;
    (void)be;
    in->qgen[pic->cur_slot]++;
    if (in->deferred) {             /* keep the block; recon_cpu_advance launches it */
        struct cpu_qpic *e;
        pthread_mutex_lock(&in->q_mu);
        if (in->q_n == in->q_cap) {
            uint32_t ncap = in->q_cap ? in->q_cap * 2 : 32, k;
            struct cpu_qpic *nq = (struct cpu_qpic *)calloc(ncap, sizeof *nq);
            for (k = 0; k < in->q_n; k++) nq[k] = in->q[(in->q_head + k) % in->q_cap];
            free(in->q); in->q = nq; in->q_cap = ncap; in->q_head = 0;
        }
        e = &in->q[(in->q_head + in->q_n) % in->q_cap];
        e->block = (uint8_t *)malloc(pic->block_used); memcpy(e->block, pic->block, pic->block_used);
        e->used = pic->block_used; e->cur_slot = pic->cur_slot; e->gate_gen = in->popped[pic->cur_slot];
        __atomic_fetch_add(&in->q_n, 1, __ATOMIC_RELEASE);
        pthread_mutex_unlock(&in->q_mu);
        return 0;
    }
    in->lgen[pic->cur_slot]++;
    return cpu_reconstruct(in, pic);
}
/* one scheduling step for every deferred instance of engine `ctx`: launch its oldest queued picture unless an output of
 * that picture's frame slot is still unreleased.  Returns the number of pictures reconstructed. */
uint32_t recon_cpu_advance(void *ctx)
{
    uint32_t i, launched = 0, n;
    cpu_inst_t *list[4096];
    pthread_mutex_lock(&g_adv_mu);
    pthread_mutex_lock(&g_reg_mu);
    n = g_reg_n; memcpy(list, g_reg, n * sizeof list[0]);
    pthread_mutex_unlock(&g_reg_mu);
    for (i = 0; i < n; i++) {
        cpu_inst_t *in = list[i];
        struct cpu_qpic q;
        h264_pic_input_t tmp;
        if (in->ctx != ctx || !__atomic_load_n(&in->q_n, __ATOMIC_ACQUIRE)) continue;
        pthread_mutex_lock(&in->q_mu);
        q = in->q[in->q_head];
        if ((int32_t)(__atomic_load_n(&in->released[q.cur_slot], __ATOMIC_ACQUIRE) - q.gate_gen) < 0) { pthread_mutex_unlock(&in->q_mu); continue; }
        in->q_head = (in->q_head + 1) % in->q_cap;
        __atomic_fetch_sub(&in->q_n, 1, __ATOMIC_RELEASE);      /* head and count move together: the parser thread appends at head + count */
        pthread_mutex_unlock(&in->q_mu);
        /* the instance's own input buffer belongs to its parser thread, which may be filling the next block right now:
         * reconstruct from a private descriptor (records / slots are only written here) */
        memset(&tmp, 0, sizeof tmp);
        tmp.mbs = in->pic.mbs; tmp.coef = in->pic.coef; tmp.coef_cap = in->pic.coef_cap;
        tmp.block = q.block; tmp.block_used = q.used; tmp.cur_slot = q.cur_slot;
        cpu_reconstruct(in, &tmp);
        free(q.block);
        __atomic_fetch_add(&in->lgen[q.cur_slot], 1, __ATOMIC_RELEASE);
        launched++;
    }
    pthread_mutex_unlock(&g_adv_mu);
    return launched;
}
static int cpu_reconstruct(cpu_inst_t *in, h264_pic_input_t *pic)
{
    uint32_t i, n = in->wm * in->hm;
    if (in->dev_parse) {            /* what kernel Kp does on the GPU: slices -> records, slots, concealment, status */
        KpPic kp; KpResult res;
        memset(&res, 0, sizeof res);
        kp.block = pic->block; kp.mbs = pic->mbs; kp.coef = pic->coef; kp.ctx = in->kctx; kp.coef_cap = pic->coef_cap; kp.pad = 0; kp.res = &res;
        kp_cpu_parse_picture(&kp, g_kp_tables);
        pic->coef_used = res.coef_used; pic->n_intra = res.n_intra; pic->n_inter = res.n_inter; pic->any_deblock = res.any_deblock;
        pic->n_conceal = res.n_conceal; pic->conceal_offset = res.conceal_offset;
        in->stat[pic->cur_slot] = res.stat;
    }
    if (g_tap.records) g_tap.records(g_tap.user, pic->mbs, n, pic->coef, pic->coef_used);
    for (i = 0; i < n; i++) if (recon_cpu_residual_mb(&pic->mbs[i], pic->coef)) in->errors |= 1;
    if (g_tap.residual) g_tap.residual(g_tap.user, pic->coef, pic->coef_used);
    recon_cpu_predict_picture(pic->mbs, pic->coef, (int)in->wm, (int)in->hm, in->frames[pic->cur_slot], in->frames);
    if (pic->n_conceal) recon_cpu_conceal_picture(pic->mbs, (const uint32_t *)(pic->coef + (size_t)pic->conceal_offset * 16), pic->n_conceal, (int)in->wm, (int)in->hm, in->frames[pic->cur_slot]);
    memcpy(in->predeblock, in->frames[pic->cur_slot], (size_t)n * 384);
    if (g_tap.predeblock) g_tap.predeblock(g_tap.user, in->predeblock, (size_t)n * 384);
    recon_cpu_deblock_picture(pic->mbs, (int)in->wm, (int)in->hm, in->frames[pic->cur_slot]);
    return 0;
}
static uint8_t *cpu_frame_host(h264_backend_t *be, void *inst, int slot, uint32_t *err)
{
    cpu_inst_t *in = (cpu_inst_t *)inst;
    (void)be;
    if (err) *err = in->errors;
    return in->frames[slot];
}
static void cpu_destroy(h264_backend_t *be) { (void)be; }
static int cpu_block_grow(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_bytes)
{
    uint32_t cap = pic->block_cap * 2 > min_bytes ? pic->block_cap * 2 : min_bytes;
    uint8_t *n = (uint8_t *)realloc(pic->block, cap);
    (void)be; (void)inst;
    if (!n) return -1;
    pic->block = n; pic->block_cap = cap;
    return 0;
}
static int cpu_frame_status(h264_backend_t *be, void *inst, int slot, h264b200_picstat_t *out)
{
    (void)be; *out = ((cpu_inst_t *)inst)->stat[slot]; return 0;
}
static uint8_t *cpu_frame_host_async(h264_backend_t *be, void *inst, int slot, uint32_t *gen)
{
    cpu_inst_t *in = (cpu_inst_t *)inst;
    (void)be;
    if (gen) *gen = in->qgen[slot];
    in->popped[slot] = in->qgen[slot];
    return in->frames[slot];
}
static int cpu_frame_wait(h264_backend_t *be, void *inst, int slot, uint32_t gen, uint32_t *err)
{
    cpu_inst_t *in = (cpu_inst_t *)inst;
    const uint32_t l = __atomic_load_n(&in->lgen[slot], __ATOMIC_ACQUIRE) & 0xffffffu;
    (void)be;
    if (err) *err = in->errors;
    if (l == (gen & 0xffffffu)) return 0;
    return (int32_t)((l - gen) << 8) < 0 ? 2 : 1;      /* 2: still queued; 1: a later picture already took the slot */
}
/* the CPU backend reconstructs at launch: a launched picture is complete */
static int cpu_frame_state(h264_backend_t *be, void *inst, int slot, uint32_t gen)
{
    int rc = cpu_frame_wait(be, inst, slot, gen, NULL);
    return rc == 2 ? 2 : 0;
}
static void cpu_frame_release(h264_backend_t *be, void *inst, int slot, uint32_t gen)
{
    cpu_inst_t *in = (cpu_inst_t *)inst;
    (void)be;
    if ((int32_t)(gen - in->released[slot]) > 0) __atomic_store_n(&in->released[slot], gen, __ATOMIC_RELEASE);
}
static uint32_t cpu_inst_pending(h264_backend_t *be, void *inst) { (void)be; return __atomic_load_n(&((cpu_inst_t *)inst)->q_n, __ATOMIC_ACQUIRE); }

/* every picture is reconstructed synchronously in pic_submit, so the asynchronous half of the interface is trivial */
h264_backend_t recon_cpu_backend(int device_parse)
{
    h264_backend_t b;
    memset(&b, 0, sizeof b);
    b.inst_create = cpu_inst_create; b.inst_destroy = cpu_inst_destroy; b.pic_begin = cpu_pic_begin; b.coef_grow = cpu_coef_grow;
    b.pic_submit = cpu_pic_submit; b.frame_host = cpu_frame_host; b.destroy = cpu_destroy;
    b.frame_host_async = cpu_frame_host_async; b.frame_wait = cpu_frame_wait;
    b.block_grow = cpu_block_grow; b.frame_status = cpu_frame_status;
    b.frame_release = cpu_frame_release; b.inst_pending = cpu_inst_pending; b.frame_state = cpu_frame_state;
    b.parse_mode = device_parse; b.inst_create_ex = cpu_inst_create_ex;
    return b;
}
static h264_backend_t g_cpu_backend;
h264_backend_t *h264_default_backend(void)
{
    if (!g_cpu_backend.inst_create) {
        const char *pm = getenv("H264B200_PARSE");
        g_cpu_backend = recon_cpu_backend(pm && !strcmp(pm, "device"));
    }
    return &g_cpu_backend;
}
