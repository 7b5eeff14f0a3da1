/* h264_cavlc_inl.h — the CAVLC residual block decoder as an inline function (it runs
 * ~100 000 times per 1080p picture; see h264_cavlc.c for the tables and the reference
 * functions it stands in for: h264bsd_cavlc.c:395-915). */
#ifndef B200_H264_CAVLC_INL_H
#define B200_H264_CAVLC_INL_H
#include "h264_bits.h"

typedef struct { uint8_t len, tc, t1, pad; } ct_entry_t;
extern ct_entry_t g_ct[3][16 * 8];
extern ct_entry_t g_ct_cdc[256];
extern uint8_t g_tz[15][512][2];
extern uint8_t g_tz_cdc[3][8][2];
extern uint8_t g_rb[7][8][2];          /* row 6: zerosLeft > 6, 3-bit codes only (len 0: longer code) */
extern int8_t  g_lvl[7][256][4];      /* {level, bits, next suffixLength, -} */

/* Decode one residual block into out[scan[i]].  `out` (16 x int16) is zeroed here when the block
 * has coefficients and left untouched when TotalCoeff is 0.  nc < 0: chroma DC.
 * Returns TotalCoeff, or -1 on a malformed block. */
static __attribute__((noinline)) int h264_cavlc_block_full(br_t *b, int nc, int max_coeff, int16_t *out, const uint8_t *scan)
{
    int tc, t1, i, sl, zeros_left, pos;
    int level[16];
    uint32_t v;

    /* ---- coeff_token ---- */
    if (b->bits < 32) br_refill(b);
    v = (uint32_t)(b->cache >> 32);
    if (nc < 0) {
        ct_entry_t e = g_ct_cdc[v >> 24];
        if (!e.len) return -1;
        br_skip(b, e.len); tc = e.tc; t1 = e.t1;
    } else if (nc < 8) {
        int lz;
        ct_entry_t e;
        if (nc < 2 && (v >> 31)) { br_skip(b, 1); return 0; }      /* the most frequent token: TotalCoeff 0 */
        if (v < 0x10000u) return -1;              /* more than 15 leading zeros: no such code */
        lz = __builtin_clz(v);
        e = g_ct[(0xaa50 >> (2 * nc)) & 3][lz * 8 + ((v >> (28 - lz)) & 7)];     /* nC 0,1 -> 0; 2,3 -> 1; 4..7 -> 2 */
        if (!e.len) return -1;
        br_skip(b, e.len); tc = e.tc; t1 = e.t1;
    } else {
        v >>= 26; br_skip(b, 6);
        if (v == 3) { tc = 0; t1 = 0; }
        else { tc = (int)(v >> 2) + 1; t1 = (int)(v & 3); if (t1 > tc) return -1; }
    }
    if (tc == 0) return 0;
    if (tc > max_coeff) return -1;
    memset(out, 0, 32);

    /* ---- levels ---- */
    sl = (tc > 10 && t1 < 3) ? 1 : 0;
    {   /* trailing ones: up to three sign bits, decoded without a loop (entries beyond t1 are overwritten below) */
        const uint32_t s = (uint32_t)(b->cache >> 61);          /* next 3 bits; b->bits >= 32 - 16 here */
        level[0] = 1 - (int)((s >> 1) & 2); level[1] = 1 - (int)(s & 2); level[2] = 1 - (int)((s << 1) & 2);
        br_skip(b, t1);
    }
    for (i = t1; i < tc; i++) {
        int lv;
        const int8_t *q;
        if (b->bits < 32) br_refill(b);
        v = (uint32_t)(b->cache >> 32);
        q = g_lvl[sl][v >> 24];
        if (q[1] && i != t1) {                    /* prefix + suffix within 8 bits: level and next suffixLength from the table */
            level[i] = q[0]; br_skip(b, q[1]); sl = q[2];
            continue;
        }
        if (q[1]) {                               /* first level after the trailing ones: levelCode += 2 when t1 < 3 */
            lv = q[0];
            br_skip(b, q[1]);
            if (t1 < 3) lv += lv > 0 ? 1 : -1;
        } else {
            int prefix, code;
            if (v < 0x10000u) return -1;          /* level_prefix > 15: not Baseline (h264bsd_cavlc.c:513-514) */
            prefix = __builtin_clz(v);
            br_skip(b, prefix + 1);
            code = (prefix < 15 ? prefix : 15) << sl;
            if (sl > 0 || prefix >= 14) {
                int size = (prefix == 14 && sl == 0) ? 4 : prefix >= 15 ? 12 : sl;
                code += (int)br_get(b, size);
            }
            if (prefix >= 15 && sl == 0) code += 15;
            if (i == t1 && t1 < 3) code += 2;
            lv = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
        }
        level[i] = lv;
        if (sl == 0) sl = 1;
        if ((lv < 0 ? -lv : lv) > (3 << (sl - 1)) && sl < 6) sl++;
    }

    /* ---- total_zeros ---- */
    if (tc < max_coeff) {
        const uint8_t *e = nc < 0 ? g_tz_cdc[tc - 1][br_peek(b, 3)] : g_tz[tc - 1][br_peek(b, 9)];
        if (!e[0]) return -1;
        br_skip(b, e[0]); zeros_left = e[1];
        if (zeros_left + tc > max_coeff) return -1;
    } else zeros_left = 0;

    /* ---- run_before + placement (highest frequency first) ---- */
    pos = zeros_left + tc - 1;
    for (i = 0; i < tc - 1 && zeros_left > 0; i++) {
        int run;
        out[scan[pos]] = (int16_t)level[i];
        {
            const uint8_t *e = g_rb[(zeros_left < 7 ? zeros_left : 7) - 1][br_peek(b, 3)];
            if (e[0]) { br_skip(b, e[0]); run = e[1]; }
            else {                               /* zerosLeft > 6 and 000 prefix: unary tail, up to 11 bits */
                int lz;
                v = br_peek(b, 11);
                if (!v) return -1;
                lz = __builtin_clz(v) - 21;      /* leading zeros within the 11 bits */
                run = lz + 4; br_skip(b, lz + 1);
            }
        }
        if (run > zeros_left) return -1;
        zeros_left -= run;
        pos -= run + 1;
    }
    for (; i < tc; i++) out[scan[pos--]] = (int16_t)level[i];       /* no zeros left: contiguous */
    return tc;
}
/* Two blocks out of three are empty in typical streams, and with sparse neighbours (nC < 2) that is the single
 * bit '1': answered at the call site, without the call into the full decoder. */
static inline int h264_cavlc_block(br_t *b, int nc, int max_coeff, int16_t *out, const uint8_t *scan)
{
    if ((unsigned)nc < 2u) {
        if (b->bits < 1) br_refill(b);
        if (b->cache >> 63) { br_skip(b, 1); return 0; }
    }
    return h264_cavlc_block_full(b, nc, max_coeff, out, scan);
}
#endif
