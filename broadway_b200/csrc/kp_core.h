/* kp_core.h — slice_data() parsing for kernel Kp: bitstream -> macroblock records + coefficient slots, ON THE DEVICE.
 *
 * Same job, same output bytes as the host parser (h264_slice.c, h264_cavlc_inl.h), i.e. the reference's
 * h264bsdDecodeSliceData (h264bsd_slice_data.c:85-235), h264bsdDecodeMacroblockLayer with mb_pred / sub_mb_pred /
 * residual / nC (h264bsd_macroblock_layer.c:133-242, :353-496, :699-869), h264bsdDecodeResidualBlockCavlc
 * (h264bsd_cavlc.c:395-915), motion vector prediction (h264bsd_inter_prediction.c:499-1031), Intra4x4PredMode
 * derivation (h264bsd_intra_prediction.c:1885-1936), the QP update (h264bsd_macroblock_layer.c:1043-1049),
 * h264bsdMarkSliceCorrupted (h264bsd_slice_data.c:302-358) and the bookkeeping half of h264bsdConceal
 * (h264bsd_conceal.c:125-255) — but organised for a GPU:
 *
 *   - one WARP per picture (slices of a picture in order, like the host); the bit-serial syntax walk runs on lane 0,
 *     all 32 lanes move data: the four neighbour records + contexts are staged into shared memory with one
 *     coalesced load per macroblock, the finished 128-byte record, the 32-byte context and the coefficient slots
 *     go out as 16-byte stores, I_PCM samples are copied by the whole warp;
 *   - the slot staging area is kept all-zero between macroblocks, so lane 0 only ever writes non-zero levels;
 *   - the bit reader refills a 64-bit window with ALIGNED 32-bit words (the host pads every RBSP to 16 bytes);
 *   - code tables (the host parser's own LUTs, built once on the host) live in shared memory.
 *
 * The file compiles twice: under nvcc for kp_parse.cuh (KP_LANES 32), and as plain C++ (KP_LANES 1) for the CPU
 * test-suite, where tests/ drive it through oracle/engine_shim to check "device records == host records" and the
 * golden MD5s without a GPU.  The product never runs the CPU build.
 */
#ifndef B200_KP_CORE_H
#define B200_KP_CORE_H
#include <stdint.h>
#include <string.h>
#include "h264b200_records.h"
#include "h264b200_slices.h"

#ifdef __CUDACC__
#define KP_FN __device__ __forceinline__
#define KP_NOINL __device__ __noinline__
#define KP_HOT __device__ __forceinline__   /* hot: inlined so that the bit reader state stays in registers */
#define KP_LANES 32
#define KP_NOUNROLL _Pragma("unroll 1")   /* code size: the hot path has to live in the instruction cache */
#define KP_SYNC() __syncwarp()
#define KP_BCAST(x) __shfl_sync(0xffffffffu, (x), 0)
#define KP_CLZ(x) __clz((int)(x))
#define KP_CTZ(x) (__ffs((int)(x)) - 1)
#define KP_BSWAP(x) __byte_perm((x), 0, 0x0123)
#define KP_COLD() asm volatile("")   /* an opaque statement keeps ptxas from if-converting the block it sits in */
#else
#define KP_FN static inline
#define KP_NOINL static
#define KP_HOT static inline
#define KP_LANES 1
#define KP_NOUNROLL
#define KP_SYNC() ((void)0)
#define KP_BCAST(x) (x)
#define KP_CLZ(x) __builtin_clz(x)
#define KP_CTZ(x) __builtin_ctz(x)
#define KP_BSWAP(x) __builtin_bswap32(x)
#define KP_COLD() ((void)0)
#endif

#include "kp_types.h"
#include <stddef.h>

#ifdef __CUDACC__
/* A pointer into shared memory whose 32-bit shared-window address the compiler must keep in a register from here on.
 * On sm_100 that address contains the CTA's rank in its cluster (S2R SR_CgaCtaId + LEA); inside the divergent lane-0
 * code ptxas cannot use the uniform datapath for it and, with 64 registers, re-materialised the whole sequence (and the
 * warp's staging address from SR_TID) at ~35 places per macroblock: 8 % of the instructions of the kernel. */
template <typename T> __device__ __forceinline__ T *kp_pin_shared(T *p)
{
    uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("" : "+r"(a));
    return (T *)__cvta_shared_to_generic(a);
}
typedef uint4 KpU4;
#else
#define kp_pin_shared(p) (p)
typedef struct { uint32_t x, y, z, w; } KpU4;
#endif

/* lane-0 state of the slice being parsed */
typedef struct {
    /* bit reader: a window of three aligned 32-bit words (big endian in registers): w0 holds the current position,
     * `sh` bits of it are consumed; w1 follows; w2 is already on its way from memory when w1 becomes w0, so the load
     * latency never sits on the decoding chain.  32 valid bits are always available to kp_peek32 (one funnel shift). */
    const uint32_t *words; uint32_t n_words, wpos; uint32_t w0, w1, w2, sh;
    uint32_t rbsp_bits, payload_bits;
    const KpTables *T; KpStage *st;
    const h264b200_slice_t *sl;
    uint32_t W, N, addr; int mbx, mby;
    int is_p, qp;
    uint32_t avail;              /* bit 0..3: macroblock A, B, C, D belongs to this slice (staged in st->nctx / nrec [0..3]) */
    uint32_t n_slots;            /* slots of the current macroblock */
    uint32_t coef_used, n_intra, n_inter, any_deblock;
    int32_t ipcm_byte;           /* >= 0: the macroblock is I_PCM, its 384 samples start at this RBSP byte */
    uint32_t skip_run; int prev_skipped;
} KpS;

/* ------------------------------------------------------------------ bits */
#ifdef __CUDACC__
#define KP_FUNNEL(hi, lo, n) __funnelshift_l((lo), (hi), (n))          /* upper 32 bits of hi:lo << (n & 31) */
#else
#define KP_FUNNEL(hi, lo, n) (((n) & 31) ? ((hi) << ((n) & 31)) | ((lo) >> (32 - ((n) & 31))) : (hi))
#endif
KP_FN uint32_t kp_word(const KpS &s, uint32_t i) { return i < s.n_words ? KP_BSWAP(s.words[i]) : 0u; }   /* zeros behind the RBSP, like the host reader */
KP_FN void kp_seek(KpS &s, uint32_t bit)
{
    const uint32_t wi = bit >> 5;
    s.w0 = kp_word(s, wi); s.w1 = kp_word(s, wi + 1); s.w2 = kp_word(s, wi + 2);
    s.wpos = wi + 3; s.sh = bit & 31;
}
KP_FN void kp_bits_init(KpS &s, const uint8_t *rbsp, uint32_t len, uint32_t bit_off, uint32_t payload_bits)
{
    s.words = (const uint32_t *)rbsp; s.n_words = (len + 3) >> 2;     /* the block pads every RBSP with zeros to 16 bytes */
    s.rbsp_bits = len * 8; s.payload_bits = payload_bits;
    kp_seek(s, bit_off);
}
KP_FN uint32_t kp_pos(const KpS &s) { return (s.wpos - 3) * 32u + s.sh; }
KP_FN uint32_t kp_peek32(const KpS &s) { return KP_FUNNEL(s.w0, s.w1, s.sh); }
KP_FN uint32_t kp_peek(const KpS &s, int n) { return kp_peek32(s) >> (32 - n); }   /* 1 <= n <= 32 */
KP_FN void kp_skip(KpS &s, int n)                                      /* 0 <= n <= 32 */
{
    s.sh += (uint32_t)n;
    if (s.sh >= 32) {                                              /* once per 32 bits: a branch, not a predicated block in every skip */
        KP_COLD();
        s.sh -= 32; s.w0 = s.w1; s.w1 = s.w2; s.w2 = kp_word(s, s.wpos); s.wpos++;
    }
}
KP_FN uint32_t kp_get(KpS &s, int n) { uint32_t v; if (n == 0) return 0; v = kp_peek(s, n); kp_skip(s, n); return v; }
KP_FN uint32_t kp_get1(KpS &s) { const uint32_t v = kp_peek32(s) >> 31; kp_skip(s, 1); return v; }
KP_FN uint32_t kp_ue(KpS &s)     /* 0xffffffff: malformed (h264_bits.h br_ue) */
{
    uint32_t v = kp_peek32(s); int lz;
    if (v & 0x80000000u) { kp_skip(s, 1); return 0; }
    if (v == 0) { kp_skip(s, 32); return 0xffffffffu; }
    lz = KP_CLZ(v);
    if (lz <= 15) { v >>= (31 - 2 * lz); kp_skip(s, 2 * lz + 1); return v - 1; }
    kp_skip(s, lz);
    v = kp_get(s, lz + 1);
    return v - 1;
}
KP_FN int32_t kp_se(KpS &s)
{
    uint32_t k = kp_ue(s);
    if (k == 0xffffffffu) return INT32_MIN;
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
}
KP_FN int kp_more_data(const KpS &s) { return kp_pos(s) < s.payload_bits; }
KP_FN int kp_overrun(const KpS &s) { return kp_pos(s) > s.rbsp_bits; }

/* ------------------------------------------------------------------ CAVLC block (h264_cavlc_inl.h) */
/* One residual block whose coeff_token is not the single bit of an empty block.  Levels go to out[scan[i]]; `out` is
 * all zero on entry.  Returns TotalCoeff or -1. */
KP_HOT int kp_cavlc_block_full(KpS &s, int nc, int max_coeff, int16_t *out, const uint8_t *scan)
{
    const KpTables *T = s.T;
    int16_t *level = s.st->lvl;
    int tc, t1, i, sl, zeros_left, pos;
    uint32_t v = kp_peek32(s);

    if (nc < 0) {
        const uint32_t e = *(const uint32_t *)T->ct_cdc[v >> 24];
        if (!(e & 0xff)) return -1;
        kp_skip(s, (int)(e & 0xff)); tc = (int)((e >> 8) & 0xff); t1 = (int)((e >> 16) & 0xff);
    } else if (nc < 8) {
        int lz; uint32_t e;
        if (v < 0x10000u) return -1;
        lz = KP_CLZ(v);
        e = *(const uint32_t *)T->ct[(0xaa50 >> (2 * nc)) & 3][lz * 8 + ((v >> (28 - lz)) & 7)];
        if (!(e & 0xff)) return -1;
        kp_skip(s, (int)(e & 0xff)); tc = (int)((e >> 8) & 0xff); t1 = (int)((e >> 16) & 0xff);
    } else {
        v >>= 26; kp_skip(s, 6);
        if (v == 3) { tc = 0; t1 = 0; }
        else { tc = (int)(v >> 2) + 1; t1 = (int)(v & 3); if (t1 > tc) return -1; }
    }
    if (tc == 0) return 0;
    if (tc > max_coeff) return -1;

    sl = (tc > 10 && t1 < 3) ? 1 : 0;
    {
        const uint32_t sg = kp_peek32(s) >> 29;                  /* the signs of up to three trailing ones */
        level[0] = (int16_t)(1 - (int)((sg >> 1) & 2)); level[1] = (int16_t)(1 - (int)(sg & 2)); level[2] = (int16_t)(1 - (int)((sg << 1) & 2));
        kp_skip(s, t1);
    }
KP_NOUNROLL
    for (i = t1; i < tc; i++) {
        int lv;
        uint32_t q;                                              /* {level, bits, next suffixLength} */
        v = kp_peek32(s);
        q = *(const uint32_t *)T->lvl[sl][v >> 24];
        if ((q & 0xff00u) && i != t1) { level[i] = (int16_t)(int8_t)q; kp_skip(s, (int)((q >> 8) & 0xff)); sl = (int)((q >> 16) & 0xff); continue; }
        if (q & 0xff00u) {
            lv = (int8_t)q;
            kp_skip(s, (int)((q >> 8) & 0xff));
            if (t1 < 3) lv += lv > 0 ? 1 : -1;
        } else {
            int prefix, code;
            if (v < 0x10000u) return -1;
            prefix = KP_CLZ(v);
            kp_skip(s, prefix + 1);
            code = (prefix < 15 ? prefix : 15) << sl;
            if (sl > 0 || prefix >= 14) {
                int size = (prefix == 14 && sl == 0) ? 4 : prefix >= 15 ? 12 : sl;
                code += (int)kp_get(s, size);
            }
            if (prefix >= 15 && sl == 0) code += 15;
            if (i == t1 && t1 < 3) code += 2;
            lv = (code & 1) ? (-code - 1) >> 1 : (code + 2) >> 1;
        }
        level[i] = (int16_t)lv;
        if (sl == 0) sl = 1;
        if ((lv < 0 ? -lv : lv) > (3 << (sl - 1)) && sl < 6) sl++;
    }

    if (tc < max_coeff) {
        const uint8_t *e = nc < 0 ? T->tz_cdc[tc - 1][kp_peek(s, 3)] : T->tz[tc - 1][kp_peek(s, 9)];
        const uint32_t ev = *(const uint16_t *)e;
        if (!(ev & 0xff)) return -1;
        kp_skip(s, (int)(ev & 0xff)); zeros_left = (int)(ev >> 8);
        if (zeros_left + tc > max_coeff) return -1;
    } else zeros_left = 0;

    pos = zeros_left + tc - 1;
KP_NOUNROLL
    for (i = 0; i < tc - 1 && zeros_left > 0; i++) {
        int run;
        out[scan[pos]] = level[i];
        {
            const uint32_t pk = kp_peek(s, 11);
            const uint32_t ev = *(const uint16_t *)T->rb[(zeros_left < 7 ? zeros_left : 7) - 1][pk >> 8];
            if (ev & 0xff) { kp_skip(s, (int)(ev & 0xff)); run = (int)(ev >> 8); }
            else {
                int lz;
                if (!pk) return -1;
                lz = KP_CLZ(pk) - 21;
                run = lz + 4; kp_skip(s, lz + 1);
            }
        }
        if (run > zeros_left) return -1;
        zeros_left -= run;
        pos -= run + 1;
    }
KP_NOUNROLL
    for (; i < tc; i++, pos--) out[scan[pos]] = level[i];
    return tc;
}
/* ------------------------------------------------------------------ motion vector prediction (h264_slice.c predict_mv) */
/* The vectors and reference indices a prediction can look at form a 5 x 6 grid (KpStage.mvg / refg): row 0 = the row
 * of 4x4 blocks above the macroblock (D, B0..B3, C), column 0 = the column to its left (A0..A3), the rest = the
 * macroblock itself.  The border is filled by the whole warp when the neighbours are staged (kp_stage_derive); the
 * interior starts "not available" and is filled as the partitions are derived, which is exactly the availability rule
 * inside a macroblock.  refg: -2 not available, -1 available but not inter, else refIdxL0. */
#define KP_G(r, c) ((r) * 12 + 3 + (c))      /* row stride 12: the four interior entries of a row start 16-byte aligned */
KP_FN int kp_median3(int a, int b, int c) { int mx = a > b ? a : b, mn = a < b ? a : b; return c > mx ? mx : c < mn ? mn : c; }

/* dir: 0 median, 1 A first (8x16 left / 16x8 bottom), 2 B first (16x8 top), 3 C first (8x16 right).  Returns {hor, ver} packed. */
KP_FN uint32_t kp_predict_mv(const KpStage *st, int x4, int y4, int w4, int ref, int dir)
{
    const int ia = KP_G(y4 + 1, x4), ib = KP_G(y4, x4 + 1);
    int ic = KP_G(y4, x4 + 1 + w4);
    const int ra = st->refg[ia], rb = st->refg[ib];
    int rc = st->refg[ic];
    if (rc == -2) { ic = KP_G(y4, x4); rc = st->refg[ic]; }      /* C not available: D */
    const uint32_t ma = st->mvg[ia], mb = st->mvg[ib], mc = st->mvg[ic];
    if (dir == 1 && ra == ref) return ma;
    if (dir == 2 && rb == ref) return mb;
    if (dir == 3 && rc == ref) return mc;
    if (rb != -2 || rc != -2 || ra == -2) {
        const int ea = ra == ref, eb = rb == ref, ec = rc == ref;
        if (ea + eb + ec != 1) {
            const int mx = kp_median3((int16_t)(ma & 0xffff), (int16_t)(mb & 0xffff), (int16_t)(mc & 0xffff));
            const int my = kp_median3((int32_t)ma >> 16, (int32_t)mb >> 16, (int32_t)mc >> 16);
            return (uint32_t)(uint16_t)mx | ((uint32_t)(uint16_t)my << 16);
        }
        return ea ? ma : eb ? mb : mc;
    }
    return ma;
}
KP_FN int kp_mv_in_range(int x, int y) { return x >= -8192 && x <= 8191 && y >= -2048 && y <= 2047; }

/* ------------------------------------------------------------------ residual */
/* nC (9.2.1) from TotalCoeff grids with one guard row above and one guard column to the left (KpStage.grid: luma, 5 rows
 * of 8, cells 0..39; one 3 x 4 grid per chroma plane, cells 40..51 and 52..63), the host parser's own layout (h264_slice.c parse_residual): the
 * guards come from macroblocks A and B (64 = not available, so that a + b >= 64 exactly when a neighbour is missing) and
 * are written by the whole warp when the neighbours are staged; the interior starts at zero. */
KP_FN int kp_nc_avg(int a, int b)
{
    int n = a + b;
    if (n < 64) n = (n + 1) >> 1;
    return n & 31;
}

/* residual( ) of 7.3.5.3 as ONE loop with ONE call of the block decoder (code size: the kernel has to live in the
 * instruction cache), over exactly the blocks the syntax holds: bit `step` of `todo` — step 0 = Intra16x16 DC, 1..16 =
 * luma4x4BlkIdx 0..15 (set by coded_block_pattern per 8x8 quadrant), 17/18 = chroma DC Cb/Cr, 19..26 = chroma AC.  What a
 * step needs (its cell in the TotalCoeff grids, the distance to the cell above, its kind, its bit in resid_mask, its scan
 * table and coefficient count) is one 64-bit entry of KpTables.step_desc, so the loop head has no branches.  Slot order and masks as include/h264b200_records.h says.  Returns 0 / -1. */
KP_HOT int kp_parse_residual(KpS &s, int cbp, int i16)
{
    KpStage *st = s.st;
    h264b200_mb_t *r = &st->rec;
    const KpTables *T = s.T;
    uint32_t slot = 0, mask = 0, nz = 0, cdc = 0;                 /* cdc: bit 0 / 1: the Cb / Cr DC block has coefficients */
    int dc_nz = 0;
    uint32_t todo = ((uint32_t)T->cbp_luma[cbp & 15] << 1) | (uint32_t)i16;
    if (cbp & 0x30) todo |= (cbp & 0x20) ? 0x7fe0000u : 0x60000u;
    r->coef_offset = s.coef_used;
KP_NOUNROLL
    while (todo) {
        const int step = KP_CTZ(todo);
        const uint32_t dsc = T->step_desc[step][0], dsc2 = T->step_desc[step][1];
        const int kind = (int)((dsc >> 12) & 3);
        const int lum16 = kind == 1 ? i16 : 0;                    /* a luma block of an Intra16x16 macroblock: 15 coefficients from scan position 1 */
        uint8_t *g = st->grid + (dsc & 0xff);
        const int maxc = (int)((dsc2 >> 16) & 0xff) - lum16;
        const uint8_t *scan = (const uint8_t *)T + (dsc2 & 0xffffu) + lum16;
        int16_t *out = st->slots + (uint32_t)((kind == 1 && dc_nz) ? step : (int)slot) * 16 + (dsc2 >> 24);   /* Intra16x16 with DC: slot 1 + luma4x4BlkIdx */
        int nc = kp_nc_avg(g[-1], *(g - ((dsc >> 8) & 15))), tc;
        if (kind == 2) nc = -1;
        todo &= todo - 1;
        /* two blocks out of three are empty, and with sparse neighbours that is the single bit '1' */
        if ((unsigned)nc < 2u && (int32_t)kp_peek32(s) < 0) { kp_skip(s, 1); tc = 0; }
        else {
            tc = kp_cavlc_block_full(s, nc, maxc, out, scan);
            if (tc < 0) return -1;
        }
        {
            const uint32_t bit = 1u << ((dsc >> 16) & 31);
            if (kind == 1) {
                if (tc) { nz |= bit; *g = (uint8_t)tc; if (!dc_nz) { mask |= bit; slot++; } }
            } else if (kind == 3) {
                if (tc) *g = (uint8_t)tc;
                if (tc || ((cdc >> ((step - 19) >> 2)) & 1)) { mask |= bit; slot++; }
            } else if (kind == 2) {
                if (tc) cdc |= 1u << (step - 17);
                if (step == 18 && cdc) {
                    mask |= H264B200_RESID_CHROMA_DC; slot++;
                    if (!(cbp & 0x20)) {                           /* DC only: an all-zero slot per block of a plane whose DC is coded */
                        if (cdc & 1) { mask |= 0xfu << 16; slot += 4; }
                        if (cdc & 2) { mask |= 0xfu << 20; slot += 4; }
                    }
                }
            } else if (tc) { dc_nz = 1; mask |= H264B200_RESID_LUMA_DC | 0xffffu; slot = 17; }   /* every luma block gets a slot (staging is zero) */
        }
    }
    r->resid_mask = mask; r->nz_mask = (uint16_t)nz;
    s.n_slots = slot;
    return 0;
}

/* ------------------------------------------------------------------ intra */
KP_FN int kp_intra_usable(const KpS &s, int nb)
{
    return ((s.avail >> nb) & 1) && !(s.sl->constrained_intra && s.st->nctx[nb].kind == H264B200_MB_INTER);
}
KP_FN int kp_pred_i4_mode(const KpS &s, int blk, int aA, int aB)
{
    const uint8_t *r2b = s.T->raster_to_blk;
    int r = r2b[blk], x4 = r & 3, y4 = r >> 2, ma, mb;
    if (x4 > 0) ma = s.st->rec.i4_mode[r2b[r - 1]];
    else {
        if (!aA) return 2;
        ma = s.st->nctx[0].kind == H264B200_MB_I4x4 ? s.st->nrec[0].i4_mode[r2b[r + 3]] : 2;
    }
    if (y4 > 0) mb = s.st->rec.i4_mode[r2b[r - 4]];
    else {
        if (!aB) return 2;
        mb = s.st->nctx[1].kind == H264B200_MB_I4x4 ? s.st->nrec[1].i4_mode[r2b[12 + x4]] : 2;
    }
    return ma < mb ? ma : mb;
}
KP_FN int kp_update_qp(KpS &s, int delta)
{
    if (delta < -26 || delta > 25) return -1;
    if (delta) { s.qp += delta; if (s.qp < 0) s.qp += 52; else if (s.qp >= 52) s.qp -= 52; }
    return 0;
}
KP_FN void kp_set_qp_fields(const KpS &s, h264b200_mb_t *r)
{
    int qc = s.qp + s.sl->chroma_qp_off;
    qc = qc < 0 ? 0 : qc > 51 ? 51 : qc;
    r->qp_y = (uint8_t)s.qp; r->qp_dbk = (uint8_t)s.qp; r->qp_c = s.T->qpc[qc];
}

/* mb_pred of an intra macroblock (mb_type 0 I4x4, 1..24 I16x16; I_PCM is handled by the caller): modes + legality.
 * Returns coded_block_pattern (>= 0), -1 on an error, -2 when coded_block_pattern is still to be read (I4x4). */
KP_HOT int kp_intra_pred(KpS &s, uint32_t mb_type)
{
    h264b200_mb_t *r = &s.st->rec; KpMbCtx *c = &s.st->ctx;
    const int aA = kp_intra_usable(s, 0), aB = kp_intra_usable(s, 1), aC = kp_intra_usable(s, 2), aD = kp_intra_usable(s, 3);
    uint32_t v;
    r->avail = (uint8_t)((aA ? H264B200_AVAIL_A : 0) | (aB ? H264B200_AVAIL_B : 0) | (aC ? H264B200_AVAIL_C : 0) | (aD ? H264B200_AVAIL_D : 0));
    if (mb_type == 0) {
        const uint8_t *r2b = s.T->raster_to_blk;
        c->kind = r->mb_class = H264B200_MB_I4x4;
KP_NOUNROLL
        for (int blk = 0; blk < 16; blk++) {
            int pred = kp_pred_i4_mode(s, blk, aA, aB), mode;
            if (kp_get1(s)) mode = pred;
            else { int rem = (int)kp_get(s, 3); mode = rem < pred ? rem : rem + 1; }
            r->i4_mode[blk] = (uint8_t)mode;
            {
                const int rr = r2b[blk], x4 = rr & 3, y4 = rr >> 2;
                const int left = x4 > 0 ? 1 : aA, up = y4 > 0 ? 1 : aB;
                const int ul = (x4 > 0 && y4 > 0) ? 1 : x4 > 0 ? aB : y4 > 0 ? aA : aD;
                /* modes 0,3,7 need up; 1,8 left; 4,5,6 all three (h264bsd_intra_prediction.c:773-823) */
                const int need_up = (0x0f9 >> mode) & 1, need_left = (0x172 >> mode) & 1, need_ul = (0x070 >> mode) & 1;
                if ((need_up && !up) || (need_left && !left) || (need_ul && !ul)) return -1;
            }
        }
    } else {
        c->kind = r->mb_class = H264B200_MB_I16x16;
        r->i16_mode = (uint8_t)((mb_type - 1) & 3);
        if ((r->i16_mode == 0 && !aB) || (r->i16_mode == 1 && !aA) || (r->i16_mode == 3 && (!aA || !aB || !aD))) return -1;
    }
    v = kp_ue(s); if (v > 3) return -1;
    r->chroma_mode = (uint8_t)v;
    if ((v == 1 && !aA) || (v == 2 && !aB) || (v == 3 && (!aA || !aB || !aD))) return -1;
    if (mb_type == 0) return -2;
    return (int)((((mb_type - 1) >> 2) % 3) << 4 | (mb_type >= 13 ? 15 : 0));
}

/* ------------------------------------------------------------------ inter */
KP_FN int kp_read_ref_idx(KpS &s, uint32_t n_active)
{
    uint32_t v;
    if (n_active <= 1) return 0;
    v = n_active - 1 > 1 ? kp_ue(s) : !kp_get1(s);               /* te(v) */
    if (v >= n_active) return -1;
    return (int)v;
}

/* partitions of a P macroblock as a list (KpStage.part): x4 | y4 << 2 | (w4-1) << 4 | (h4-1) << 6 | quadrant << 8 | dir << 10 */
#define KP_PART(x, y, w, h, q, d) ((uint16_t)((x) | ((y) << 2) | (((w) - 1) << 4) | (((h) - 1) << 6) | ((q) << 8) | ((d) << 10)))

/* mb_pred / sub_mb_pred + vector derivation of a P macroblock; mb_type 0..4, or 5 = P_Skip (nothing is read).  0 / -1. */
KP_HOT int kp_inter_pred(KpS &s, uint32_t mb_type)
{
    KpStage *st = s.st;
    h264b200_mb_t *r = &st->rec; KpMbCtx *c = &st->ctx;
    const uint32_t n_active = s.sl->num_ref_idx_active;
    uint16_t *part = st->part;
    int n_part, n_ref, skip_zero = 0;
    c->kind = r->mb_class = H264B200_MB_INTER;
    /* the partition list, and how many ref_idx are read (one per 16x16 / 16x8 / 8x16 partition or 8x8 quadrant) */
    if (mb_type == 0 || mb_type == 5) { part[0] = KP_PART(0, 0, 4, 4, 0, 0); n_part = 1; n_ref = 1; r->part_flags = 31; }
    else if (mb_type == 1) { part[0] = KP_PART(0, 0, 4, 2, 0, 2); part[1] = KP_PART(0, 2, 4, 2, 2, 1); n_part = 2; n_ref = 2; r->part_flags = 15; }
    else if (mb_type == 2) { part[0] = KP_PART(0, 0, 2, 4, 0, 1); part[1] = KP_PART(2, 0, 2, 4, 1, 3); n_part = 2; n_ref = 2; r->part_flags = 15; }
    else {
        n_part = 0; n_ref = 4;
KP_NOUNROLL
        for (int q = 0; q < 4; q++) {
            const uint32_t v = kp_ue(s);
            const int ox = (q & 1) * 2, oy = (q >> 1) * 2;
            if (v > 3) return -1;
            if (v == 0) { part[n_part++] = KP_PART(ox, oy, 2, 2, q, 0); r->part_flags |= (uint8_t)(1 << q); }
            else if (v == 1) { part[n_part++] = KP_PART(ox, oy, 2, 1, q, 0); part[n_part++] = KP_PART(ox, oy + 1, 2, 1, q, 0); }
            else if (v == 2) { part[n_part++] = KP_PART(ox, oy, 1, 2, q, 0); part[n_part++] = KP_PART(ox + 1, oy, 1, 2, q, 0); }
            else {
                part[n_part++] = KP_PART(ox, oy, 1, 1, q, 0); part[n_part++] = KP_PART(ox + 1, oy, 1, 1, q, 0);
                part[n_part++] = KP_PART(ox, oy + 1, 1, 1, q, 0); part[n_part++] = KP_PART(ox + 1, oy + 1, 1, 1, q, 0);
            }
        }
    }
    /* ref_idx_l0: quadrants covered by reference k are  16x16: all;  16x8: 2k, 2k+1;  8x16: k, k+2;  8x8: k */
KP_NOUNROLL
    for (int k = 0; k < n_ref; k++) {
        const int ref = (mb_type >= 4) ? 0 : kp_read_ref_idx(s, n_active);
        const int slot = (ref >= 0 && ref <= 16) ? s.sl->ref_slot[ref] : -1;
        const uint32_t qm = n_ref == 1 ? 15u : n_ref == 4 ? 1u << k : mb_type == 1 ? 3u << (2 * k) : 5u << k;
        if (ref < 0 || slot < 0) return -1;                       /* bad index, or a missing / non-existing picture (h264bsd_dpb.c:846-860) */
        for (int q = 0; q < 4; q++) if ((qm >> q) & 1) { c->ref_idx[q] = (int8_t)ref; r->ref_slot[q] = (uint8_t)slot; }
    }
    /* mvd_l0, in partition order, before any vector is derived */
    if (mb_type == 5) {
        /* P_Skip: zero vector when A or B is missing or is a zero vector to reference 0 (h264bsd_inter_prediction.c:521-527) */
        const int ra = st->refg[KP_G(1, 0)], rb = st->refg[KP_G(0, 1)];
        st->mvd[0] = 0;
        skip_zero = ra == -2 || rb == -2 || (ra == 0 && st->mvg[KP_G(1, 0)] == 0) || (rb == 0 && st->mvg[KP_G(0, 1)] == 0);
    } else {
KP_NOUNROLL
        for (int k = 0; k < n_part; k++) {
            const uint32_t dx = (uint32_t)kp_se(s), dy = (uint32_t)kp_se(s);
            st->mvd[k] = (dx & 0xffffu) | (dy << 16);             /* only the low 16 bits reach the int16 vector */
        }
    }
KP_NOUNROLL
    for (int k = 0; k < n_part; k++) {
        const uint32_t pt = part[k];
        const int x4 = pt & 3, y4 = (pt >> 2) & 3, w4 = ((pt >> 4) & 3) + 1, h4 = ((pt >> 6) & 3) + 1;
        const int ref = c->ref_idx[(pt >> 8) & 3];
        const uint32_t pred = skip_zero ? 0u : kp_predict_mv(st, x4, y4, w4, ref, (pt >> 10) & 3);
        const int mx = (int16_t)((pred & 0xffffu) + (st->mvd[k] & 0xffffu)), my = (int16_t)((pred >> 16) + (st->mvd[k] >> 16));
        const uint32_t mv = (uint32_t)(uint16_t)mx | ((uint32_t)(uint16_t)my << 16);
        if (!kp_mv_in_range(mx, my)) return -1;
        /* the interior of the grid is also the record's vector array (kp_stage_out copies its rows) */
KP_NOUNROLL
        for (int j = y4; j < y4 + h4; j++) {
            uint32_t *row = &st->mvg[KP_G(j + 1, x4 + 1)];
            int8_t *rrow = &st->refg[KP_G(j + 1, x4 + 1)];
            if (w4 == 4) {
                KpU4 v4; v4.x = v4.y = v4.z = v4.w = mv;
                *(KpU4 *)row = v4;
                *(uint32_t *)rrow = 0x01010101u * (uint32_t)(uint8_t)ref;
            } else if (w4 == 2) {
                row[0] = mv; row[1] = mv;
                *(uint16_t *)rrow = (uint16_t)(0x0101u * (uint32_t)(uint8_t)ref);
            } else { row[0] = mv; rrow[0] = (int8_t)ref; }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ one macroblock (lane 0) */
#define KP_MB_OK        0
#define KP_MB_FAIL      1    /* syntax error: the context of the macroblock is cleared, nothing else is stored */
#define KP_MB_FAIL_KEEP 2    /* the macroblock had been decoded before (h264_slice.c: `if (c->decoded) return -1`): nothing is touched */

/* neighbours are staged (kp_stage_in, kp_stage_derive); returns KP_MB_* and leaves record / context / slots in the staging area */
KP_FN int kp_parse_mb(KpS &s)
{
    KpStage *st = s.st;
    h264b200_mb_t *r = &st->rec; KpMbCtx *c = &st->ctx;
    const h264b200_slice_t *sl = s.sl;
    const uint16_t sid = sl->slice_id;
    uint32_t mb_type;
    int cbp = 0, i16 = 0;
    if (st->old.decoded) return KP_MB_FAIL_KEEP;
    c->slice_id = sid; r->slice_id = sid;
    r->chroma_qp_off = sl->chroma_qp_off;
    r->dbk_off_a = sl->alpha_off; r->dbk_off_b = sl->beta_off;
    r->dbk_idc = sl->disable_deblocking_idc;
    s.n_slots = 0; s.ipcm_byte = -1;
    if (s.is_p && !s.prev_skipped) {
        s.skip_run = kp_ue(s);
        if (s.skip_run == 0xffffffffu || s.skip_run > s.N - s.addr) return KP_MB_FAIL;
        if (s.skip_run) s.prev_skipped = 1;
    }
    if (s.skip_run) { s.skip_run--; mb_type = 5; }
    else {
        mb_type = kp_ue(s);
        s.prev_skipped = 0;
        if (mb_type > (s.is_p ? 30u : 25u)) return KP_MB_FAIL;
        /* one numbering for both slice types: 0..4 P partitions, 5 = P_Skip, 6.. = intra (I4x4, 24 x I16x16, I_PCM) */
        if (!s.is_p) mb_type += 6;
        else if (mb_type >= 5) mb_type += 1;
    }
    if (mb_type <= 5) {
        s.n_inter++;
        if (kp_inter_pred(s, mb_type)) return KP_MB_FAIL;
        if (mb_type < 5) {
            const uint32_t v = kp_ue(s);
            if (v > 47) return KP_MB_FAIL;
            cbp = s.T->cbp_map[v][1];
        }
    } else {
        const uint32_t it = mb_type - 6;                          /* 0 I4x4, 1..24 I16x16, 25 I_PCM */
        c->ref_idx[0] = c->ref_idx[1] = c->ref_idx[2] = c->ref_idx[3] = -1;
        s.n_intra++;
        if (it == 25) {
            uint32_t byte_pos;
            r->avail = (uint8_t)((kp_intra_usable(s, 0) ? H264B200_AVAIL_A : 0) | (kp_intra_usable(s, 1) ? H264B200_AVAIL_B : 0) |
                                 (kp_intra_usable(s, 2) ? H264B200_AVAIL_C : 0) | (kp_intra_usable(s, 3) ? H264B200_AVAIL_D : 0));
            c->kind = r->mb_class = H264B200_MB_IPCM;
            while (kp_pos(s) & 7) if (kp_get1(s)) return KP_MB_FAIL;      /* pcm_alignment_zero_bit */
            byte_pos = kp_pos(s) >> 3;
            if (byte_pos + 384 > (s.rbsp_bits >> 3)) return KP_MB_FAIL;
            r->coef_offset = s.coef_used;
            s.ipcm_byte = (int32_t)byte_pos;                       /* the warp copies the samples when the macroblock is stored */
            s.n_slots = 12;
            kp_seek(s, (byte_pos + 384) * 8);                     /* reposition behind the samples */
            r->nz_mask = 0xffff;                                   /* TotalCoeff 16 everywhere: kp_stage_out */
            kp_set_qp_fields(s, r);
            r->qp_dbk = 0;                                         /* h264bsd_macroblock_layer.c:1003 */
            goto mb_done;
        }
        cbp = kp_intra_pred(s, it);
        if (cbp == -1) return KP_MB_FAIL;
        if (cbp == -2) {
            const uint32_t v = kp_ue(s);
            if (v > 47) return KP_MB_FAIL;
            cbp = s.T->cbp_map[v][0];
        } else i16 = 1;
    }
    if (cbp || i16) {
        if (kp_update_qp(s, kp_se(s))) return KP_MB_FAIL;
        kp_set_qp_fields(s, r);
        if (kp_parse_residual(s, cbp, i16)) return KP_MB_FAIL;
    } else kp_set_qp_fields(s, r);
mb_done:
    if (kp_overrun(s)) return KP_MB_FAIL;
    if (sl->disable_deblocking_idc != 1) {
        int fl = H264B200_DBK_INNER;
        if (s.mbx > 0 && (sl->disable_deblocking_idc != 2 || st->nctx[0].slice_id == sid)) fl |= H264B200_DBK_LEFT;
        if (s.mby > 0 && (sl->disable_deblocking_idc != 2 || st->nctx[1].slice_id == sid)) fl |= H264B200_DBK_TOP;
        r->dbk_flags = (uint8_t)fl;
        s.any_deblock = 1;
    }
    c->decoded = 1;
    return KP_MB_OK;
}

/* ------------------------------------------------------------------ warp-cooperative data movement */

/* stage the neighbours A, B, C, D of macroblock `addr` (records + contexts) and the old context of `addr` itself;
 * clear the record / context being built */
KP_FN void kp_stage_in(int lane, const KpPic &p, KpStage *st, uint32_t addr, int mbx, int mby, uint32_t W)
{
    for (int v = lane; v < 32; v += KP_LANES) {
        const int nb = v >> 3, part = v & 7;
        const int ok = nb == 0 ? mbx > 0 : nb == 1 ? mby > 0 : nb == 2 ? (mby > 0 && mbx + 1 < (int)W) : (mby > 0 && mbx > 0);
        const uint32_t na = nb == 0 ? addr - 1 : nb == 1 ? addr - W : nb == 2 ? addr - W + 1 : addr - W - 1;
        if (ok) ((KpU4 *)&st->nrec[nb])[part] = ((const KpU4 *)&p.mbs[na])[part];
        if (part < 2) {
            KpU4 z = {0, 0, 0, 0};
            if (ok) z = ((const KpU4 *)&p.ctx[na])[part];
            ((KpU4 *)&st->nctx[nb])[part] = z;
        }
        if (nb == 0 && part >= 2 && part < 4) ((KpU4 *)&st->old)[part - 2] = ((const KpU4 *)&p.ctx[addr])[part - 2];
        if (nb == 1 && part >= 2 && part < 4) { KpU4 z = {0, 0, 0, 0}; ((KpU4 *)&st->ctx)[part - 2] = z; }
        if (nb == 2) { KpU4 z = {0, 0, 0, 0}; ((KpU4 *)&st->rec)[part] = z; }
    }
}

/* Where the border cells of the per-macroblock grids come from is the same for every macroblock: computed once per
 * picture into KpStage.gsrc / msrc (0: an interior cell; else 1 + neighbour | index << 4 [| vector index << 8]). */
KP_FN void kp_stage_plan(int lane, KpStage *st)
{
    for (int v = lane; v < 64; v += KP_LANES) {
        uint32_t g = 0, m = 0;
        if (v < 40) {                                             /* luma: 5 rows of 8, block (x, y) at 9 + x + 8 y */
            const int rr = v >> 3, c = v & 7;
            if (rr == 0 && c >= 1 && c <= 4) { const int i = c - 1; g = 2u | (uint32_t)(10 + i + 2 * (i >> 1)) << 4; }       /* B: blocks 10 11 14 15 */
            else if (rr >= 1 && c == 0) { const int i = rr - 1; g = 1u | (uint32_t)(5 + 2 * i + 4 * (i >> 1)) << 4; }       /* A: blocks 5 7 13 15 */
        } else {                                                  /* chroma: per plane 3 rows of 4, block (x, y) at 5 + x + 4 y */
            const int w = v - 40, pl = w >= 12, i = w - 12 * pl, rr = i >> 2, c = i & 3;
            if (rr == 0 && (c == 1 || c == 2)) g = 2u | (uint32_t)(16 + 4 * pl + 1 + c) << 4;
            else if (rr >= 1 && c == 0) g = 1u | (uint32_t)(16 + 4 * pl + 2 * rr - 1) << 4;
        }
        if (v < 60) {                                             /* motion vector grid, KP_G */
            const int rr = v / 12, c = v - 12 * rr - 3;
            if (rr == 0) {
                if (c == 0) m = 4u | 3u << 4 | 15u << 8;                                           /* D: its bottom right block */
                else if (c >= 1 && c <= 4) m = 2u | (uint32_t)(2 + ((c - 1) >> 1)) << 4 | (uint32_t)(11 + c) << 8;   /* B: bottom row */
                else if (c == 5) m = 3u | 2u << 4 | 12u << 8;                                      /* C: bottom left block */
            } else if (c == 0) m = 1u | (uint32_t)(((rr - 1) >> 1) * 2 + 1) << 4 | (uint32_t)((rr - 1) * 4 + 3) << 8;   /* A: right column */
        }
        st->gsrc[v] = (uint16_t)g; st->msrc[v] = (uint16_t)m;
    }
}

/* after kp_stage_in (and a warp sync): which neighbours belong to the slice; the guard cells of the TotalCoeff grids
 * (nC) from macroblocks A and B and zeros inside; in P slices the border of the motion vector grid (KP_G) from the
 * neighbour records, "not available" inside.  Every lane computes the same availability word. */
KP_FN uint32_t kp_stage_derive(int lane, KpStage *st, uint16_t sid, int is_p, int mbx, int mby, uint32_t W)
{
    uint32_t av = 0;
    if (mbx > 0 && st->nctx[0].slice_id == sid) av |= 1;
    if (mby > 0) {
        if (st->nctx[1].slice_id == sid) av |= 2;
        if (mbx + 1 < (int)W && st->nctx[2].slice_id == sid) av |= 4;
        if (mbx > 0 && st->nctx[3].slice_id == sid) av |= 8;
    }
    for (int v = lane; v < 64; v += KP_LANES) {
        const uint32_t d = st->gsrc[v];
        int val = 0;
        if (d) { const int nb = (int)(d & 15) - 1; val = ((av >> nb) & 1) ? st->nctx[nb].tc[d >> 4] : 64; }
        st->grid[v] = (uint8_t)val;
    }
    if (is_p) {
        for (int v = lane; v < 60; v += KP_LANES) {
            const uint32_t d = st->msrc[v];
            int ref = -2;
            uint32_t mv = 0;
            if (d) {
                const int nb = (int)(d & 15) - 1;
                if ((av >> nb) & 1) {
                    ref = -1;
                    if (st->nctx[nb].kind == H264B200_MB_INTER) { ref = st->nctx[nb].ref_idx[(d >> 4) & 15]; mv = ((const uint32_t *)st->nrec[nb].mv)[d >> 8]; }
                }
            }
            st->refg[v] = (int8_t)ref; st->mvg[v] = mv;
        }
    }
    return av;
}

/* store the finished macroblock: the record (its vector half = the interior rows of the motion vector grid), the
 * context (TotalCoeff gathered from the nC grids), the coefficient slots (re-zeroing the staging area) or I_PCM samples */
KP_FN void kp_stage_out(int lane, const KpPic &p, KpStage *st, const KpTables *T, uint32_t addr, uint32_t coef_off, uint32_t n_slots,
                        int32_t ipcm_byte, const uint8_t *rbsp)
{
    const int inter = st->rec.mb_class == H264B200_MB_INTER;
    for (int v = lane; v < 16; v += KP_LANES) {
        if (v < 4) ((KpU4 *)&p.mbs[addr])[v] = ((const KpU4 *)&st->rec)[v];
        else if (v < 8) {
            KpU4 row = {0, 0, 0, 0};
            if (inter) row = *(const KpU4 *)&st->mvg[KP_G(v - 3, 1)];
            ((KpU4 *)&p.mbs[addr])[v] = row;
        } else {
            const int w = v - 8;
            uint32_t val;
            if (w >= 6) val = ((const uint32_t *)&st->ctx)[w];
            else if (ipcm_byte >= 0) val = 0x10101010u;
            else if (w < 4) {
                const uint8_t *ix = T->lc_idx + 4 * w;
                val = (uint32_t)st->grid[ix[0]] | ((uint32_t)st->grid[ix[1]] << 8) | ((uint32_t)st->grid[ix[2]] << 16) | ((uint32_t)st->grid[ix[3]] << 24);
            } else {
                const uint8_t *g = st->grid + 40 + 12 * (w - 4);
                val = (uint32_t)g[5] | ((uint32_t)g[6] << 8) | ((uint32_t)g[9] << 16) | ((uint32_t)g[10] << 24);
            }
            ((uint32_t *)&p.ctx[addr])[w] = val;
        }
    }
    if (ipcm_byte >= 0) {
        uint32_t *dst = (uint32_t *)(p.coef + (size_t)coef_off * 16);
        const uint8_t *src = rbsp + ipcm_byte;
        for (int i = lane; i < 96; i += KP_LANES)
            dst[i] = (uint32_t)src[4 * i] | ((uint32_t)src[4 * i + 1] << 8) | ((uint32_t)src[4 * i + 2] << 16) | ((uint32_t)src[4 * i + 3] << 24);
    } else {
        KpU4 *dst = (KpU4 *)(p.coef + (size_t)coef_off * 16);
        KpU4 *src = (KpU4 *)st->slots;
        const KpU4 z = {0, 0, 0, 0};
        for (uint32_t i = (uint32_t)lane; i < 2 * n_slots; i += KP_LANES) { dst[i] = src[i]; src[i] = z; }
    }
}
KP_FN void kp_stage_clear(int lane, KpStage *st)
{
    const KpU4 z = {0, 0, 0, 0};
    for (int i = lane; i < (int)(sizeof st->slots / 16); i += KP_LANES) ((KpU4 *)st->slots)[i] = z;
}

/* ------------------------------------------------------------------ failed slices and lost macroblocks (lane 0) */
/* h264_decoder.c mark_slice_corrupted (h264bsdMarkSliceCorrupted, h264bsd_slice_data.c:302-358) */
KP_NOINL void kp_mark_slice_corrupted(const KpPic &p, const h264b200_slice_t *sl, const uint8_t *map, uint32_t W, uint32_t N, uint32_t slice_last_mb)
{
    const uint16_t sid = sl->slice_id;
    uint32_t cur = sl->first_mb;
    if (cur >= N) return;
    if (slice_last_mb) {
        uint32_t i = slice_last_mb - 1, cnt = 0, lim = W > 10 ? W : 10;
        while (i > cur) {
            if (p.ctx[i].slice_id == sid && ++cnt >= lim) break;
            i--;
        }
        cur = i;
    }
    do {
        if (p.ctx[cur].slice_id != sid || !p.ctx[cur].decoded) break;
        p.ctx[cur].decoded = 0;
        if (map) {
            const uint8_t grp = map[cur];
            do cur++; while (cur < N && map[cur] != grp);
            if (cur >= N) cur = 0;
        } else cur = cur + 1 < N ? cur + 1 : 0;
    } while (cur);
}

KP_FN void kp_zero_rec(h264b200_mb_t *r) { KpU4 z = {0, 0, 0, 0}; for (int i = 0; i < 8; i++) ((KpU4 *)r)[i] = z; }

/* h264_decoder.c conceal_picture: records for the macroblocks no slice delivered.  Returns their number. */
KP_NOINL uint32_t kp_conceal_picture(const KpPic &p, const KpTables *T, const h264b200_pichdr_t *hdr, uint32_t num_decoded, KpResult &res)
{
    const uint32_t W = hdr->width_mbs, H = hdr->height_mbs, N = W * H;
    const int whole = num_decoded == 0;
    const int ref = hdr->conceal_as_p ? hdr->conceal_ref_slot : -1;
    uint32_t i, n = 0;
    res.n_conceal = 0;
    if (ref >= 0 || whole) {
        for (i = 0; i < N; i++) {
            h264b200_mb_t *r = &p.mbs[i];
            if (p.ctx[i].decoded) continue;
            kp_zero_rec(r);
            n++;
            if (ref >= 0) {
                r->mb_class = H264B200_MB_INTER; r->part_flags = 31;
                for (int q = 0; q < 4; q++) r->ref_slot[q] = (uint8_t)ref;
                r->qp_y = r->qp_dbk = 40; r->qp_c = T->qpc[40];
                r->flags = H264B200_MBF_DBK_AS_INTRA;
                if (!whole) {
                    r->dbk_flags = (uint8_t)(H264B200_DBK_INNER | (i % W ? H264B200_DBK_LEFT : 0) | (i >= W ? H264B200_DBK_TOP : 0));
                    res.any_deblock = 1;
                }
                res.n_inter++;
            } else {
                uint32_t *dst;
                if (res.coef_used + 12 > p.coef_cap) { r->mb_class = H264B200_MB_MISSING; continue; }
                r->mb_class = H264B200_MB_IPCM; r->coef_offset = res.coef_used; r->nz_mask = 0xffff;
                dst = (uint32_t *)(p.coef + (size_t)res.coef_used * 16);
                for (int k = 0; k < 96; k++) dst[k] = 0x80808080u;
                res.coef_used += 12; res.n_intra++;
            }
        }
        if (whole) for (i = 0; i < N; i++) p.mbs[i].dbk_flags = 0;
        return n;
    }
    {
        uint32_t lost = N - num_decoded, need = (lost * 4 + 31) / 32, *list, row, col, j;
        if (res.coef_used + need > p.coef_cap) {
            for (i = 0; i < N; i++) if (!p.ctx[i].decoded) { kp_zero_rec(&p.mbs[i]); p.mbs[i].mb_class = H264B200_MB_MISSING; }
            return lost;
        }
        list = (uint32_t *)(p.coef + (size_t)res.coef_used * 16);
        res.conceal_offset = res.coef_used;
        for (i = 0; i < N && !p.ctx[i].decoded; i++) ;
        row = i / W; col = i % W;
#define KP_CONCEAL_ONE(addr_) do { \
            const uint32_t a_ = (addr_), y_ = a_ / W, x_ = a_ % W; \
            h264b200_mb_t *r_ = &p.mbs[a_]; \
            kp_zero_rec(r_); \
            r_->mb_class = H264B200_MB_CONCEAL; \
            r_->avail = (uint8_t)((y_ > 0 && p.ctx[a_ - W].decoded ? H264B200_CN_ABOVE : 0) | (y_ + 1 < H && p.ctx[a_ + W].decoded ? H264B200_CN_BELOW : 0) | \
                                  (x_ > 0 && p.ctx[a_ - 1].decoded ? H264B200_CN_LEFT : 0) | (x_ + 1 < W && p.ctx[a_ + 1].decoded ? H264B200_CN_RIGHT : 0)); \
            r_->qp_y = r_->qp_dbk = 40; r_->qp_c = T->qpc[40]; \
            r_->dbk_flags = (uint8_t)(H264B200_DBK_INNER | (x_ ? H264B200_DBK_LEFT : 0) | (y_ ? H264B200_DBK_TOP : 0)); \
            p.ctx[a_].decoded = 1; list[n++] = a_; } while (0)
        for (j = col; j-- > 0;) KP_CONCEAL_ONE(row * W + j);
        for (j = col + 1; j < W; j++) if (!p.ctx[row * W + j].decoded) KP_CONCEAL_ONE(row * W + j);
        if (row) for (j = 0; j < W; j++) for (i = row; i-- > 0;) KP_CONCEAL_ONE(i * W + j);
        for (i = row + 1; i < H; i++) for (j = 0; j < W; j++) if (!p.ctx[i * W + j].decoded) KP_CONCEAL_ONE(i * W + j);
#undef KP_CONCEAL_ONE
        res.n_conceal = n;
        res.coef_used += (n * 4 + 31) / 32;
        res.any_deblock = 1;
    }
    return n;
}

/* ------------------------------------------------------------------ one picture (the whole warp) */
KP_FN void kp_parse_picture(int lane, const KpPic &p, KpStage *st, const KpTables *T)
{
    const h264b200_pichdr_t *hdr = (const h264b200_pichdr_t *)p.block;
    const uint32_t W = hdr->width_mbs, H = hdr->height_mbs, N = W * H, n_slices = hdr->n_slices;
    uint32_t num_decoded = 0, off = sizeof(h264b200_pichdr_t), flags = 0;
    KpS s;
    {   /* contexts: nothing decoded yet */
        const KpU4 z = {0, 0, 0, 0};
        for (uint32_t i = (uint32_t)lane; i < 2 * N; i += KP_LANES) ((KpU4 *)p.ctx)[i] = z;
    }
    kp_stage_clear(lane, st);
    kp_stage_plan(lane, st);
    s.T = T; s.st = st; s.W = W; s.N = N;
    s.coef_used = 0; s.n_intra = s.n_inter = s.any_deblock = 0;
    KP_SYNC();

    for (uint32_t k = 0; k < n_slices; k++) {
        const h264b200_slice_t *sl = (const h264b200_slice_t *)(p.block + off);
        const uint8_t *rbsp = (const uint8_t *)(sl + 1);
        const uint8_t *map = sl->map_off ? (const uint8_t *)sl + sl->map_off : NULL;
        uint32_t addr = sl->first_mb, next_linear = 0xffffffffu, mb_count = 0, slice_last_mb = 0;
        int more = 1, failed = 0;
        off += sl->size;
        if (num_decoded == N) continue;                           /* the picture is complete: later slices of the access unit are ignored */
        s.sl = sl; s.is_p = sl->is_p; s.qp = sl->slice_qp;
        s.skip_run = 0; s.prev_skipped = 0;
        s.mbx = s.mby = 0;
        if (lane == 0) kp_bits_init(s, rbsp, sl->rbsp_len, sl->bit_off, sl->payload_bits);
        if (addr >= N) { failed = 1; more = 0; }
        while (more) {
            int rc = 0; uint32_t n_slots = 0, coef_off = 0; int32_t ipcm = -1;
            if (addr != next_linear) { s.mbx = (int)(addr % W); s.mby = (int)(addr / W); }
            else if (++s.mbx == (int)W) { s.mbx = 0; s.mby++; }
            next_linear = addr + 1;
            s.addr = addr;
            kp_stage_in(lane, p, st, addr, s.mbx, s.mby, W);
            KP_SYNC();
            s.avail = kp_stage_derive(lane, st, sl->slice_id, s.is_p, s.mbx, s.mby, W);
            KP_SYNC();
            if (lane == 0) {
                s.T = kp_pin_shared(T); s.st = kp_pin_shared(st);
                rc = kp_parse_mb(s);
                n_slots = s.n_slots; coef_off = s.coef_used; ipcm = s.ipcm_byte;
                if (rc == KP_MB_OK) {
                    s.coef_used += n_slots;
                    more = kp_more_data(s) || s.skip_run;
                } else if (rc == KP_MB_FAIL) {
                    /* h264_slice.c: memset(c, 0) ... c->slice_id = 0 */
                    KpU4 z = {0, 0, 0, 0};
                    ((KpU4 *)&st->ctx)[0] = z; ((KpU4 *)&st->ctx)[1] = z;
                }
            }
            KP_SYNC();
            rc = KP_BCAST(rc); more = KP_BCAST(more);
            n_slots = KP_BCAST(n_slots); coef_off = KP_BCAST(coef_off); ipcm = KP_BCAST(ipcm);
            if (rc == KP_MB_OK) kp_stage_out(lane, p, st, T, addr, coef_off, n_slots, ipcm, rbsp);
            else {
                if (rc == KP_MB_FAIL) for (int v = lane; v < 2; v += KP_LANES) ((KpU4 *)&p.ctx[addr])[v] = ((const KpU4 *)&st->ctx)[v];
                kp_stage_clear(lane, st);
                failed = 1; more = 0;
                KP_SYNC();
                break;
            }
            KP_SYNC();
            mb_count++;
            if (!s.is_p) slice_last_mb = addr;
            if (map) { const uint8_t grp = map[addr]; do addr++; while (addr < N && map[addr] != grp); }
            else addr++;
            if (more && addr >= N) { failed = 1; more = 0; }
        }
        if (!failed && num_decoded + mb_count > N) failed = 1;
        if (!failed) num_decoded += mb_count;
        else {
            flags |= H264B200_PS_SLICE_ERROR;
            if (lane == 0) kp_mark_slice_corrupted(p, sl, map, W, N, slice_last_mb);
            KP_SYNC();
        }
    }

    if (lane == 0) {
        KpResult res;
        res.coef_used = s.coef_used; res.n_intra = s.n_intra; res.n_inter = s.n_inter; res.any_deblock = s.any_deblock;
        res.n_conceal = 0; res.conceal_offset = 0; res.pad[0] = res.pad[1] = 0;
        res.stat.err_mbs = 0; res.stat.flags = flags; res.stat.decoded_mbs = num_decoded;
        if (num_decoded != N) {
            res.stat.flags |= H264B200_PS_INCOMPLETE;
            if (hdr->tentative) res.stat.flags |= H264B200_PS_DROPPED;
            res.stat.err_mbs = kp_conceal_picture(p, T, hdr, num_decoded, res);
        }
        res.stat.coef_slots = res.coef_used;
        *p.res = res;
    }
    KP_SYNC();
}

#endif
