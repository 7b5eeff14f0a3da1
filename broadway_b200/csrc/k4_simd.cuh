/* k4_simd.cuh — the H.264 edge filter on TWO lines at once, one line per 16-bit half of
 * a 32-bit register (sm_100a: VABSDIFF4, VIMNMX.S16x2, PRMT; everything else is plain
 * 32-bit integer arithmetic arranged so that no carry or borrow crosses the halves).
 *
 * Same arithmetic as 8.7.2.3 / 8.7.2.4 and the reference's FilterVerLumaEdge /
 * FilterHorLuma / FilterVerChromaEdge / FilterHorChroma (h264bsd_deblocking.c:649-1121);
 * both lines of a register belong to the same 4-sample (luma) / 2-sample (chroma) edge
 * segment, so boundary strength and thresholds are scalars per lane.
 *
 * The file also compiles as plain C++ (no CUDA): tests/test_k4_simd_cpu.py builds it with
 * g++ and checks dbk_edge2 against a scalar statement of the filter on random and
 * boundary inputs — the packed-arithmetic tricks are verified before they reach a GPU.
 */
#pragma once
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define K4S_FN __device__ __forceinline__
K4S_FN uint32_t k4s_absd4(uint32_t a, uint32_t b) { return __vabsdiffu4(a, b); }
K4S_FN uint32_t k4s_min2(uint32_t a, uint32_t b) { uint32_t r; asm("min.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
K4S_FN uint32_t k4s_max2(uint32_t a, uint32_t b) { uint32_t r; asm("max.s16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
/* prmt.b32 itself, not __byte_perm: the intrinsic keeps only 3 bits per selector nibble and would drop the sign-replication bit */
K4S_FN uint32_t k4s_perm(uint32_t a, uint32_t b, uint32_t s) { uint32_t r; asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s)); return r; }
#else
#define K4S_FN static inline
K4S_FN uint32_t k4s_absd4(uint32_t a, uint32_t b)
{
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) { int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i); }
    return r;
}
K4S_FN uint32_t k4s_min2(uint32_t a, uint32_t b)
{
    int16_t al = (int16_t)a, ah = (int16_t)(a >> 16), bl = (int16_t)b, bh = (int16_t)(b >> 16);
    return (uint32_t)(uint16_t)(al < bl ? al : bl) | ((uint32_t)(uint16_t)(ah < bh ? ah : bh) << 16);
}
K4S_FN uint32_t k4s_max2(uint32_t a, uint32_t b)
{
    int16_t al = (int16_t)a, ah = (int16_t)(a >> 16), bl = (int16_t)b, bh = (int16_t)(b >> 16);
    return (uint32_t)(uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
}
K4S_FN uint32_t k4s_perm(uint32_t a, uint32_t b, uint32_t s)
{
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t sel = (s >> (4 * i)) & 0xf, byte = (uint32_t)(src >> (8 * (sel & 7))) & 0xff;
        if (sel & 8) byte = (byte & 0x80) ? 0xff : 0x00;      /* sign replication mode */
        r |= byte << (8 * i);
    }
    return r;
}
#endif

#define K4S_K  0x00010001u      /* 1 in both halves */
#define K4S_H  0x80008000u
#define K4S_LO 0x00ff00ffu
#define K4S_9  0x01ff01ffu

/* 0xffff in every half whose bit 15 is set */
K4S_FN uint32_t k4s_signmask(uint32_t t) { return k4s_perm(t, 0, 0xbb99); }
K4S_FN uint32_t k4s_sel(uint32_t m, uint32_t a, uint32_t b) { return (a & m) | (b & ~m); }   /* one LOP3 */

/* One edge of two lines.  v[0..3] = p3..p0, v[4..7] = q0..q3, each register = (line A | line B << 16),
 * samples 0..255.  bs 0..4 (same for both lines); thr = alpha | beta << 8; tcw = tc0(bS=1) | tc0(bS=2) << 8 |
 * tc0(bS=3) << 16.  Chroma lines (luma == false) only ever change p0 and q0.  any_weak / any_strong are
 * warp-uniform hints: the variant no lane of the warp needs is skipped.
 *
 * Packed tricks (all verified on the CPU by the test above):
 *   x < t   <=>  bit 15 of (x + 0x8000 - t) clear            (x, t in 0..255; one add with a per-edge constant)
 *   signed intermediates carry a bias (+256 after the shift) so that plain 32-bit add/sub/shift work per half;
 *   clip to 0..255 happens in the biased domain [256, 511], whose low byte is the clipped value. */
K4S_FN void dbk_edge2(uint32_t *v, int bs, uint32_t thr, uint32_t tcw, bool luma, bool any_weak, bool any_strong)
{
    const uint32_t p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    const uint32_t alpha = bs ? (thr & 0xff) : 0u, beta = (thr >> 8) & 0xff;     /* alpha 0: nothing passes, like bS 0 */
    const uint32_t HA = K4S_H - alpha * K4S_K, HB = K4S_H - beta * K4S_K;
    const uint32_t ad = k4s_absd4(p0, q0);
    const uint32_t off = k4s_signmask((ad + HA) | (k4s_absd4(p1, p0) + HB) | (k4s_absd4(q1, q0) + HB));   /* 0xffff: line not filtered */
    const uint32_t nap = luma ? k4s_signmask(k4s_absd4(p2, p0) + HB) : 0xffffffffu;                     /* 0xffff: ap false */
    const uint32_t naq = luma ? k4s_signmask(k4s_absd4(q2, q0) + HB) : 0xffffffffu;
    uint32_t n0 = p0, n1 = p1, n2 = p2, m0 = q0, m1 = q1, m2 = q2;
    const uint32_t C256 = 256u * K4S_K;
    if (any_weak) {
        const uint32_t offw = (bs < 4) ? off : 0xffffffffu;
        const uint32_t tc0 = (tcw >> ((8 * (bs - 1)) & 31)) & 0xff;
        const uint32_t TC0 = tc0 * K4S_K;
        const uint32_t tc = TC0 + (luma ? ((~nap & K4S_K) + (~naq & K4S_K)) : K4S_K);
        /* delta + 256 = (4(q0-p0) + (p1-q1) + 4 + 2048) >> 3, clipped to 256 -/+ tc */
        uint32_t D = ((q0 << 2) + p1 + 2052u * K4S_K) - ((p0 << 2) + q1);
        D = (D >> 3) & K4S_9;
        D = k4s_min2(k4s_max2(D, C256 - tc), C256 + tc);
        const uint32_t tp = k4s_min2(k4s_max2(p0 + D, C256), K4S_9);                 /* p0 + delta, clipped, + 256 */
        const uint32_t tq = k4s_min2(k4s_max2(q0 + 2u * C256 - D, C256), K4S_9);     /* q0 - delta, clipped, + 256 */
        const uint32_t sel = ~offw & K4S_LO;
        n0 = k4s_sel(sel, tp, p0); m0 = k4s_sel(sel, tq, q0);
        const uint32_t avg = ((p0 + q0 + K4S_K) >> 1) & K4S_LO;
        const uint32_t lo0 = C256 - TC0, hi0 = C256 + TC0;
        uint32_t E = (p2 + avg + 2u * C256) - (p1 << 1), F = (q2 + avg + 2u * C256) - (q1 << 1);
        E = k4s_min2(k4s_max2((E >> 1) & K4S_9, lo0), hi0);
        F = k4s_min2(k4s_max2((F >> 1) & K4S_9, lo0), hi0);
        n1 = k4s_sel(~(offw | nap) & K4S_LO, p1 + E, p1);
        m1 = k4s_sel(~(offw | naq) & K4S_LO, q1 + F, q1);
    }
    if (any_strong) {
        const uint32_t offs = (bs == 4) ? off : 0xffffffffu;
        const uint32_t nsmall = k4s_signmask(ad + (K4S_H - ((alpha >> 2) + 2u) * K4S_K));
        const uint32_t sp = ~(offs | nap | nsmall), sq = ~(offs | naq | nsmall), st = ~offs;
        const uint32_t s = p0 + q0, up = p1 + s, uq = q1 + s;
        const uint32_t sp0 = ((p2 + q1 + 4u * K4S_K + (up << 1)) >> 3) & K4S_LO, sp1 = ((p2 + up + 2u * K4S_K) >> 2) & K4S_LO;
        const uint32_t sp2 = ((((p3 + p2) << 1) + p2 + up + 4u * K4S_K) >> 3) & K4S_LO;
        const uint32_t sq0 = ((q2 + p1 + 4u * K4S_K + (uq << 1)) >> 3) & K4S_LO, sq1 = ((q2 + uq + 2u * K4S_K) >> 2) & K4S_LO;
        const uint32_t sq2 = ((((q3 + q2) << 1) + q2 + uq + 4u * K4S_K) >> 3) & K4S_LO;
        const uint32_t wp0 = (((p1 << 1) + p0 + q1 + 2u * K4S_K) >> 2) & K4S_LO, wq0 = (((q1 << 1) + q0 + p1 + 2u * K4S_K) >> 2) & K4S_LO;
        n0 = k4s_sel(sp, sp0, k4s_sel(st, wp0, n0)); n1 = k4s_sel(sp, sp1, n1); n2 = k4s_sel(sp, sp2, n2);
        m0 = k4s_sel(sq, sq0, k4s_sel(st, wq0, m0)); m1 = k4s_sel(sq, sq1, m1); m2 = k4s_sel(sq, sq2, m2);
    }
    v[1] = n2; v[2] = n1; v[3] = n0; v[4] = m0; v[5] = m1; v[6] = m2;
}

/* bytes <-> two-line registers.
 * k4s_unpack_rows: words a, b = four consecutive samples of line A and of line B  ->  out[j] = a_j | b_j << 16
 * k4s_pack_rows:   the inverse */
K4S_FN void k4s_unpack_rows(uint32_t a, uint32_t b, uint32_t *out)
{
    const uint32_t e02 = k4s_perm(a, b, 0x6420), e13 = k4s_perm(a, b, 0x7531);     /* [a0 a2 b0 b2], [a1 a3 b1 b3] */
    out[0] = e02 & K4S_LO; out[2] = (e02 >> 8) & K4S_LO;
    out[1] = e13 & K4S_LO; out[3] = (e13 >> 8) & K4S_LO;
}
K4S_FN void k4s_pack_rows(const uint32_t *in, uint32_t *a, uint32_t *b)
{
    const uint32_t f01 = k4s_perm(in[0], in[1], 0x6240), f23 = k4s_perm(in[2], in[3], 0x6240);   /* [a0 a1 b0 b1], [a2 a3 b2 b3] */
    *a = k4s_perm(f01, f23, 0x5410); *b = k4s_perm(f01, f23, 0x7632);
}
/* two adjacent samples of one row (a 16-bit load) = the same position of two adjacent column lines */
K4S_FN uint32_t k4s_unpack_pair(uint32_t h) { return k4s_perm(h, 0, 0x4140); }
K4S_FN uint32_t k4s_pack_pair(uint32_t v) { return k4s_perm(v, 0, 0x4420); }
