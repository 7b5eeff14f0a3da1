/* k2_math.cuh — the luma interpolation arithmetic of K2 (k2_inter.cuh) for one 4x4 block, on registers only.
 *
 * Same results as 8.4.2.2.1 and the reference's nine interpolators (h264bsd_reconstruct.c:491-1791):
 * the 9x9 window arrives as three byte-aligned words per row (bytes 0..8), and the sixteen fractional
 * positions are ONE formula, out = (S * (3 - n) + 1) >> 1 with S the sum of the n in {1, 2} operands the
 * position uses among {integer sample G', horizontal half b', vertical half h', centre j}; the primed
 * operands sit one row lower / one column right when the fraction is 3/4.
 *   - horizontal 6-tap sums (1,-5,20,20,-5,1) of a row: two dp4a on packed bytes per sum;
 *   - vertical 6-tap sums of raw samples: two samples per instruction on biased 16-bit lanes,
 *     (a+f) + 20(c+d) + 2560 - 5(b+e) stays within [0, 65535] per lane;
 *   - centre sample j: the vertical filter over the UNCLIPPED horizontal sums, (x + 512) >> 10.
 * any_b / any_h / any_j say which operand families have to be computed at all (in the kernel: warp votes).
 *
 * The file also compiles as plain C++ (no CUDA): tests/test_k2_math_cpu.py builds it with g++ and checks
 * every fractional position against a scalar statement of the standard's formulas on random and extreme
 * windows — the packed-arithmetic tricks are verified before they reach a GPU.
 */
#pragma once
#include <stdint.h>

#ifdef __CUDA_ARCH__
#define K2M_FN __device__ __forceinline__
K2M_FN uint32_t k2m_funnel_r(uint32_t lo, uint32_t hi, int sh) { return __funnelshift_r(lo, hi, sh); }
/* four unsigned bytes of a times four signed bytes of b, accumulated (dp4a.u32.s32) */
K2M_FN int k2m_dp4a_us(uint32_t a, int b, int c) { int d; asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
K2M_FN int k2m_clip255(int v) { return min(max(v, 0), 255); }
#else
#define K2M_FN static inline
K2M_FN uint32_t k2m_funnel_r(uint32_t lo, uint32_t hi, int sh) { return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31)); }
K2M_FN int k2m_dp4a_us(uint32_t a, int b, int c)
{
    for (int i = 0; i < 4; i++) c += (int)((a >> (8 * i)) & 0xff) * (int)(int8_t)((uint32_t)b >> (8 * i));
    return c;
}
K2M_FN int k2m_clip255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
#endif

K2M_FN int k2m_tap6(int a, int b, int c, int d, int e, int f) { return (a + f) - 5 * (b + e) + 20 * (c + d); }

/* r0/r1/r2: window rows 0..8 as bytes 0..3 / 4..7 / 8 of the row.  fx, fy: quarter-sample fractions.
 * use_*: the operands THIS position takes (all false: the block is not an inter block, out is don't-care);
 * n_ops = how many of them (1 or 2).  out[py]: the four predicted samples of row py, low byte first. */
K2M_FN void k2m_luma4x4(const uint32_t (&r0)[9], const uint32_t (&r1)[9], const uint32_t (&r2)[9], int fx, int fy,
                        bool use_g, bool use_b, bool use_h, bool use_j, int n_ops, bool any_b, bool any_h, bool any_j,
                        uint32_t (&out)[4])
{
    const int dn = fy == 3, rt = fx == 3;
    /* ---- the 4 samples at columns x+rt .. x+rt+3 of every window row (G' and the inputs of h') ---- */
    uint32_t cw[9];
#pragma unroll
    for (int r = 0; r < 9; r++) cw[r] = k2m_funnel_r(r0[r], r1[r], 8 * (2 + rt));

    /* ---- horizontal 6-tap sums: hs[r][k] for output column k of window row r ---- */
    int hs[9][4];
    if (any_b || any_j) {
        const int T0 = 0x1414fb01, T1 = 0x000001fb;      /* (1,-5,20,20) and (-5,1,0,0) as signed bytes, low byte first */
#pragma unroll
        for (int r = 0; r < 9; r++) {
            if (!any_j && (r < 2 || r > 6)) { hs[r][0] = hs[r][1] = hs[r][2] = hs[r][3] = 0; continue; }
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const uint32_t lo = k ? k2m_funnel_r(r0[r], r1[r], 8 * k) : r0[r];
                const uint32_t hi = k ? k2m_funnel_r(r1[r], r2[r], 8 * k) : r1[r];
                hs[r][k] = k2m_dp4a_us(lo, T0, k2m_dp4a_us(hi, T1, 0));
            }
        }
    }
#ifndef __CUDA_ARCH__
    else { for (int r = 0; r < 9; r++) for (int k = 0; k < 4; k++) hs[r][k] = 0; }     /* the host compiler cannot see that it is unused */
#endif

    /* ---- per output row: operands and the final blend ---- */
#pragma unroll
    for (int py = 0; py < 4; py++) {
        int bq[4] = {0, 0, 0, 0}, hq[4] = {0, 0, 0, 0}, jq[4] = {0, 0, 0, 0};
        if (any_b) {
#pragma unroll
            for (int k = 0; k < 4; k++) bq[k] = k2m_clip255(((dn ? hs[py + 3][k] : hs[py + 2][k]) + 16) >> 5);
        }
        if (any_h) {
#pragma unroll
            for (int half = 0; half < 2; half++) {
                uint32_t e[6];
#pragma unroll
                for (int t = 0; t < 6; t++) e[t] = (cw[py + t] >> (8 * half)) & 0x00ff00ffu;
                const uint32_t s = (e[0] + e[5] + 0x0a000a00u) + 20u * (e[2] + e[3]) - 5u * (e[1] + e[4]);
                hq[half] = k2m_clip255(((int)(s & 0xffff) - 2560 + 16) >> 5);
                hq[half + 2] = k2m_clip255(((int)(s >> 16) - 2560 + 16) >> 5);
            }
        }
        if (any_j) {
#pragma unroll
            for (int k = 0; k < 4; k++)
                jq[k] = k2m_clip255((k2m_tap6(hs[py][k], hs[py + 1][k], hs[py + 2][k], hs[py + 3][k], hs[py + 4][k], hs[py + 5][k]) + 512) >> 10);
        }
        const uint32_t gw = dn ? cw[py + 3] : cw[py + 2];
        uint32_t pk = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            int s = 0;
            if (use_g) s += (gw >> (8 * k)) & 0xff;
            if (use_b) s += bq[k];
            if (use_h) s += hq[k];
            if (use_j) s += jq[k];
            const int v = (s * (3 - n_ops) + 1) >> 1;
            pk |= (uint32_t)v << (8 * k);
        }
        out[py] = pk;
    }
}
