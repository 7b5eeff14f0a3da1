/* k2_inter.cuh — kernel family 2: inter prediction + residual add, frame-parallel
 * over every P macroblock of every picture in the batch.
 *
 * Device replacement of h264bsdInterPrediction's partition walk
 * (h264bsd_inter_prediction.c:364-487), h264bsdPredictSamples and the nine luma
 * interpolators (h264bsd_reconstruct.c:1819-1941, :491-1791), PredictChroma
 * (:110-476), the edge-clamped block fetch h264bsdFillBlock (:2222-2314) and
 * h264bsdWriteOutputBlocks (h264bsd_image.c:171-343).
 *
 * ONE THREAD PER 4x4 LUMA BLOCK, everything in registers, no shared memory.
 * Prediction is per-sample independent, so every partition shape is sixteen
 * 4x4 blocks with their own vector (the record stores final vectors per 4x4
 * block).  A half-warp is one macroblock (lane & 15 = raster block index), a
 * warp two horizontally adjacent macroblocks, so each luma row store of a warp
 * is one full 32-byte sector.  Per thread:
 *   - the 9x9 reference window is fetched as 9 rows x 3 aligned 32-bit words
 *     through the read-only path (windows of neighbouring blocks overlap in L1)
 *     and byte-aligned with funnel shifts; windows that leave the picture take a
 *     per-sample clamped path (= the reference's coordinate clamp);
 *   - horizontal 6-tap sums (1,-5,20,20,-5,1) of a row are eight dp4a on packed
 *     bytes; vertical 6-tap sums of raw samples run two samples per instruction
 *     on biased 16-bit lanes; the centre sample j is the vertical filter over the
 *     unclipped horizontal sums with (x+512)>>10 (8.4.2.2.1);
 *   - the 16 fractional positions are ONE formula: out = (S*(3-n)+1)>>1 with S
 *     the sum of the n in {1,2} operands the position uses among {integer
 *     sample G', horizontal half b', vertical half h', centre j}, where the
 *     primed operands sit one row lower / one column right when the fraction is
 *     3/4.  Which operand families a warp needs at all is decided by warp votes,
 *     so streams with coherent motion skip most of the arithmetic;
 *   - chroma is the 1/8-pel bilinear blend of a 3x3 window (8.4.2.2.2), two
 *     2x2 blocks (Cb, Cr) per thread;
 *   - the residual K1 left in the block's coefficient slot is added with
 *     clipping and the block is stored: 4 x 32-bit luma, 4 x 16-bit chroma.
 * HBM per inter macroblock: 384 B reference (unique) + 384 B written + 128 B
 * record + 32 B per coded block.
 */
#pragma once
#include "k_common.cuh"
#include "k2_math.cuh"

#define K2_THREADS 128

/* clamped single-sample fetch: the reference's out-of-picture behaviour (h264bsdFillBlock) */
__device__ __forceinline__ uint32_t ref_px(const uint8_t *pl, int w, int h, int x, int y)
{
    x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
    return __ldg(pl + (size_t)y * w + x);
}

/* 8 CTAs per SM at 64 registers (a few spilled words): measured 2.19 ms per 256 pictures against 2.24 ms at 80 and
 * 2.33 ms at 96 registers without spills — the loads of the reference window are what the extra warps hide.
 * (An L2 prefetch of the records of later CTAs made no difference.)
 *
 * Tried and dropped (round 1, all bit-exact, default bench, ms per 256 pictures against 2.19 for this kernel; ncu of
 * this kernel: 1 540 warp instructions per thread, issue slots 64 % busy, DRAM traffic = algorithmic bytes):
 *   - sorting the 4x4 blocks of a CTA by fractional-position class (G | b | b+h | h | h+j | j | j+b) so that the
 *     warp votes above switch most of the arithmetic off: 2.69 (256-block tiles, one run per warp) and 3.15
 *     (512-block tiles, chunks of 32 claimed dynamically so the warps finish together).  A sorted warp gathers
 *     blocks from up to 32 macroblocks, so its window loads no longer share sectors; the two neighbouring
 *     macroblocks of an unsorted warp overlap heavily and that locality is worth more than the instructions;
 *   - the three window words as two 8-byte loads plus selects: 2.31;
 *   - prefetch.global.L1 of the residual slots at the top (ptxas sinks the loads to their use): 2.21, no change.
 * Round 2, what-if bounds (profiles/r02_k2_whatif_ab.json; build flags below, wrong pictures, timing only): with the luma
 * window for free (K2_WHATIF_NO_WINDOW_LOADS) the kernel takes 1.73 ms, with the interpolation replaced by a copy
 * (K2_WHATIF_NO_MATH) 1.33 ms.  So staging the window through shared memory / TMA can win 21 % at the very most, and
 * the larger lever is the 1.33 ms this thread-per-4x4 mapping costs as a pure data mover (per-thread record decoding,
 * 33 narrow loads, 4-byte stores): a warp-per-macroblock-pair mapping with 16-byte row accesses is the next step. */
__global__ void __launch_bounds__(K2_THREADS, 8) k2_inter(Batch b)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, blk = lane & 15, bx = blk & 3, by = blk >> 2;
    const uint32_t g0 = (blockIdx.x * K2_THREADS + threadIdx.x) >> 4;        /* batch-wide macroblock index */
    const bool in_range = g0 < b.total_mbs;
    const uint32_t g = in_range ? g0 : b.total_mbs - 1;
    const PicJob &job = b.jobs[find_job(b, g)];
    const uint32_t mbi = g - job.mb_base;
    const h264b200_mb_t *mb = job.mbs + mbi;
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(mb);
    const bool inter = in_range && (__ldg(rec) & 0xff) == H264B200_MB_INTER;
    if (!__any_sync(FULL, inter)) return;

    const int W = job.wm * 16, H = job.hm * 16, CW = W >> 1, CH = H >> 1;
    const int mbx = mbi % job.wm, mby = mbi / job.wm;
    const size_t ysize = (size_t)W * H, csize = ysize >> 2;
    const uint32_t mvw = __ldg(rec + 16 + blk);                              /* mv[blk] = {hor, ver} */
    const int mvx = (int)(short)(mvw & 0xffff), mvy = (int)mvw >> 16;
    const uint32_t slots = __ldg(rec + 6);                                   /* ref_slot[4] */
    const uint8_t *ref = job.frames + (size_t)((slots >> (8 * ((by >> 1) * 2 + (bx >> 1)))) & 0xff) * job.frame_bytes;
    const uint32_t mask = __ldg(rec + 4);
    const int16_t *coef = job.coef + (size_t)__ldg(rec + 3) * 16;
    const int fx = mvx & 3, fy = mvy & 3;

    /* residual of this block: requested now, consumed after the interpolation */
    const int bi = (blk & 1) | ((blk & 2) << 1) | ((blk & 4) >> 1) | (blk & 8);                /* raster -> luma4x4BlkIdx */
    const bool has_res = inter && ((mask >> bi) & 1);
    int4 res_lo = make_int4(0, 0, 0, 0), res_hi = make_int4(0, 0, 0, 0);
    if (has_res) {
        const int4 *rs = reinterpret_cast<const int4 *>(coef + slot_index(mask, bi) * 16);
        res_lo = __ldg(rs); res_hi = __ldg(rs + 1);
    }

    /* ================================ luma ================================ */
    uint32_t out_rows[4];
    {
        const int x0 = mbx * 16 + bx * 4 + (mvx >> 2) - 2, y0 = mby * 16 + by * 4 + (mvy >> 2) - 2;   /* window origin */
        /* which operand families does this position use (see header) */
        const bool useJ = inter && ((fx == 2 && fy != 0) || (fy == 2 && fx != 0));
        const bool useB = inter && fx != 0 && fy != 2;
        const bool useH = inter && fy != 0 && fx != 2;
        const bool useG = inter && !useJ && !(useB && useH) && !(fx == 2) && !(fy == 2);
        const int n_ops = (int)useJ + (int)useB + (int)useH + (int)useG;      /* 1 or 2 */
        const bool anyJ = __any_sync(FULL, useJ), anyB = __any_sync(FULL, useB), anyH = __any_sync(FULL, useH);

        /* ---- window rows as three byte-aligned words: bytes 0..8 of the row ---- */
        uint32_t r0[9], r1[9], r2[9];
        const int sh = x0 & 3, xa = x0 - sh;
        const bool inside = inter && xa >= 0 && xa + 12 <= W && y0 >= 0 && y0 + 9 <= H;
#ifdef K2_WHATIF_NO_WINDOW_LOADS
        /* measurement only (tools/build_cuda_variant.sh, wrong pictures): the luma window costs nothing — no loads, no
         * alignment shifts — which bounds from below what ANY staging scheme (shared memory, TMA) could make of this kernel */
        if (inter) {
#pragma unroll
            for (int r = 0; r < 9; r++) { r0[r] = mvw * (uint32_t)(r + 1); r1[r] = mvw ^ (uint32_t)(r * 0x01010101); r2[r] = mvw >> r; }
        } else
#endif
        if (inside) {
            const uint8_t *p = ref + (size_t)y0 * W + xa;
#pragma unroll
            for (int r = 0; r < 9; r++) {
                const uint32_t *q = reinterpret_cast<const uint32_t *>(p + (size_t)r * W);
                const uint32_t w0 = __ldg(q), w1 = __ldg(q + 1), w2 = __ldg(q + 2);
                r0[r] = __funnelshift_r(w0, w1, 8 * sh); r1[r] = __funnelshift_r(w1, w2, 8 * sh); r2[r] = w2 >> (8 * sh);
            }
        } else if (inter) {
#pragma unroll 1
            for (int r = 0; r < 9; r++) {
                uint32_t a = 0, c = 0;
                for (int k = 0; k < 4; k++) { a |= ref_px(ref, W, H, x0 + k, y0 + r) << (8 * k); c |= ref_px(ref, W, H, x0 + 4 + k, y0 + r) << (8 * k); }
                const uint32_t e = ref_px(ref, W, H, x0 + 8, y0 + r);
                /* dynamic row index into a register array is not possible: scatter with a switch */
#pragma unroll
                for (int k = 0; k < 9; k++) if (k == r) { r0[k] = a; r1[k] = c; r2[k] = e; }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 9; r++) { r0[r] = r1[r] = r2[r] = 0; }
        }

        /* ---- interpolation on registers (k2_math.cuh; CPU-checked by tests/test_k2_math_cpu.py) ---- */
#ifdef K2_WHATIF_NO_MATH
        /* measurement only: the window is fetched and aligned but the luma interpolation is replaced by a copy of the
         * integer samples — what the kernel costs as a pure data mover */
        (void)useG; (void)useB; (void)useH; (void)useJ; (void)n_ops; (void)anyB; (void)anyH; (void)anyJ;
#pragma unroll
        for (int py = 0; py < 4; py++) out_rows[py] = __funnelshift_r(r0[py + 2], r1[py + 2], 16) ^ (r2[py] & 0u);
#else
        k2m_luma4x4(r0, r1, r2, fx, fy, useG, useB, useH, useJ, n_ops, anyB, anyH, anyJ, out_rows);
#endif
    }

    /* ---- luma residual + store ---- */
    if (inter) {
        uint8_t *dst = job.cur + (size_t)(mby * 16 + by * 4) * W + mbx * 16 + bx * 4;
        if (has_res) {
            const int rw[8] = {res_lo.x, res_lo.y, res_lo.z, res_lo.w, res_hi.x, res_hi.y, res_hi.z, res_hi.w};
#pragma unroll
            for (int py = 0; py < 4; py++) {
                const uint32_t p = out_rows[py];
                const int v0 = clip255((int)(p & 0xff) + (int)(short)(rw[2 * py] & 0xffff)), v1 = clip255((int)((p >> 8) & 0xff) + (rw[2 * py] >> 16));
                const int v2 = clip255((int)((p >> 16) & 0xff) + (int)(short)(rw[2 * py + 1] & 0xffff)), v3 = clip255((int)(p >> 24) + (rw[2 * py + 1] >> 16));
                out_rows[py] = (uint32_t)v0 | ((uint32_t)v1 << 8) | ((uint32_t)v2 << 16) | ((uint32_t)v3 << 24);
            }
        }
#pragma unroll
        for (int py = 0; py < 4; py++) *reinterpret_cast<uint32_t *>(dst + (size_t)py * W) = out_rows[py];
    }

    /* ================================ chroma ================================ */
    if (inter) {
        const int xc = mbx * 8 + bx * 2 + (mvx >> 3), yc = mby * 8 + by * 2 + (mvy >> 3);
        const int cfx = mvx & 7, cfy = mvy & 7;
        const int w00 = (8 - cfx) * (8 - cfy), w01 = cfx * (8 - cfy), w10 = (8 - cfx) * cfy, w11 = cfx * cfy;
        const int sh = xc & 3, xa = xc - sh;
        const bool inside = xa >= 0 && xa + 8 <= CW && yc >= 0 && yc + 3 <= CH;
        const int cb0 = 16 + (by >> 1) * 2 + (bx >> 1);                      /* Cb 4x4 block holding this 2x2 */
        const int roff = (by & 1) * 8 + (bx & 1) * 2;                        /* its position inside that block */
#pragma unroll
        for (int pl = 0; pl < 2; pl++) {
            const uint8_t *rp = ref + ysize + (pl ? csize : 0);
            uint32_t t[3];                                                   /* bytes 0..2 of each window row */
            if (inside) {
#pragma unroll
                for (int r = 0; r < 3; r++) {
                    const uint32_t *q = reinterpret_cast<const uint32_t *>(rp + (size_t)(yc + r) * CW + xa);
                    t[r] = __funnelshift_r(__ldg(q), __ldg(q + 1), 8 * sh);
                }
            } else {
#pragma unroll
                for (int r = 0; r < 3; r++)
                    t[r] = ref_px(rp, CW, CH, xc, yc + r) | (ref_px(rp, CW, CH, xc + 1, yc + r) << 8) | (ref_px(rp, CW, CH, xc + 2, yc + r) << 16);
            }
            const int cb = cb0 + 4 * pl;
            const bool has_r = (mask >> cb) & 1;
            const int16_t *rs = coef + slot_index(mask, cb) * 16 + roff;
            uint8_t *dst = job.cur + ysize + (pl ? csize : 0) + (size_t)(mby * 8 + by * 2) * CW + mbx * 8 + bx * 2;
#pragma unroll
            for (int y = 0; y < 2; y++) {
                const int a0 = t[y] & 0xff, a1 = (t[y] >> 8) & 0xff, a2 = (t[y] >> 16) & 0xff;
                const int c0 = t[y + 1] & 0xff, c1 = (t[y + 1] >> 8) & 0xff, c2 = (t[y + 1] >> 16) & 0xff;
                int v0 = (w00 * a0 + w01 * a1 + w10 * c0 + w11 * c1 + 32) >> 6;
                int v1 = (w00 * a1 + w01 * a2 + w10 * c1 + w11 * c2 + 32) >> 6;
                if (has_r) {
                    const int rr = *reinterpret_cast<const int *>(rs + 4 * y);
                    v0 = clip255(v0 + (int)(short)(rr & 0xffff)); v1 = clip255(v1 + (rr >> 16));
                }
                *reinterpret_cast<uint16_t *>(dst + (size_t)y * CW) = (uint16_t)(v0 | (v1 << 8));
            }
        }
    }
}
