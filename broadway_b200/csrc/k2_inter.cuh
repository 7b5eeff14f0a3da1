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

#define K2_THREADS 128

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return (a + f) - 5 * (b + e) + 20 * (c + d); }

/* four unsigned bytes of a times four signed bytes of b, accumulated (dp4a.u32.s32) */
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

/* clamped single-sample fetch: the reference's out-of-picture behaviour (h264bsdFillBlock) */
__device__ __forceinline__ uint32_t ref_px(const uint8_t *pl, int w, int h, int x, int y)
{
    x = min(max(x, 0), w - 1); y = min(max(y, 0), h - 1);
    return __ldg(pl + (size_t)y * w + x);
}

/* The work of one thread: the 4x4 luma block `blk` (raster) of batch-wide macroblock g0 and the two 2x2 chroma
 * blocks under it.  STAGED == false: results go straight to the frame.  STAGED == true (k2_inter, below): they go
 * to the CTA's staging arrays sy / sc at index `slot`, to be stored by the thread that owns the block's position. */
template <bool STAGED>
__device__ __forceinline__ void k2_block(const Batch &b, uint32_t g0, int blk, int4 *sy, uint2 *sc, int slot)
{
    const unsigned FULL = 0xffffffffu;
    const int bx = blk & 3, by = blk >> 2;
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
        const int dn = fy == 3, rt = fx == 3;
        const bool outer = !STAGED || anyH || anyJ;   /* window rows 0, 1, 7, 8 are only inputs of the vertical filters (skipping them did not pay in the unsorted kernel) */

        /* ---- window rows as three byte-aligned words: bytes 0..8 of the row ---- */
        uint32_t r0[9], r1[9], r2[9];
        const int sh = x0 & 3, xa = x0 - sh;
        const bool inside = inter && xa >= 0 && xa + 12 <= W && y0 >= 0 && y0 + 9 <= H;
        if (inside) {
            const uint8_t *p = ref + (size_t)y0 * W + xa;
#pragma unroll
            for (int r = 0; r < 9; r++) {
                if (!outer && (r < 2 || r > 6)) { r0[r] = r1[r] = r2[r] = 0; continue; }      /* warp uniform */
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

        /* ---- the 4 samples at columns x+rt .. x+rt+3 of every window row (G' and the inputs of h') ---- */
        uint32_t cw[9];
#pragma unroll
        for (int r = 0; r < 9; r++) cw[r] = __funnelshift_r(r0[r], r1[r], 8 * (2 + rt));

        /* ---- horizontal 6-tap sums: hs[r][k] for output column k of window row r ---- */
        int hs[9][4];
        if (anyB || anyJ) {
            const int T0 = 0x1414fb01, T1 = 0x000001fb;      /* (1,-5,20,20) and (-5,1,0,0) as signed bytes, low byte first */
#pragma unroll
            for (int r = 0; r < 9; r++) {
                if (!anyJ && (r < 2 || r > 6)) { hs[r][0] = hs[r][1] = hs[r][2] = hs[r][3] = 0; continue; }
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t lo = k ? __funnelshift_r(r0[r], r1[r], 8 * k) : r0[r];
                    const uint32_t hi = k ? __funnelshift_r(r1[r], r2[r], 8 * k) : r1[r];
                    hs[r][k] = dp4a_us(lo, T0, dp4a_us(hi, T1, 0));
                }
            }
        }

        /* ---- per output row: operands and the final blend ---- */
#pragma unroll
        for (int py = 0; py < 4; py++) {
            int bq[4] = {0, 0, 0, 0}, hq[4] = {0, 0, 0, 0}, jq[4] = {0, 0, 0, 0};
            if (anyB) {
#pragma unroll
                for (int k = 0; k < 4; k++) bq[k] = clip255(((dn ? hs[py + 3][k] : hs[py + 2][k]) + 16) >> 5);
            }
            if (anyH) {
                /* vertical 6-tap over raw samples, two samples per instruction on biased 16-bit lanes:
                 * (a+f) + 20(c+d) + 2560 - 5(b+e) stays within [0, 65535] per lane */
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    uint32_t e[6];
#pragma unroll
                    for (int t = 0; t < 6; t++) e[t] = (cw[py + t] >> (8 * half)) & 0x00ff00ffu;
                    const uint32_t s = (e[0] + e[5] + 0x0a000a00u) + 20u * (e[2] + e[3]) - 5u * (e[1] + e[4]);
                    hq[half] = clip255(((int)(s & 0xffff) - 2560 + 16) >> 5);
                    hq[half + 2] = clip255(((int)(s >> 16) - 2560 + 16) >> 5);
                }
            }
            if (anyJ) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    jq[k] = clip255((tap6(hs[py][k], hs[py + 1][k], hs[py + 2][k], hs[py + 3][k], hs[py + 4][k], hs[py + 5][k]) + 512) >> 10);
            }
            const uint32_t gw = dn ? cw[py + 3] : cw[py + 2];
            uint32_t pk = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) {
                int s = 0;
                if (useG) s += (gw >> (8 * k)) & 0xff;
                if (useB) s += bq[k];
                if (useH) s += hq[k];
                if (useJ) s += jq[k];
                const int v = (s * (3 - n_ops) + 1) >> 1;
                pk |= (uint32_t)v << (8 * k);
            }
            out_rows[py] = pk;
        }
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
        if (STAGED) sy[slot] = make_int4((int)out_rows[0], (int)out_rows[1], (int)out_rows[2], (int)out_rows[3]);
        else {
#pragma unroll
            for (int py = 0; py < 4; py++) *reinterpret_cast<uint32_t *>(dst + (size_t)py * W) = out_rows[py];
        }
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
        uint32_t cpk[2];                                                     /* per plane: row 0 | row 1 << 16 */
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
            cpk[pl] = 0;
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
                if (STAGED) cpk[pl] |= (uint32_t)(v0 | (v1 << 8)) << (16 * y);
                else *reinterpret_cast<uint16_t *>(dst + (size_t)y * CW) = (uint16_t)(v0 | (v1 << 8));
            }
        }
        if (STAGED) sc[slot] = make_uint2(cpk[0], cpk[1]);
    }
}

/* A half-warp is one macroblock, a warp two horizontally adjacent macroblocks; every thread keeps the block its
 * index names and stores it itself.  8 CTAs per SM at 64 registers (a few spilled words): measured 2.19 ms per 256
 * pictures against 2.24 ms at 80 and 2.33 ms at 96 registers without spills — the loads of the reference window are
 * what the extra warps hide.  (An L2 prefetch of the records of later CTAs made no difference.) */
__global__ void __launch_bounds__(K2_THREADS, 8) k2_inter(Batch b)
{
    k2_block<false>(b, (blockIdx.x * K2_THREADS + threadIdx.x) >> 4, threadIdx.x & 15, nullptr, nullptr, 0);
}

/* EXPERIMENT, not the default (H264B200_K2=sorted selects it; bit-exact, the GPU parity suite passes with it):
 * the blocks of a CTA (16 macroblocks = 256 blocks) are SORTED BY FRACTIONAL-POSITION CLASS before they are
 * interpolated.  Which operand families (horizontal half b, vertical half h, centre j) a warp computes is a warp
 * vote; with one macroblock pair per warp and vectors that differ per partition, nearly every warp needs all
 * three and every thread pays for the worst position.  After a counting sort over the seven classes
 *     G | b | b+h | h | h+j | j | j+b      (neighbours differ by one family; class 7: not an inter block)
 * a warp holds one or two classes and the votes switch most of the arithmetic off; results return through
 * shared memory to the thread that owns the block's position, so the frame is still written as full 32-byte
 * sectors per warp row store.
 * Measured on B200, default bench, per 256 pictures: 2.69 ms against 2.22 ms for k2_inter.  The kernel is bound by
 * the latency of its dependent loads (record -> window -> residual; ncu: 40 % of the samples on the long
 * scoreboard), not by the arithmetic the sort removes, and the warps of a CTA now finish at very different times
 * (a G warp does a sixth of a j+h warp's work) while their registers stay allocated until the slowest one ends,
 * so fewer warps are left to hide that latency.  A version that balances the classes across warps (several tiles
 * per CTA, chunks of sorted blocks claimed dynamically) is the open follow-up. */
#define K2S_THREADS 256
__global__ void __launch_bounds__(K2S_THREADS, 4) k2_inter_sorted(Batch b)
{
    __shared__ int4 sy[K2S_THREADS];
    __shared__ uint2 sc[K2S_THREADS];
    __shared__ uint16_t perm[K2S_THREADS];
    __shared__ uint32_t hist[64];                /* [class][warp] counts, then exclusive offsets */
    const unsigned FULL = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t mb0 = blockIdx.x * (K2S_THREADS / 16);
    const uint32_t gA = mb0 + (tid >> 4);
    const bool inA = gA < b.total_mbs;
    const uint32_t gc = inA ? gA : b.total_mbs - 1;
    bool interA;
    int cls = 7;
    {
        const PicJob &job = b.jobs[find_job(b, gc)];
        const uint32_t *rec = reinterpret_cast<const uint32_t *>(job.mbs + (gc - job.mb_base));
        interA = inA && (__ldg(rec) & 0xff) == H264B200_MB_INTER;
        if (interA) {
            const uint32_t mvw = __ldg(rec + 16 + (tid & 15));
            const int fx = mvw & 3, fy = (mvw >> 16) & 3;
            const int useJ = (fx == 2 && fy != 0) || (fy == 2 && fx != 0), useB = fx != 0 && fy != 2, useH = fy != 0 && fx != 2;
            cls = (0x04652310u >> (4 * (useB | (useH << 1) | (useJ << 2)))) & 7;
        }
    }
    if (!__syncthreads_or(interA)) return;
    /* counting sort, stable in thread order: rank inside the warp from ballots, class/warp offsets from one warp scan */
    int rank = 0; uint32_t cnt = 0;
#pragma unroll
    for (int c = 0; c < 8; c++) {
        const unsigned m = __ballot_sync(FULL, cls == c);
        if (cls == c) rank = __popc(m & ((1u << lane) - 1u));
        if (lane == c) cnt = __popc(m);
    }
    if (lane < 8) hist[lane * 8 + warp] = cnt;
    __syncthreads();
    if (warp == 0) {
        const uint32_t a = hist[2 * lane], c2 = hist[2 * lane + 1];
        uint32_t incl = a + c2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(FULL, incl, d); if (lane >= d) incl += v; }
        hist[2 * lane] = incl - a - c2; hist[2 * lane + 1] = incl - c2;
    }
    __syncthreads();
    perm[hist[cls * 8 + warp] + rank] = (uint16_t)tid;
    __syncthreads();
    const int t2 = perm[tid];
    k2_block<true>(b, mb0 + (t2 >> 4), t2 & 15, sy, sc, t2);
    __syncthreads();
    if (interA) {
        const PicJob &job = b.jobs[find_job(b, gc)];
        const uint32_t mbi = gc - job.mb_base;
        const int W = job.wm * 16, H = job.hm * 16, CW = W >> 1;
        const int mbx = mbi % job.wm, mby = mbi / job.wm, blk = tid & 15, bx = blk & 3, by = blk >> 2;
        const size_t ysize = (size_t)W * H, csize = ysize >> 2;
        const int4 y = sy[tid];
        const uint2 c = sc[tid];
        uint8_t *dst = job.cur + (size_t)(mby * 16 + by * 4) * W + mbx * 16 + bx * 4;
        *reinterpret_cast<uint32_t *>(dst) = (uint32_t)y.x;
        *reinterpret_cast<uint32_t *>(dst + (size_t)W) = (uint32_t)y.y;
        *reinterpret_cast<uint32_t *>(dst + (size_t)2 * W) = (uint32_t)y.z;
        *reinterpret_cast<uint32_t *>(dst + (size_t)3 * W) = (uint32_t)y.w;
        uint8_t *dc = job.cur + ysize + (size_t)(mby * 8 + by * 2) * CW + mbx * 8 + bx * 2;
        *reinterpret_cast<uint16_t *>(dc) = (uint16_t)c.x;
        *reinterpret_cast<uint16_t *>(dc + CW) = (uint16_t)(c.x >> 16);
        *reinterpret_cast<uint16_t *>(dc + csize) = (uint16_t)c.y;
        *reinterpret_cast<uint16_t *>(dc + csize + CW) = (uint16_t)(c.y >> 16);
    }
}
