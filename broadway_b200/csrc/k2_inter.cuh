/* k2_inter.cuh — kernel family 2: inter prediction + residual add, frame-parallel
 * over every P macroblock of every picture in the batch.
 *
 * Device replacement of h264bsdInterPrediction's partition walk
 * (h264bsd_inter_prediction.c:364-487), h264bsdPredictSamples and the nine luma
 * interpolators (h264bsd_reconstruct.c:1819-1941, :491-1791), PredictChroma
 * (:110-476), the edge-clamped block fetch h264bsdFillBlock (:2222-2314) and
 * h264bsdWriteOutputBlocks (h264bsd_image.c:171-343).
 *
 * One CTA (128 threads) per macroblock.  Prediction is per-sample independent,
 * so every partition shape is handled as sixteen 4x4 blocks with their own
 * vector (the record stores final vectors per 4x4 block):
 *   1. the 16 (9x9 luma) and 32 (3x3 chroma) reference windows, apron included,
 *      are staged into shared memory with coordinate clamping (= the reference's
 *      out-of-frame behaviour);
 *   2. each thread produces two luma samples and one chroma sample from shared
 *      memory: 6-tap (1,-5,20,20,-5,1) half samples with (x+16)>>5, centre sample
 *      from unclipped intermediates with (x+512)>>10, quarter samples as rounded
 *      averages (8.4.2.2.1); chroma bilinear 1/8 pel (8.4.2.2.2);
 *   3. residual (already transformed by K1) is added with clipping and the
 *      macroblock is written from shared memory with 16-byte stores.
 * HBM per inter macroblock: 384 B reference (unique) + 384 B written + 128 B
 * record + 32 B per coded block.
 */
#pragma once
#include "k_common.cuh"

#define K2_THREADS 128
#define K2_LP 12                     /* luma window pitch (9 columns used) */

struct __align__(16) K2Smem {
    h264b200_mb_t rec;
    uint8_t luma[16][9][K2_LP];
    uint8_t chroma[2][16][3][4];
    __align__(16) uint8_t out_y[16][16];
    __align__(16) uint8_t out_c[2][8][8];
};

__device__ __forceinline__ int tap6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

/* t: the 9x9 window of one 4x4 block; integer sample G of output (lx,ly) is t[ly+2][lx+2] */
__device__ __forceinline__ int luma_sample(const uint8_t (*t)[K2_LP], int lx, int ly, int fx, int fy)
{
    const int x = lx + 2, y = ly + 2;
#define PX(dx, dy) ((int)t[y + (dy)][x + (dx)])
#define HB1(dy) tap6(PX(-2, dy), PX(-1, dy), PX(0, dy), PX(1, dy), PX(2, dy), PX(3, dy))
#define VH1(dx) tap6(PX(dx, -2), PX(dx, -1), PX(dx, 0), PX(dx, 1), PX(dx, 2), PX(dx, 3))
    if ((fx | fy) == 0) return PX(0, 0);
    if (fy == 0) {                                   /* a, b, c */
        int b = clip255((HB1(0) + 16) >> 5);
        return fx == 2 ? b : (b + (fx == 1 ? PX(0, 0) : PX(1, 0)) + 1) >> 1;
    }
    if (fx == 0) {                                   /* d, h, n */
        int h = clip255((VH1(0) + 16) >> 5);
        return fy == 2 ? h : (h + (fy == 1 ? PX(0, 0) : PX(0, 1)) + 1) >> 1;
    }
    if (fx == 2 || fy == 2) {                        /* f, i, j, k, q: all need the centre sample j */
        int j = clip255((tap6(HB1(-2), HB1(-1), HB1(0), HB1(1), HB1(2), HB1(3)) + 512) >> 10);
        if (fx == 2 && fy == 2) return j;
        if (fx == 2) return (j + clip255((HB1(fy == 1 ? 0 : 1) + 16) >> 5) + 1) >> 1;      /* f: b above, q: s below */
        return (j + clip255(((fx == 1 ? VH1(0) : VH1(1)) + 16) >> 5) + 1) >> 1;            /* i: h left, k: m right */
    }
    {                                                /* e, g, p, r: diagonal quarter samples */
        int b = clip255((HB1(fy == 1 ? 0 : 1) + 16) >> 5);
        int h = clip255(((fx == 1 ? VH1(0) : VH1(1)) + 16) >> 5);
        return (b + h + 1) >> 1;
    }
#undef PX
#undef HB1
#undef VH1
}

__global__ void __launch_bounds__(K2_THREADS) k2_inter(Batch b)
{
    __shared__ K2Smem s;
    const int tid = threadIdx.x;
    const uint32_t g = blockIdx.x;
    const PicJob &job = b.jobs[find_job(b, g)];
    const uint32_t mbi = g - job.mb_base;
    const h264b200_mb_t *mb = job.mbs + mbi;
    if (__ldg(reinterpret_cast<const uint8_t *>(mb)) != H264B200_MB_INTER) return;      /* CTA-uniform */
    if (tid < 8) reinterpret_cast<int4 *>(&s.rec)[tid] = __ldg(reinterpret_cast<const int4 *>(mb) + tid);
    __syncthreads();

    const int W = job.wm * 16, H = job.hm * 16, CW = W >> 1, CH = H >> 1;
    const int mbx = mbi % job.wm, mby = mbi / job.wm;
    const size_t ysize = (size_t)W * H, csize = (size_t)CW * CH;

    /* ---- 1. stage reference windows (coordinate clamp = h264bsdFillBlock) ---- */
    for (int e = tid; e < 16 * 81; e += K2_THREADS) {
        int blk = e / 81, r = e - blk * 81, ry = r / 9, rx = r - ry * 9;
        int bx = blk & 3, by = blk >> 2;
        const uint8_t *ref = job.frames + (size_t)s.rec.ref_slot[(by >> 1) * 2 + (bx >> 1)] * job.frame_bytes;
        int x = mbx * 16 + bx * 4 + (s.rec.mv[blk][0] >> 2) - 2 + rx;
        int y = mby * 16 + by * 4 + (s.rec.mv[blk][1] >> 2) - 2 + ry;
        x = min(max(x, 0), W - 1); y = min(max(y, 0), H - 1);
        s.luma[blk][ry][rx] = __ldg(ref + (size_t)y * W + x);
    }
    for (int e = tid; e < 2 * 16 * 9; e += K2_THREADS) {
        int pl = e / 144, r0 = e - pl * 144, blk = r0 / 9, r = r0 - blk * 9, ry = r / 3, rx = r - ry * 3;
        int bx = blk & 3, by = blk >> 2;
        const uint8_t *ref = job.frames + (size_t)s.rec.ref_slot[(by >> 1) * 2 + (bx >> 1)] * job.frame_bytes + ysize + (pl ? csize : 0);
        int x = mbx * 8 + bx * 2 + (s.rec.mv[blk][0] >> 3) + rx;
        int y = mby * 8 + by * 2 + (s.rec.mv[blk][1] >> 3) + ry;
        x = min(max(x, 0), CW - 1); y = min(max(y, 0), CH - 1);
        s.chroma[pl][blk][ry][rx] = __ldg(ref + (size_t)y * CW + x);
    }
    __syncthreads();

    /* ---- 2. interpolate + residual ---- */
    const uint32_t mask = s.rec.resid_mask;
    const int16_t *coef = job.coef + (size_t)s.rec.coef_offset * 16;
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {           /* a warp covers two whole 4x4 blocks: <= 2-way divergence */
        int blk = pass * 8 + (tid >> 4), p = tid & 15, lx = p & 3, ly = p >> 2;
        int v = luma_sample(s.luma[blk], lx, ly, s.rec.mv[blk][0] & 3, s.rec.mv[blk][1] & 3);
        int bi = (blk & 1) | ((blk & 2) << 1) | ((blk & 4) >> 1) | (blk & 8);            /* raster -> luma4x4BlkIdx */
        if ((mask >> bi) & 1) v = clip255(v + coef[slot_index(mask, bi) * 16 + p]);
        s.out_y[(blk >> 2) * 4 + ly][(blk & 3) * 4 + lx] = (uint8_t)v;
    }
    {
        int pl = tid >> 6, q = tid & 63, blk = q >> 2, x = q & 1, y = (q >> 1) & 1;
        int fx = s.rec.mv[blk][0] & 7, fy = s.rec.mv[blk][1] & 7;
        const uint8_t (*t)[4] = s.chroma[pl][blk];
        int v = ((8 - fx) * (8 - fy) * t[y][x] + fx * (8 - fy) * t[y][x + 1] + (8 - fx) * fy * t[y + 1][x] + fx * fy * t[y + 1][x + 1] + 32) >> 6;
        int bx = blk & 3, by = blk >> 2, cb = 16 + 4 * pl + (by >> 1) * 2 + (bx >> 1);
        if ((mask >> cb) & 1) v = clip255(v + coef[slot_index(mask, cb) * 16 + ((by & 1) * 2 + y) * 4 + (bx & 1) * 2 + x]);
        s.out_c[pl][by * 2 + y][bx * 2 + x] = (uint8_t)v;
    }
    __syncthreads();

    /* ---- 3. write the macroblock: 16-byte luma rows, 8-byte chroma rows ---- */
    if (tid < 16) {
        *reinterpret_cast<int4 *>(job.cur + (size_t)(mby * 16 + tid) * W + mbx * 16) = *reinterpret_cast<const int4 *>(s.out_y[tid]);
    } else if (tid < 32) {
        int pl = (tid - 16) >> 3, r = tid & 7;
        *reinterpret_cast<int2 *>(job.cur + ysize + (pl ? csize : 0) + (size_t)(mby * 8 + r) * CW + mbx * 8) = *reinterpret_cast<const int2 *>(s.out_c[pl][r]);
    }
}
