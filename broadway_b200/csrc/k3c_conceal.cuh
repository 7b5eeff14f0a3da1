/* k3c_conceal.cuh — spatial concealment of lost macroblocks (I pictures, or P pictures without a usable
 * reference): device replacement of ConcealMb's interpolation path and Transform
 * (h264bsd_conceal.c:330-631).  Runs after K2/K3 and before K4, only for pictures that lost slices.
 *
 * The reference conceals in a fixed serial order and every concealed macroblock becomes a source for the
 * next ones, so this is one warp per picture walking the host-prepared order list (the host knows which
 * macroblocks were decoded; include/h264b200_records.h H264B200_MB_CONCEAL).  Per macroblock and plane: the
 * 4 (2) sample sums along each usable side give a DC and two gradient terms, a reduced 4x4 inverse
 * transform spreads them over a 4x4 grid, each grid value fills a 4x4 (2x2) patch.
 */
#pragma once
#include "k_common.cuh"

/* grid[16] from the side sums; g = samples per sum (4 luma, 2 chroma) */
__device__ __forceinline__ void conceal_grid(int (&fp)[16], const int *a, const int *bl, const int *l, const int *r,
                                             bool A, bool B, bool L, bool R, bool luma)
{
    int j = 0, hor = 0, ver = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) fp[k] = 0;
    const int sa = a[0] + a[1] + a[2] + a[3], sb = bl[0] + bl[1] + bl[2] + bl[3], sl = l[0] + l[1] + l[2] + l[3], sr = r[0] + r[1] + r[2] + r[3];
    if (A) { j++; hor++; fp[0] += sa; fp[1] += a[0] + a[1] - a[2] - a[3]; }
    if (B) { j++; hor++; fp[0] += sb; fp[1] += bl[0] + bl[1] - bl[2] - bl[3]; }
    if (L) { j++; ver++; fp[0] += sl; fp[4] += l[0] + l[1] - l[2] - l[3]; }
    if (R) { j++; ver++; fp[0] += sr; fp[4] += r[0] + r[1] - r[2] - r[3]; }
    const int sh = luma ? 5 : 4, base = luma ? 3 : 2;
    if (!hor && L && R) fp[1] = (sl - sr) >> sh; else if (hor) fp[1] >>= (base + hor);
    if (!ver && A && B) fp[4] = (sa - sb) >> sh; else if (ver) fp[4] >>= (base + ver);
    const int dsh = luma ? 4 : 3;
    if (j == 3) fp[0] = (21 * fp[0]) >> (luma ? 10 : 9); else fp[0] >>= (dsh + (j == 1 ? 0 : j == 2 ? 1 : 2));
    /* Transform (:590-631): only DC, lowest horizontal and lowest vertical term can be non-zero */
    if (!fp[1] && !fp[4]) {
#pragma unroll
        for (int k = 1; k < 16; k++) fp[k] = fp[0];
        return;
    }
    const int t0 = fp[0], t1 = fp[1], v = fp[4];
    fp[0] = t0 + t1; fp[1] = t0 + (t1 >> 1); fp[2] = t0 - (t1 >> 1); fp[3] = t0 - t1;
    fp[4] = fp[5] = fp[6] = fp[7] = v;
#pragma unroll
    for (int c = 0; c < 4; c++) {
        const int u0 = fp[c], u1 = fp[4 + c];
        fp[c] = u0 + u1; fp[4 + c] = u0 + (u1 >> 1); fp[8 + c] = u0 - (u1 >> 1); fp[12 + c] = u0 - u1;
    }
}

__global__ void __launch_bounds__(32) k3c_conceal(Batch b)
{
    const PicJob &job = b.jobs[blockIdx.x];
    if (!job.n_conceal) return;
    const int lane = threadIdx.x;
    const int W = job.wm * 16, H = job.hm * 16;
    const size_t ysize = (size_t)W * H, csize = ysize >> 2;
    for (uint32_t e = 0; e < job.n_conceal; e++) {
        const uint32_t addr = __ldg(job.conceal_list + e);
        const int mbx = addr % job.wm, mby = addr / job.wm;
        const int fl = __ldg(reinterpret_cast<const uint8_t *>(job.mbs + addr) + 8);       /* record byte 8: avail = H264B200_CN_* */
        const bool A = fl & H264B200_CN_ABOVE, B = fl & H264B200_CN_BELOW, L = fl & H264B200_CN_LEFT, R = fl & H264B200_CN_RIGHT;
#pragma unroll
        for (int comp = 0; comp < 3; comp++) {
            const bool luma = comp == 0;
            const int S = luma ? 16 : 8, g = S >> 2, st = luma ? W : W >> 1;
            uint8_t *P = job.cur + (luma ? 0 : ysize + (comp == 2 ? csize : 0));
            const int x0 = mbx * S, y0 = mby * S;
            /* lane k (0..3): above sums, 4..7: below, 8..11: left, 12..15: right */
            int sum = 0;
            if (lane < 16) {
                const int side = lane >> 2, k = lane & 3;
                const bool on = side == 0 ? A : side == 1 ? B : side == 2 ? L : R;
                if (on) for (int t = 0; t < g; t++) {
                    const int o = k * g + t;
                    const uint8_t *p = side == 0 ? P + (size_t)(y0 - 1) * st + x0 + o : side == 1 ? P + (size_t)(y0 + S) * st + x0 + o
                                     : side == 2 ? P + (size_t)(y0 + o) * st + x0 - 1 : P + (size_t)(y0 + o) * st + x0 + S;
                    sum += __ldcg(p);
                }
            }
            int a[4], bl[4], l[4], r[4], fp[16];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                a[k] = __shfl_sync(0xffffffffu, sum, k); bl[k] = __shfl_sync(0xffffffffu, sum, 4 + k);
                l[k] = __shfl_sync(0xffffffffu, sum, 8 + k); r[k] = __shfl_sync(0xffffffffu, sum, 12 + k);
            }
            conceal_grid(fp, a, bl, l, r, A, B, L, R, luma);
            if (luma) {                                   /* lane: row lane>>1, columns (lane&1)*8..+7 = two grid cells */
                const int y = lane >> 1, cx = (lane & 1) * 2, gy = y >> 2;
                int v0 = 0, v1 = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) { if (k == 4 * gy + cx) v0 = fp[k]; if (k == 4 * gy + cx + 1) v1 = fp[k]; }
                v0 = clip255(v0); v1 = clip255(v1);
                *reinterpret_cast<uint2 *>(P + (size_t)(y0 + y) * st + x0 + (lane & 1) * 8) = make_uint2(0x01010101u * (uint32_t)v0, 0x01010101u * (uint32_t)v1);
            } else if (lane < 16) {                       /* lane: row lane>>1 (0..7), columns (lane&1)*4..+3 = two grid cells of 2 */
                const int y = lane >> 1, cx = (lane & 1) * 2, gy = y >> 1;
                int v0 = 0, v1 = 0;
#pragma unroll
                for (int k = 0; k < 16; k++) { if (k == 4 * gy + cx) v0 = fp[k]; if (k == 4 * gy + cx + 1) v1 = fp[k]; }
                v0 = clip255(v0); v1 = clip255(v1);
                *reinterpret_cast<uint32_t *>(P + (size_t)(y0 + y) * st + x0 + (lane & 1) * 4) = 0x00000101u * (uint32_t)v0 | 0x01010000u * (uint32_t)v1;
            }
        }
        __syncwarp();
        __threadfence_block();
    }
}
