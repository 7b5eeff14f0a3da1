/* k5_rgba.cuh — output formatting on the device: crop to the SPS cropping rectangle and convert
 * I420 -> RGBA, fused into what feeds the device-to-host copy.
 *
 * Device replacement of the wrapper's optional converter (templates/DecoderPost.js:324-565,
 * yuv2rgbcalc :514-560: BT.601 fixed point, 1192(y-16), 1634(v-128), 832(v-128), 400(u-128),
 * 2066(u-128), >>10, clamp, bytes R,G,B,255; chroma taken from (x>>1, y>>1)) and of the cropping the
 * reference only REPORTS (h264bsd_decoder.c:886-917 h264bsdCroppingParams) and leaves to the caller.
 * Pure streaming: 1.5 B read + 4 B written per sample; one thread per 2 rows x 4 columns writes two
 * 16-byte vectors.  HBM-bound.
 */
#pragma once
#include "k_common.cuh"

struct RgbaJob { const uint8_t *frame; uint8_t *out; int W, H, cl, ct, cw, ch; };

__device__ __forceinline__ uint32_t yuv2rgba(int y, int u, int v)
{
    const int a0 = 1192 * (y - 16), a1 = 1634 * (v - 128), a2 = 832 * (v - 128), a3 = 400 * (u - 128), a4 = 2066 * (u - 128);
    const int r = clip255((a0 + a1) >> 10), g = clip255((a0 - a2 - a3) >> 10), b = clip255((a0 + a4) >> 10);
    return 0xff000000u | ((uint32_t)b << 16) | ((uint32_t)g << 8) | (uint32_t)r;
}

__global__ void __launch_bounds__(256) k5_rgba(RgbaJob j)
{
    const int qw = (j.cw + 3) >> 2, qh = (j.ch + 1) >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= qw * qh) return;
    const int qx = idx % qw, qy = idx / qw;
    const int x0 = 4 * qx, y0 = 2 * qy;                         /* inside the cropped picture */
    const int sx = j.cl + x0, sy = j.ct + y0;                   /* inside the coded frame (both even) */
    const uint8_t *Y = j.frame + (size_t)sy * j.W + sx;
    const uint8_t *U = j.frame + (size_t)j.W * j.H + (size_t)(sy >> 1) * (j.W >> 1) + (sx >> 1);
    const uint8_t *V = U + ((size_t)j.W * j.H >> 2);
    const int n = min(4, j.cw - x0);                            /* columns this thread really has */
    int u[2], v[2];
    u[0] = __ldg(U); v[0] = __ldg(V);
    u[1] = n > 2 ? __ldg(U + 1) : u[0]; v[1] = n > 2 ? __ldg(V + 1) : v[0];
#pragma unroll
    for (int r = 0; r < 2; r++) {
        if (y0 + r >= j.ch) break;
        uint32_t px[4];
#pragma unroll
        for (int k = 0; k < 4; k++) px[k] = k < n ? yuv2rgba(__ldg(Y + (size_t)r * j.W + k), u[k >> 1], v[k >> 1]) : 0;
        uint32_t *o = reinterpret_cast<uint32_t *>(j.out) + (size_t)(y0 + r) * j.cw + x0;
        if (n == 4 && (j.cw & 3) == 0) *reinterpret_cast<uint4 *>(o) = make_uint4(px[0], px[1], px[2], px[3]);
        else for (int k = 0; k < n; k++) o[k] = px[k];
    }
}
