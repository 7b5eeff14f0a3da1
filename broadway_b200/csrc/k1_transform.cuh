/* k1_transform.cuh — kernel family 1: dequantisation + inverse transforms.
 *
 * Device replacement of ProcessResidual (h264bsd_macroblock_layer.c:1343-1424),
 * h264bsdProcessBlock (h264bsd_transform.c:94-231), h264bsdProcessLumaDc
 * (:252-335) and h264bsdProcessChromaDc (:356-398), batched over every
 * macroblock of every picture of a launch.
 *
 * Eight lanes per macroblock (four macroblocks per warp): a P macroblock
 * carries ~7 coded 4x4 blocks out of 24, so one lane per POTENTIAL block would
 * leave three quarters of a warp idle.  Lane j of the octet takes the j-th,
 * (j+8)-th, ... coded block of the macroblock (set bits of resid_mask, which is
 * also the order of the coefficient slots): two 16-byte loads of its int16 slot,
 * dequant with LevelScale(qP%6,pos) << qP/6, row and column butterflies in
 * registers, (x+32)>>6, two 16-byte stores IN PLACE.  Levels arrive already in
 * raster order (the host parser un-zig-zags while it writes).
 * DC terms: a block of an Intra16x16 macroblock (or a chroma block of a
 * macroblock with chroma DC) needs ONE element of the 4x4 (2x2) Hadamard of
 * the DC slot; the lane computes that element directly from the 16 (4) DC
 * levels — cheaper than a cross-lane transform that most macroblocks never use.
 * The reference's DC-only / first-row fast paths (:188-227) are arithmetic
 * shortcuts of the same transform and are not reproduced.  A residual outside
 * [-512,511] (the reference's error return, :181-185) raises bit 0 of the batch
 * error word.
 * HBM: 32 B in + 32 B out per coded block, plus the first 32-byte sector of the
 * 128-byte macroblock record.
 */
#pragma once
#include "k_common.cuh"

__device__ __forceinline__ void idct4x4_regs(int (&d)[16])
{
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int e0 = d[4*i] + d[4*i+2], e1 = d[4*i] - d[4*i+2], e2 = (d[4*i+1] >> 1) - d[4*i+3], e3 = d[4*i+1] + (d[4*i+3] >> 1);
        d[4*i] = e0 + e3; d[4*i+1] = e1 + e2; d[4*i+2] = e1 - e2; d[4*i+3] = e0 - e3;
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int e0 = d[i] + d[8+i], e1 = d[i] - d[8+i], e2 = (d[4+i] >> 1) - d[12+i], e3 = d[4+i] + (d[12+i] >> 1);
        d[i] = (e0 + e3 + 32) >> 6; d[4+i] = (e1 + e2 + 32) >> 6; d[8+i] = (e1 - e2 + 32) >> 6; d[12+i] = (e0 - e3 + 32) >> 6;
    }
}

/* element (row r, column c) of H * X * H with H = [1 1 1 1; 1 1 -1 -1; 1 -1 -1 1; 1 -1 1 -1] (8.5.10): sign tables by row */
__device__ __forceinline__ int hadamard4_elem(const int16_t *x, int r, int c)
{
    /* sign(k, i) of H[k][i]: k=0: ++++, k=1: ++--, k=2: +--+, k=3: +-+-; four bits per row, bit i set = minus */
    const unsigned sr = (0xA6C0u >> (4 * r)) & 0xfu, sc = (0xA6C0u >> (4 * c)) & 0xfu;
    int acc = 0;
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int v = x[4 * i + j];
            acc += (((sr >> i) ^ (sc >> j)) & 1) ? -v : v;
        }
    return acc;
}

/* Occupancy, measured on B200 per 256-picture launch of the default bench: 4 / 5 / 6 / 8 CTAs per SM
 * (61 / 48 / 40 / 32 registers, the last with spills) = 0.518 / 0.449 / 0.408 / 0.410 ms. */
#ifndef K1_MINB
#define K1_MINB 6
#endif
__global__ void __launch_bounds__(256, K1_MINB) k1_transform(Batch b)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t g = t >> 3;                                   /* batch-wide macroblock index: 8 lanes each */
    const int l8 = threadIdx.x & 7;
    if (g >= b.total_mbs) return;
    const PicJob &job = b.jobs[find_job(b, g)];
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(job.mbs + (g - job.mb_base));
    const uint32_t w0 = __ldg(rec), mask = __ldg(rec + 4);
    const int cls = w0 & 0xff;
    uint32_t blocks = mask & 0xffffffu;
    if (cls == H264B200_MB_IPCM || cls == H264B200_MB_MISSING || !blocks) return;
    const uint32_t coef_off = __ldg(rec + 3);
    const int qp = (w0 >> 8) & 0xff, qpc = (w0 >> 16) & 0xff;
    const bool has_ldc = mask & H264B200_RESID_LUMA_DC, has_cdc = mask & H264B200_RESID_CHROMA_DC;
    const int16_t *mb_in = job.coef_in + (size_t)coef_off * 16;
    int16_t *mb_out = job.coef + (size_t)coef_off * 16;
    const int16_t *cdc_slot = mb_in + 16 * ((has_ldc ? 1 : 0) + __popc(mask & 0xffffu));

    /* strip the l8 lowest set bits: this lane's first block */
    for (int k = 0; k < l8 && blocks; k++) blocks &= blocks - 1;
    int j = l8;                                                  /* ordinal of the block among the coded blocks */
    while (blocks) {
        const int blk = __ffs(blocks) - 1;
        const bool chroma = blk >= 16;
        const int q = chroma ? qpc : qp, sh = q / 6, m = q % 6;
        const int ls0 = H264_LEVEL_SCALE[m][0], ls1 = H264_LEVEL_SCALE[m][1], ls2 = H264_LEVEL_SCALE[m][2];
        const size_t so = (size_t)((has_ldc ? 1 : 0) + j + ((chroma && has_cdc) ? 1 : 0)) * 16;
        const int4 lo = *reinterpret_cast<const int4 *>(mb_in + so), hi = *reinterpret_cast<const int4 *>(mb_in + so + 8);
        int d[16];
        {
            const int wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
            for (int i = 0; i < 8; i++) { d[2*i] = (int)(short)(wv[i] & 0xffff); d[2*i+1] = wv[i] >> 16; }
        }
        /* LevelScale class per raster position: 0 2 0 2 / 2 1 2 1 / 0 2 0 2 / 2 1 2 1 */
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int c = ((i & 5) == 0) ? ls0 : ((i & 5) == 5) ? ls1 : ls2;
            d[i] = (d[i] * c) << sh;
        }
        if (!chroma && has_ldc) {
            /* Intra16x16: DC of the block at raster position (bx4, by4) = element (by4, bx4) of the Hadamard of the DC slot */
            const int bx4 = (blk & 1) | ((blk >> 1) & 2), by4 = ((blk >> 1) & 1) | ((blk >> 2) & 2);
            const int v = hadamard4_elem(mb_in, by4, bx4) * H264_LEVEL_SCALE[qp % 6][0];
            d[0] = qp >= 12 ? v << (qp / 6 - 2) : (v + (1 << (1 - qp / 6))) >> (2 - qp / 6);
        } else if (chroma && has_cdc) {
            /* 2x2: f = [c0+c1+c2+c3, c0-c1+c2-c3, c0+c1-c2-c3, c0-c1-c2+c3] for blocks 0..3 of the plane */
            const int k = (blk - 16) & 3;
            const int16_t *c = cdc_slot + 4 * ((blk - 16) >> 2);
            const int c0 = c[0], c1 = c[1], c2 = c[2], c3 = c[3];
            const int f = c0 + ((k & 1) ? -c1 : c1) + ((k & 2) ? -c2 : c2) + ((k == 1 || k == 2) ? -c3 : c3);
            const int v = f * H264_LEVEL_SCALE[qpc % 6][0];
            d[0] = qpc >= 6 ? v << (qpc / 6 - 1) : v >> 1;
        }
        idct4x4_regs(d);
        int bad = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) bad |= (d[i] < -512) | (d[i] > 511);
        if (bad) atomicOr(b.error_flags, 1u);
        int4 o0, o1;
        o0.x = (d[0] & 0xffff) | (d[1] << 16);   o0.y = (d[2] & 0xffff) | (d[3] << 16);
        o0.z = (d[4] & 0xffff) | (d[5] << 16);   o0.w = (d[6] & 0xffff) | (d[7] << 16);
        o1.x = (d[8] & 0xffff) | (d[9] << 16);   o1.y = (d[10] & 0xffff) | (d[11] << 16);
        o1.z = (d[12] & 0xffff) | (d[13] << 16); o1.w = (d[14] & 0xffff) | (d[15] << 16);
        *reinterpret_cast<int4 *>(mb_out + so) = o0; *reinterpret_cast<int4 *>(mb_out + so + 8) = o1;
        /* advance to this lane's next block: 8 set bits further */
        for (int k = 0; k < 8 && blocks; k++) blocks &= blocks - 1;
        j += 8;
    }
}
