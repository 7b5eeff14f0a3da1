/* k1_transform.cuh — kernel family 1: dequantisation + inverse transforms.
 *
 * Device replacement of ProcessResidual (h264bsd_macroblock_layer.c:1343-1424),
 * h264bsdProcessBlock (h264bsd_transform.c:94-231), h264bsdProcessLumaDc
 * (:252-335) and h264bsdProcessChromaDc (:356-398), batched over every
 * macroblock of every picture of a launch.  One warp per macroblock:
 *   - lanes 0..15 hold the Intra16x16 DC matrix and run the 4x4 Hadamard as
 *     warp-shuffle butterflies (xor 1,2 = rows; xor 4,8 = columns);
 *   - lanes 0..7 hold the two 2x2 chroma DC matrices (xor 1,2);
 *   - lane l < 24 then owns 4x4 block l (luma4x4BlkIdx 0..15, Cb 16..19,
 *     Cr 20..23): two 16-byte loads of its int16 slot, dequant with
 *     LevelScale(qP%6,pos) << qP/6, the DC injected by shuffle, row and column
 *     butterflies in registers, (x+32)>>6, two 16-byte stores IN PLACE.
 * Levels arrive already in raster order (the host parser un-zig-zags while it
 * writes), so there is no scatter here.  The reference's DC-only / first-row
 * fast paths (:188-227) are arithmetic shortcuts of the same transform and are
 * not reproduced.  A residual outside [-512,511] (the reference's error return,
 * :181-185) raises bit 0 of the batch error word.
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

__global__ void __launch_bounds__(256) k1_transform(Batch b)
{
    const unsigned FULL = 0xffffffffu;
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (g >= b.total_mbs) return;
    const PicJob &job = b.jobs[find_job(b, g)];
    const uint32_t *rec = reinterpret_cast<const uint32_t *>(job.mbs + (g - job.mb_base));
    uint32_t w0 = __ldg(rec), coef_off = __ldg(rec + 3), mask = __ldg(rec + 4);
    int cls = w0 & 0xff;
    if (cls == H264B200_MB_IPCM || cls == H264B200_MB_MISSING || (mask & 0x3ffffffu) == 0) return;   /* warp-uniform */
    int qp = (w0 >> 8) & 0xff, qpc = (w0 >> 16) & 0xff;
    const int16_t *base = job.coef_in + (size_t)coef_off * 16;
    bool has_ldc = mask & H264B200_RESID_LUMA_DC, has_cdc = mask & H264B200_RESID_CHROMA_DC;

    /* ---- Intra16x16 luma DC: 4x4 Hadamard by shuffles, then scaling (8.5.10) ---- */
    int dcy = 0;
    if (has_ldc) {                                   /* warp-uniform branch: all lanes shuffle */
        int v = lane < 16 ? base[lane] : 0, t;
        t = __shfl_xor_sync(FULL, v, 1); v = (lane & 1) ? t - v : v + t;
        t = __shfl_xor_sync(FULL, v, 2); v = (lane & 2) ? t - v : v + t;
        /* natural-order WHT -> H = [1 1 1 1; 1 1 -1 -1; 1 -1 -1 1; 1 -1 1 -1]: take y0,y2,y3,y1 */
        v = __shfl_sync(FULL, v, (lane & ~3) | ((0x78 >> (2 * (lane & 3))) & 3));
        t = __shfl_xor_sync(FULL, v, 4); v = (lane & 4) ? t - v : v + t;
        t = __shfl_xor_sync(FULL, v, 8); v = (lane & 8) ? t - v : v + t;
        v = __shfl_sync(FULL, v, (lane & ~12) | (((0x78 >> (2 * ((lane >> 2) & 3))) & 3) << 2));
        v *= H264_LEVEL_SCALE[qp % 6][0];
        dcy = qp >= 12 ? v << (qp / 6 - 2) : (v + (1 << (1 - qp / 6))) >> (2 - qp / 6);   /* lane = raster block position */
        base += 16;
    }
    /* block lane l needs the DC of raster position BLK_TO_RASTER[l] */
    int my_dc = __shfl_sync(FULL, dcy, lane < 16 ? ((lane & 1) | ((lane & 2) << 1) | ((lane & 4) >> 1) | (lane & 8)) : 0);

    /* ---- chroma DC: two 2x2 transforms (8.5.11) ---- */
    int dcc = 0;
    const int16_t *cdc_slot = base + 16 * __popc(mask & 0xffffu);
    if (has_cdc) {
        int v = lane < 8 ? cdc_slot[lane] : 0, t;
        t = __shfl_xor_sync(FULL, v, 1); v = (lane & 1) ? t - v : v + t;
        t = __shfl_xor_sync(FULL, v, 2); v = (lane & 2) ? t - v : v + t;
        v *= H264_LEVEL_SCALE[qpc % 6][0];
        dcc = qpc >= 6 ? v << (qpc / 6 - 1) : v >> 1;   /* lane 4*plane + k */
    }
    int my_cdc = __shfl_sync(FULL, dcc, lane >= 16 && lane < 24 ? lane - 16 : 0);

    if (lane >= 24 || !((mask >> lane) & 1)) return;
    bool chroma = lane >= 16;
    int q = chroma ? qpc : qp;
    int sh = q / 6, m = q % 6;
    int ls0 = H264_LEVEL_SCALE[m][0], ls1 = H264_LEVEL_SCALE[m][1], ls2 = H264_LEVEL_SCALE[m][2];
    const size_t so = ((size_t)coef_off + slot_index(mask, lane)) * 16;
    const int16_t *slot_in = job.coef_in + so;
    int16_t *slot = job.coef + so;
    int4 lo = *reinterpret_cast<const int4 *>(slot_in), hi = *reinterpret_cast<const int4 *>(slot_in + 8);
    int d[16];
    {
        const int wv[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int i = 0; i < 8; i++) { d[2*i] = (int)(short)(wv[i] & 0xffff); d[2*i+1] = wv[i] >> 16; }
    }
    /* LevelScale class per raster position: 0 2 0 2 / 2 1 2 1 / 0 2 0 2 / 2 1 2 1 */
#pragma unroll
    for (int i = 0; i < 16; i++) {
        int c = ((i & 5) == 0) ? ls0 : ((i & 5) == 5) ? ls1 : ls2;
        d[i] = (d[i] * c) << sh;
    }
    if (chroma) { if (has_cdc) d[0] = my_cdc; }
    else if (has_ldc) d[0] = my_dc;
    idct4x4_regs(d);
    int bad = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) bad |= (d[i] < -512) | (d[i] > 511);
    if (bad) atomicOr(b.error_flags, 1u);
    lo.x = (d[0] & 0xffff) | (d[1] << 16);   lo.y = (d[2] & 0xffff) | (d[3] << 16);
    lo.z = (d[4] & 0xffff) | (d[5] << 16);   lo.w = (d[6] & 0xffff) | (d[7] << 16);
    hi.x = (d[8] & 0xffff) | (d[9] << 16);   hi.y = (d[10] & 0xffff) | (d[11] << 16);
    hi.z = (d[12] & 0xffff) | (d[13] << 16); hi.w = (d[14] & 0xffff) | (d[15] << 16);
    *reinterpret_cast<int4 *>(slot) = lo; *reinterpret_cast<int4 *>(slot + 8) = hi;
}
