/* k3_intra.cuh — kernel family 3: intra prediction + residual add as a macroblock
 * wavefront (also writes I_PCM macroblocks).
 *
 * Device replacement of h264bsdIntraPrediction (h264bsd_intra_prediction.c:
 * 475-532): neighbour fetch h264bsdGetNeighbourPels (:544-613), Intra16x16
 * (:626-686, modes :999-1148), Intra4x4 (:700-832, Get4x4NeighbourPels
 * :1387-1480, nine modes :1492-1829), chroma (:844-914, :1159-1376),
 * h264bsdAddResidual (:926-988) and h264bsdWriteMacroblock (h264bsd_image.c:
 * 80-143).  Runs AFTER K2: an intra macroblock may predict from inter
 * neighbours of the same picture (unconstrained intra), and always from
 * UNFILTERED samples (deblocking is K4).
 *
 * Parallelisation: macroblock (x,y) needs (x-1,y), (x-1,y-1), (x,y-1),
 * (x+1,y-1).  Each warp owns one macroblock ROW of one picture at a time and
 * walks its intra macroblocks left to right; before touching macroblock x it
 * spins until the row above has published progress >= min(x+2, width).  Rows
 * are handed out through an atomic ticket in (row-major, picture-minor) order,
 * so a warp only ever waits on tickets lower than its own, which are held by
 * resident warps: forward progress needs no co-residency guarantee beyond
 * that.  Hand-over is st.release.gpu on the row's progress counter and acquire
 * polls with back-off on the row above (k_common.cuh); samples produced by other
 * warps are read with ld.global.cg (L2), never through the non-coherent L1.
 * Inside a macroblock: I16x16/chroma -> 8 / 4 samples per lane; I4x4 -> the
 * sixteen 4x4 blocks in decoding order, 16 lanes each, through a shared-memory
 * tile that also holds the neighbour row/column.
 */
#pragma once
#include "k_common.cuh"

#define K3_WARPS 4
#define K3_PUBLISH 2                /* macroblocks per progress hand-over (the release is a memory barrier) */
#define K3_TP 48                    /* tile pitch (multiple of 16): interior at columns 16..31, up-right 32..35, left column 15 */

struct __align__(16) K3Warp {
    h264b200_mb_t rec;
    __align__(16) uint8_t tile[17][K3_TP];       /* row 0 = samples above the macroblock */
    __align__(16) uint8_t ctile[2][9][24];       /* chroma: interior at columns 8..15, left column 7 */
    uint8_t i4taps[9][16];                       /* H264_I4_TAPS copied to shared memory: indexed by lane, constant memory would serialise */
    __align__(16) int16_t res[26][16];           /* the macroblock's residual slots, staged once (the I4x4 chain must not wait on HBM 16 times) */
};

__device__ __forceinline__ uint8_t ldcg_u8(const uint8_t *p) { return __ldcg(p); }

/* plane prediction parameters from a neighbour row/column held in shared memory.
 * top(i)/left(i) for i in -1..n-1.  Returns a, b, c of 8.3.3.4 / 8.3.4.4. */
template <int N, typename FT, typename FL>
__device__ __forceinline__ void plane_params(FT top, FL left, int &a, int &bb, int &cc)
{
    const int half = N / 2;
    int hh = 0, vv = 0;
#pragma unroll
    for (int i = 0; i < half; i++) {
        hh += (i + 1) * (top(half + i) - top(half - 2 - i));
        vv += (i + 1) * (left(half + i) - left(half - 2 - i));
    }
    a = 16 * (left(N - 1) + top(N - 1));
    if (N == 16) { bb = (5 * hh + 32) >> 6; cc = (5 * vv + 32) >> 6; }
    else { bb = (34 * hh + 32) >> 6; cc = (34 * vv + 32) >> 6; }
}

/* recv: the macroblock's record, one 16-byte piece in each of lanes 0..7 (fetched by the caller before it waited) */
__device__ void k3_macroblock(const PicJob &job, K3Warp &w, int mbx, int mby, int lane, int4 recv)
{
    const int W = job.wm * 16, H = job.hm * 16, CW = W >> 1;
    const size_t ysize = (size_t)W * H, csize = ysize >> 2;
    if (lane < 8) reinterpret_cast<int4 *>(&w.rec)[lane] = recv;
    __syncwarp();
    const int cls = w.rec.mb_class;
    uint8_t *Y = job.cur + (size_t)mby * 16 * W + mbx * 16;
    const int16_t *coef = job.coef + (size_t)w.rec.coef_offset * 16;

    if (cls == H264B200_MB_IPCM) {                     /* raw samples sit in the INPUT slots (K1 never touches them) */
        const uint8_t *src = reinterpret_cast<const uint8_t *>(job.coef_in + (size_t)w.rec.coef_offset * 16);
        if (lane < 16) *reinterpret_cast<int4 *>(Y + (size_t)lane * W) = __ldg(reinterpret_cast<const int4 *>(src) + lane);
        else {
            int pl = (lane - 16) >> 3, r = lane & 7;
            *reinterpret_cast<int2 *>(job.cur + ysize + (pl ? csize : 0) + (size_t)(mby * 8 + r) * CW + mbx * 8) =
                __ldg(reinterpret_cast<const int2 *>(src + 256 + 64 * pl) + r);
        }
        return;
    }
    const uint32_t mask = w.rec.resid_mask;
    /* residual slots of this macroblock (K1 output) -> registers now, shared memory below: one exposed latency */
    const int n_vec = 2 * __popc(mask & 0x3ffffffu);                /* 16-byte vectors: two per slot, DC slots included */
    int4 rv0 = make_int4(0, 0, 0, 0), rv1 = rv0;
    if (lane < n_vec) rv0 = __ldg(reinterpret_cast<const int4 *>(coef) + lane);
    if (lane + 32 < n_vec) rv1 = __ldg(reinterpret_cast<const int4 *>(coef) + lane + 32);
    const bool aA = w.rec.avail & H264B200_AVAIL_A, aB = w.rec.avail & H264B200_AVAIL_B;
    const bool aC = w.rec.avail & H264B200_AVAIL_C, aD = w.rec.avail & H264B200_AVAIL_D;

    /* ---- neighbour samples into the tile: row 0 (corner, 16 above, 4 above-right), column 15 ---- */
    {
        int c = lane;                                  /* columns -1..19 relative to the macroblock: 21 samples */
        if (c < 21) {
            int rel = c - 1;
            bool ok = rel < 0 ? aD : rel < 16 ? aB : aC;
            w.tile[0][16 + rel] = ok ? ldcg_u8(Y - W + rel) : 0;
        }
        if (lane < 16) w.tile[1 + lane][15] = aA ? ldcg_u8(Y + (size_t)lane * W - 1) : 0;
    }
    if (lane < n_vec) reinterpret_cast<int4 *>(&w.res[0][0])[lane] = rv0;
    if (lane + 32 < n_vec) reinterpret_cast<int4 *>(&w.res[0][0])[lane + 32] = rv1;
    __syncwarp();

    if (cls == H264B200_MB_I16x16) {
        const int mode = w.rec.i16_mode;
        int dc = 128, pa = 0, pb = 0, pc = 0;
        if (mode == 2) {
            int st = lane < 16 ? w.tile[0][16 + lane] : 0, sl = lane < 16 ? w.tile[1 + lane][15] : 0;
#pragma unroll
            for (int o = 16; o; o >>= 1) { st += __shfl_xor_sync(0xffffffffu, st, o); sl += __shfl_xor_sync(0xffffffffu, sl, o); }
            if (aA && aB) dc = (st + sl + 16) >> 5; else if (aA) dc = (sl + 8) >> 4; else if (aB) dc = (st + 8) >> 4;
        } else if (mode == 3) {
            plane_params<16>([&](int i) { return (int)w.tile[0][16 + i]; }, [&](int i) { return (int)w.tile[1 + i][15]; }, pa, pb, pc);
        }
        /* lane -> row lane>>1, columns (lane&1)*8 .. +7 */
        const int y = lane >> 1, x0 = (lane & 1) * 8;
        uint32_t outw[2];
#pragma unroll
        for (int hq = 0; hq < 2; hq++) {
            const int bx4 = (x0 >> 2) + hq, by4 = y >> 2;
            const int bi = (bx4 & 1) | ((by4 & 1) << 1) | ((bx4 & 2) << 1) | ((by4 & 2) << 2);     /* luma4x4BlkIdx */
            const bool has_r = (mask >> bi) & 1;
            const int16_t *rs = &w.res[slot_index(mask, bi)][(y & 3) * 4];
            uint32_t pk = 0;
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int x = x0 + hq * 4 + i;
                int v;
                if (mode == 0) v = w.tile[0][16 + x];
                else if (mode == 1) v = w.tile[1 + y][15];
                else if (mode == 2) v = dc;
                else v = clip255((pa + pb * (x - 7) + pc * (y - 7) + 16) >> 5);
                if (has_r) v = clip255(v + rs[i]);
                pk |= (uint32_t)v << (8 * i);
            }
            outw[hq] = pk;
        }
        *reinterpret_cast<uint2 *>(Y + (size_t)y * W + x0) = make_uint2(outw[0], outw[1]);
    } else {
        /* ---- Intra4x4: 16 blocks in decoding order through the tile.  Every sample of the eight directional modes
         * is a 1/2/3-tap filter over consecutive entries of one edge array E = [L3 L3 L2 L1 L0 TL T0..T7 T7]
         * (H264_I4_TAPS, generated by tools/gen_i4_tables.py): no per-mode code, no divergence inside a block.
         * A block only reads samples outside itself, so one warp barrier per block is enough. ---- */
        /* everything that does not depend on reconstructed samples is fetched before the serial chain: per block this
         * lane's tap descriptor and residual value (the chain then is: 3 tile loads -> filter -> store -> barrier) */
        const int l16 = lane & 15;
        int tp_all[16], rs_all[16];
#pragma unroll
        for (int blk = 0; blk < 16; blk++) {
            const int mode = w.rec.i4_mode[blk];
            tp_all[blk] = mode == 2 ? 0 : w.i4taps[mode][l16];
            rs_all[blk] = ((mask >> blk) & 1) ? (int)w.res[slot_index(mask, blk)][l16] : 0;
        }
#pragma unroll
        for (int blk = 0; blk < 16; blk++) {
            const int x4 = (blk & 1) | ((blk >> 1) & 2), y4 = ((blk >> 1) & 1) | ((blk >> 2) & 2);
            const bool has_left = x4 > 0 || aA, has_top = y4 > 0 || aB;
            bool has_ur;
            if (y4 == 0) has_ur = x4 < 3 ? aB : aC;
            else has_ur = (0x5744u >> blk) & 1;        /* above-right block already decoded inside this macroblock */
            const int tx = 16 + 4 * x4, ty = 1 + 4 * y4;
            if (lane < 16) {
                const int tp = tp_all[blk];
                int v;
                if (tp == 0) {                         /* DC (8.3.1.2.3) */
                    const uint8_t *top = &w.tile[ty - 1][tx];
                    const int st = top[0] + top[1] + top[2] + top[3];
                    const int sl = w.tile[ty][tx - 1] + w.tile[ty + 1][tx - 1] + w.tile[ty + 2][tx - 1] + w.tile[ty + 3][tx - 1];
                    v = has_top && has_left ? (st + sl + 4) >> 3 : has_left ? (sl + 2) >> 2 : has_top ? (st + 2) >> 2 : 128;
                } else {
                    const int i0 = tp & 15, n = tp >> 4;
                    int e[3];
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const int i = min(i0 + k, 14);
                        if (i >= 5) {                  /* TL, T0..T7 (T4..T7 repeat T3 when the block above-right is not available) */
                            int c = min(i - 6, 7);
                            if (c > 3 && !has_ur) c = 3;
                            e[k] = w.tile[ty - 1][tx + c];
                        } else e[k] = w.tile[ty + min(3, 4 - i)][tx - 1];
                    }
                    v = n == 1 ? e[0] : n == 2 ? (e[0] + e[1] + 1) >> 1 : (e[0] + 2 * e[1] + e[2] + 2) >> 2;
                }
                if ((mask >> blk) & 1) v = clip255(v + rs_all[blk]);
                w.tile[ty + (lane >> 2)][tx + (lane & 3)] = (uint8_t)v;
            }
            __syncwarp();
        }
        if (lane < 16) *reinterpret_cast<int4 *>(Y + (size_t)lane * W) = *reinterpret_cast<const int4 *>(&w.tile[1 + lane][16]);
    }

    /* ---- chroma 8x8 x 2 ---- */
    {
        const int pl = lane >> 4, l16 = lane & 15;
        uint8_t *C = job.cur + ysize + (pl ? csize : 0) + (size_t)mby * 8 * CW + mbx * 8;
        if (l16 < 9) { int rel = l16 - 1; bool ok = rel < 0 ? aD : aB; w.ctile[pl][0][8 + rel] = ok ? ldcg_u8(C - CW + rel) : 0; }
        if (l16 < 8) w.ctile[pl][1 + l16][7] = aA ? ldcg_u8(C + (size_t)l16 * CW - 1) : 0;
        __syncwarp();
        const int mode = w.rec.chroma_mode;
        const int y = l16 >> 1, x0 = (l16 & 1) * 4;                 /* lane -> one 4-sample row segment */
        const uint8_t (*ct)[24] = w.ctile[pl];
        int dc = 128, pa = 0, pb = 0, pc = 0;
        if (mode == 0) {
            const int xo = x0, yo = y & 4;
            int st = ct[0][8 + xo] + ct[0][9 + xo] + ct[0][10 + xo] + ct[0][11 + xo];
            int sl = ct[1 + yo][7] + ct[2 + yo][7] + ct[3 + yo][7] + ct[4 + yo][7];
            const bool diag = (xo == 0) == (yo == 0);              /* blocks 0 and 3 */
            if (diag) { if (aA && aB) dc = (st + sl + 4) >> 3; else if (aB) dc = (st + 2) >> 2; else if (aA) dc = (sl + 2) >> 2; }
            else if (yo == 0) { if (aB) dc = (st + 2) >> 2; else if (aA) dc = (sl + 2) >> 2; }   /* block 1 */
            else { if (aA) dc = (sl + 2) >> 2; else if (aB) dc = (st + 2) >> 2; }                 /* block 2 */
        } else if (mode == 3) {
            plane_params<8>([&](int i) { return (int)ct[0][8 + i]; }, [&](int i) { return (int)ct[1 + i][7]; }, pa, pb, pc);
        }
        const int cb = 16 + 4 * pl + (y >> 2) * 2 + (x0 >> 2);
        const bool has_r = (mask >> cb) & 1;
        const int16_t *rs = &w.res[slot_index(mask, cb)][(y & 3) * 4];
        uint32_t pk = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int x = x0 + i;
            int v;
            if (mode == 0) v = dc;
            else if (mode == 1) v = ct[1 + y][7];
            else if (mode == 2) v = ct[0][8 + x];
            else v = clip255((pa + pb * (x - 3) + pc * (y - 3) + 16) >> 5);
            if (has_r) v = clip255(v + rs[i]);
            pk |= (uint32_t)v << (8 * i);
        }
        *reinterpret_cast<uint32_t *>(C + (size_t)y * CW + x0) = pk;
    }
    __syncwarp();
}

/* Occupancy: the kernel is latency bound (ncu: issue slots 21 % busy), so resident warps are throughput up to the point
 * where spills lengthen the serial Intra4x4 chain.  Measured on B200, 256 all-intra 1080p pictures per launch:
 * 6 / 8 / 10 / 12 / 16 CTAs per SM (80 / 64 / 48 / 40 / 32 registers) = 3.88 / 3.51 / 3.72 / 4.22 / 4.65 ms. */
#ifndef K3_MINB
#define K3_MINB 8
#endif
__global__ void __launch_bounds__(K3_WARPS * 32, K3_MINB) k3_intra(Batch b)
{
    __shared__ K3Warp sm[K3_WARPS];
    const int lane = threadIdx.x & 31;
    K3Warp &w = sm[threadIdx.x >> 5];
    for (int i = lane; i < 9 * 16; i += 32) (&w.i4taps[0][0])[i] = H264_I4_TAPS[i >> 4][i & 15];
    __syncwarp();
    const uint32_t n_tasks = (uint32_t)b.n_jobs * (uint32_t)b.max_hm;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&b.tickets[0], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        const int row = t / b.n_jobs;
        const PicJob &job = b.jobs[t - (uint32_t)row * b.n_jobs];
        if (row >= job.hm || job.n_intra == 0) continue;
        const int wm = job.wm;
        const h264b200_mb_t *rowrec = job.mbs + (size_t)row * wm;
        const int32_t *above = job.progress + row - 1;
        int seen = row > 0 ? 0 : 0x7fffffff, published = 0;
        for (int x0 = 0; x0 < wm; x0 += 32) {
            const int cls = (x0 + lane < wm) ? __ldg(reinterpret_cast<const uint8_t *>(rowrec + x0 + lane)) : 0;
            unsigned m = __ballot_sync(0xffffffffu, cls == H264B200_MB_I4x4 || cls == H264B200_MB_I16x16 || cls == H264B200_MB_IPCM);
            while (m) {
                const int x = x0 + __ffs(m) - 1;
                m &= m - 1;
                /* the record is requested before the hand-over so that its latency overlaps the release and the wait */
                int4 recv = make_int4(0, 0, 0, 0);
                if (lane < 8) recv = __ldg(reinterpret_cast<const int4 *>(rowrec + x) + lane);
                if (x - published >= K3_PUBLISH) { wf_publish2(job.progress + row, x, lane); published = x; }   /* everything left of x is final */
                wf_wait2(above, min(x + 2, wm), seen, lane);
                k3_macroblock(job, w, x, row, lane, recv);
            }
        }
        wf_publish2(job.progress + row, wm, lane);
    }
}
