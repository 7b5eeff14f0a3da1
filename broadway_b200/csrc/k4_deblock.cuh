/* k4_deblock.cuh — kernel family 4: the in-loop deblocking filter as a macroblock
 * wavefront.
 *
 * Device replacement of h264bsdFilterPicture (h264bsd_deblocking.c:574-639):
 * boundary strengths GetBoundaryStrengths / EdgeBoundaryStrength /
 * InnerBoundaryStrength (:1134-1370, :394-410, :331-354), thresholds
 * GetLumaEdgeThresholds / GetChromaEdgeThresholds (:1381-1532), and the edge
 * filters FilterLuma / FilterChroma (:1542-1736, :649-1121).  The per-macroblock
 * filtering flags (GetMbFilteringFlags :288-319) come resolved in the record.
 *
 * The filter is order dependent (macroblock raster order is normative): the
 * top edge of (x,y) reads samples of (x,y-1) that the left edge of (x+1,y-1)
 * has already modified, so (x,y) needs (x-1,y), (x,y-1) and (x+1,y-1) complete:
 * the same 2:1 wavefront as intra prediction, run with the same ticketed
 * row-per-warp scheme (k3_intra.cuh).  Per macroblock the warp
 *   1. derives the 32 boundary strengths, one per lane (2 directions x 4 edges
 *      x 4 segments) from the current/left/top records;
 *   2. pulls the 20x20 luma and two 12x12 chroma windows (macroblock + 4
 *      samples left and above) from L2 into shared memory with 32-bit loads;
 *   3. filters vertical edges then horizontal edges in shared memory: lanes
 *      0..15 each own one luma line, lanes 16..31 one chroma line (Cb rows,
 *      Cr rows), all four (two) edges of the line in sequence;
 *   4. writes the window back and publishes progress.
 * HBM per macroblock: 384 B read + 384 B written + 128 B record (neighbour
 * records and the 4-sample halos are L2 hits).
 */
#pragma once
#include "k_common.cuh"
#include "k3_intra.cuh"      /* wf_wait / wf_publish */

#define K4_WARPS 4
#define K4_LP 20             /* luma window pitch: 5 words, conflict-free for one line per lane */
#define K4_CP 12

struct __align__(16) K4Warp {
    h264b200_mb_t rec[2];    /* current / left (ping-pong) */
    h264b200_mb_t top;
    __align__(4) uint8_t y[20][K4_LP];
    __align__(4) uint8_t c[2][12][K4_CP];
    uint8_t bs[2][4][4];     /* [dir][edge][segment] */
    uint8_t alpha[2][3], beta[2][3], idxa[2][3];   /* [luma/chroma][left, top, inner] */
};

/* one line of samples across an edge; q0 at pix, neighbours at +-step */
__device__ __forceinline__ void dbk_line(uint8_t *pix, int step, int bs, int alpha, int beta, int tc0, bool luma)
{
    const int p0 = pix[-step], p1 = pix[-2 * step], q0 = pix[0], q1 = pix[step];
    if (abs(p0 - q0) >= alpha || abs(p1 - p0) >= beta || abs(q1 - q0) >= beta) return;
    if (bs < 4) {
        int tc;
        if (luma) {
            const int p2 = pix[-3 * step], q2 = pix[2 * step];
            const bool ap = abs(p2 - p0) < beta, aq = abs(q2 - q0) < beta;
            tc = tc0 + ap + aq;
            if (ap) pix[-2 * step] = (uint8_t)(p1 + clip3i(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1));
            if (aq) pix[step] = (uint8_t)(q1 + clip3i(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1));
        } else tc = tc0 + 1;
        const int d = clip3i(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        pix[-step] = (uint8_t)clip255(p0 + d);
        pix[0] = (uint8_t)clip255(q0 - d);
    } else if (luma) {
        const int p2 = pix[-3 * step], q2 = pix[2 * step], p3 = pix[-4 * step], q3 = pix[3 * step];
        const bool small = abs(p0 - q0) < ((alpha >> 2) + 2);
        if (abs(p2 - p0) < beta && small) {
            pix[-step] = (uint8_t)((p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3);
            pix[-2 * step] = (uint8_t)((p2 + p1 + p0 + q0 + 2) >> 2);
            pix[-3 * step] = (uint8_t)((2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3);
        } else pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        if (abs(q2 - q0) < beta && small) {
            pix[0] = (uint8_t)((p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3);
            pix[step] = (uint8_t)((p0 + q0 + q1 + q2 + 2) >> 2);
            pix[2 * step] = (uint8_t)((2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3);
        } else pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    } else {
        pix[-step] = (uint8_t)((2 * p1 + p0 + q1 + 2) >> 2);
        pix[0] = (uint8_t)((2 * q1 + q0 + p1 + 2) >> 2);
    }
}

__device__ __forceinline__ bool rec_intra(const h264b200_mb_t &m) { return m.mb_class != H264B200_MB_INTER; }

/* bS between 4x4 block rp of macroblock p and block rq of macroblock q (raster indices) */
__device__ __forceinline__ int dbk_bs(const h264b200_mb_t &p, int rp, const h264b200_mb_t &q, int rq, bool mb_edge)
{
    if (rec_intra(p) || rec_intra(q)) return mb_edge ? 4 : 3;
    const int bp = (rp & 1) | ((rp & 2) << 1) | ((rp & 4) >> 1) | (rp & 8), bq = (rq & 1) | ((rq & 2) << 1) | ((rq & 4) >> 1) | (rq & 8);
    if (((p.nz_mask >> bp) | (q.nz_mask >> bq)) & 1) return 2;
    if (p.ref_slot[(rp >> 3) * 2 + ((rp & 3) >> 1)] != q.ref_slot[(rq >> 3) * 2 + ((rq & 3) >> 1)]) return 1;
    if (abs(p.mv[rp][0] - q.mv[rq][0]) >= 4 || abs(p.mv[rp][1] - q.mv[rq][1]) >= 4) return 1;
    return 0;
}

__device__ void k4_macroblock(const PicJob &job, K4Warp &w, int cur, int mbx, int mby, int lane)
{
    const int W = job.wm * 16, H = job.hm * 16, CW = W >> 1;
    const size_t ysize = (size_t)W * H, csize = ysize >> 2;
    const h264b200_mb_t &q = w.rec[cur], &left = w.rec[cur ^ 1], &top = w.top;
    const int fl = q.dbk_flags;
    const bool f_left = (fl & H264B200_DBK_LEFT) && left.mb_class != H264B200_MB_MISSING;
    const bool f_top = (fl & H264B200_DBK_TOP) && top.mb_class != H264B200_MB_MISSING;
    const bool f_inner = fl & H264B200_DBK_INNER;

    /* ---- 1. boundary strengths: lane = dir*16 + edge*4 + segment ---- */
    {
        const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
        const int rq = dir ? e * 4 + k : k * 4 + e;
        int bsv = 0;
        if (e == 0) {
            if (dir ? f_top : f_left) bsv = dbk_bs(dir ? top : left, dir ? 12 + k : k * 4 + 3, q, rq, true);
        } else if (f_inner) bsv = dbk_bs(q, dir ? rq - 4 : rq - 1, q, rq, false);
        w.bs[dir][e][k] = (uint8_t)bsv;
        const unsigned any = __ballot_sync(0xffffffffu, bsv != 0);
        if (!any) return;                              /* nothing to filter (h264bsd_deblocking.c:611) */
        if (lane < 6) {                                /* thresholds: [luma/chroma][left, top, inner] */
            const int ch = lane / 3, which = lane - ch * 3;
            int qp_q = q.qp_dbk, qp_p = which == 0 ? left.qp_dbk : which == 1 ? top.qp_dbk : q.qp_dbk;
            if (ch) {                                  /* both chroma QPs use the CURRENT macroblock's offset (:1489-1515) */
                qp_q = H264_QPC[clip3i(0, 51, qp_q + q.chroma_qp_off)];
                qp_p = H264_QPC[clip3i(0, 51, qp_p + q.chroma_qp_off)];
            }
            const int av = (qp_p + qp_q + 1) >> 1;
            const int ia = clip3i(0, 51, av + q.dbk_off_a), ib = clip3i(0, 51, av + q.dbk_off_b);
            w.alpha[ch][which] = H264_ALPHA[ia]; w.beta[ch][which] = H264_BETA[ib]; w.idxa[ch][which] = (uint8_t)ia;
        }
    }

    /* ---- 2. windows from L2: luma [-4,16) x [-4,16), chroma [-4,8) x [-4,8) ---- */
    uint8_t *Y = job.cur + (size_t)mby * 16 * W + mbx * 16;
    for (int i = lane; i < 100; i += 32) {
        const int r = i / 5, cw = i - r * 5;
        uint32_t v = 0;
        if ((r >= 4 || mby > 0) && (cw >= 1 || mbx > 0))
            v = __ldcg(reinterpret_cast<const uint32_t *>(Y + (ptrdiff_t)(r - 4) * W + (cw - 1) * 4));
        *reinterpret_cast<uint32_t *>(&w.y[r][cw * 4]) = v;
    }
    uint8_t *C0 = job.cur + ysize + (size_t)mby * 8 * CW + mbx * 8;
    for (int i = lane; i < 72; i += 32) {
        const int pl = i / 36, j = i - pl * 36, r = j / 3, cw = j - r * 3;
        uint32_t v = 0;
        if ((r >= 4 || mby > 0) && (cw >= 1 || mbx > 0))
            v = __ldcg(reinterpret_cast<const uint32_t *>(C0 + (pl ? csize : 0) + (ptrdiff_t)(r - 4) * CW + (cw - 1) * 4));
        *reinterpret_cast<uint32_t *>(&w.c[pl][r][cw * 4]) = v;
    }
    __syncwarp();

    /* ---- 3. filter: vertical edges (dir 0), then horizontal edges (dir 1) ---- */
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        if (lane < 16) {
            for (int e = 0; e < 4; e++) {
                const int bsv = w.bs[dir][e][lane >> 2];
                if (!bsv) continue;
                const int which = e ? 2 : dir;
                uint8_t *pix = dir ? &w.y[4 + 4 * e][4 + lane] : &w.y[4 + lane][4 + 4 * e];
                dbk_line(pix, dir ? K4_LP : 1, bsv, w.alpha[0][which], w.beta[0][which], bsv < 4 ? H264_TC0[w.idxa[0][which]][bsv - 1] : 0, true);
            }
        } else {
            const int pl = (lane - 16) >> 3, i = lane & 7;
            for (int e = 0; e < 4; e += 2) {
                const int bsv = w.bs[dir][e][i >> 1];
                if (!bsv) continue;
                const int which = e ? 2 : dir;
                uint8_t *pix = dir ? &w.c[pl][4 + 2 * e][4 + i] : &w.c[pl][4 + i][4 + 2 * e];
                dbk_line(pix, dir ? K4_CP : 1, bsv, w.alpha[1][which], w.beta[1][which], bsv < 4 ? H264_TC0[w.idxa[1][which]][bsv - 1] : 0, false);
            }
        }
        __syncwarp();
    }

    /* ---- 4. write back: rows -3..-1 x cols 0..15, rows 0..15 x cols -4..15 ---- */
    for (int i = lane; i < 92; i += 32) {
        int r, cw;
        if (i < 12) { r = 1 + i / 4; cw = 1 + (i & 3); if (mby == 0) continue; }
        else { const int j = i - 12; r = 4 + j / 5; cw = j % 5; if (cw == 0 && mbx == 0) continue; }
        *reinterpret_cast<uint32_t *>(Y + (ptrdiff_t)(r - 4) * W + (cw - 1) * 4) = *reinterpret_cast<const uint32_t *>(&w.y[r][cw * 4]);
    }
    for (int i = lane; i < 56; i += 32) {
        const int pl = i / 28, j = i - pl * 28;
        int r, cw;
        if (j < 4) { r = 2 + (j >> 1); cw = 1 + (j & 1); if (mby == 0) continue; }      /* rows -2,-1 */
        else { const int k = j - 4; r = 4 + k / 3; cw = k % 3; if (cw == 0 && mbx == 0) continue; }
        *reinterpret_cast<uint32_t *>(C0 + (pl ? csize : 0) + (ptrdiff_t)(r - 4) * CW + (cw - 1) * 4) = *reinterpret_cast<const uint32_t *>(&w.c[pl][r][cw * 4]);
    }
}

__global__ void __launch_bounds__(K4_WARPS * 32) k4_deblock(Batch b)
{
    __shared__ K4Warp sm[K4_WARPS];
    const int lane = threadIdx.x & 31;
    K4Warp &w = sm[threadIdx.x >> 5];
    const uint32_t n_tasks = (uint32_t)b.n_jobs * (uint32_t)b.max_hm;
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&b.tickets[1], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        const int row = t / b.n_jobs;
        const PicJob &job = b.jobs[t - (uint32_t)row * b.n_jobs];
        if (row >= job.hm || !job.any_deblock) continue;
        const int wm = job.wm;
        const h264b200_mb_t *rowrec = job.mbs + (size_t)row * wm;
        int32_t *prog = job.progress + job.hm;           /* second half: the first hm counters belong to K3 */
        int cur = 0;
        for (int x = 0; x < wm; x++, cur ^= 1) {
            if (lane < 8) reinterpret_cast<int4 *>(&w.rec[cur])[lane] = __ldg(reinterpret_cast<const int4 *>(rowrec + x) + lane);
            else if (lane < 16 && row > 0) reinterpret_cast<int4 *>(&w.top)[lane - 8] = __ldg(reinterpret_cast<const int4 *>(rowrec + x - wm) + (lane - 8));
            __syncwarp();
            if (w.rec[cur].dbk_flags && w.rec[cur].mb_class != H264B200_MB_MISSING) {
                wf_wait(prog, row, min(x + 2, wm));
                k4_macroblock(job, w, cur, x, row, lane);
            }
            wf_publish(prog, row, x + 1, lane);
        }
    }
}
