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
 * a 2:1 wavefront.  One warp owns one macroblock row of one picture (rows are
 * handed out by an atomic ticket, row-major / picture-minor, so a warp only
 * waits on lower tickets = resident warps) and walks it left to right.
 *
 * The walk is SOFTWARE PIPELINED, because what bounds a wavefront is the
 * latency of one macroblock step, not bandwidth: while macroblock x is being
 * filtered out of shared memory, everything macroblock x+1 needs is already in
 * flight into registers — its record, the record above, its 16x16 + 2x8x8
 * unfiltered samples (nobody touches them before this warp does) and, when the
 * row above is far enough ahead (the usual case), the 4 (2 chroma) sample rows
 * above it.  The 4-sample column to the left is carried over in shared memory.
 * Row-to-row hand-over uses st.release.gpu / ld.acquire.gpu on a progress
 * counter (no __threadfence, no L1 invalidate); samples other warps produced
 * are read with ld.global.cg.
 * Per macroblock the warp
 *   1. derives the 32 boundary strengths, one per lane (2 directions x 4 edges
 *      x 4 segments) from the current/left/top records;
 *   2. filters vertical edges then horizontal edges in the shared window: lanes
 *      0..15 own one luma line each (4 edges in sequence), lanes 16..31 one
 *      chroma line each (Cb rows, Cr rows; 2 edges), in ONE loop so that all 32
 *      lanes work together;
 *   3. writes the window back and publishes progress.
 * HBM per macroblock: 384 B read + 384 B written + 128 B record (neighbour
 * records and the 4-sample halos are L2 hits).
 */
#pragma once
#include "k_common.cuh"

#define K4_WARPS 4
#define K4_LP 20             /* luma window pitch: 5 words, conflict-free for one line per lane */
#define K4_CP 12
#define K4_PUBLISH 2         /* macroblocks per progress hand-over */

struct __align__(16) K4Warp {
    h264b200_mb_t rec[2];    /* current / left (ping-pong) */
    h264b200_mb_t top;
    __align__(4) uint8_t y[20][K4_LP];      /* rows/cols 0..3: samples above / left of the macroblock */
    __align__(4) uint8_t c[2][12][K4_CP];
    uint8_t bs[2][4][4];     /* [dir][edge][segment] */
    uint32_t thr[2][3];      /* [luma/chroma][left, top, inner]: alpha | beta << 8 | tc0(bS=1) << 16 */
    uint32_t tc0[2][3];      /* tc0(bS=1) | tc0(bS=2) << 8 | tc0(bS=3) << 16 */
};

/* everything of the NEXT macroblock that can be fetched ahead, one register set per lane */
struct K4Pre { int4 rec; uint32_t y0, y1, c, top; };

/* One edge of one line, entirely in registers (8.7.2.3 / 8.7.2.4; h264bsd_deblocking.c:649-1121).
 * v[0..3] = p3..p0, v[4..7] = q0..q3.  Straight-line code: every per-lane decision is a select, never a
 * branch (32 lanes hold 32 different lines); only `any_weak` / `any_strong`, warp-uniform votes, skip
 * the variant no lane needs.  Chroma lanes (luma == false) only ever change p0 and q0. */
__device__ __forceinline__ void dbk_edge(int *v, int bs, uint32_t thr, uint32_t tcw, bool luma, bool any_weak, bool any_strong)
{
    const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    const int alpha = thr & 0xff, beta = (thr >> 8) & 0xff;
    const int ad = abs(p0 - q0);
    const bool on = (bs != 0) & (ad < alpha) & (abs(p1 - p0) < beta) & (abs(q1 - q0) < beta);
    const bool ap = luma & (abs(p2 - p0) < beta), aq = luma & (abs(q2 - q0) < beta);
    int n0 = p0, n1 = p1, n2 = p2, m0 = q0, m1 = q1, m2 = q2;
    if (any_weak) {
        const bool wk = on & (bs < 4);
        const int tc0 = (tcw >> ((8 * (bs - 1)) & 31)) & 0xff;
        const int tc = tc0 + (luma ? (int)ap + (int)aq : 1);
        const int d = clip3i(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        const int avg = (p0 + q0 + 1) >> 1;
        const int e1 = clip3i(-tc0, tc0, (p2 + avg - (p1 << 1)) >> 1), f1 = clip3i(-tc0, tc0, (q2 + avg - (q1 << 1)) >> 1);
        n0 = wk ? clip255(p0 + d) : p0; m0 = wk ? clip255(q0 - d) : q0;
        n1 = (wk & ap) ? p1 + e1 : p1;  m1 = (wk & aq) ? q1 + f1 : q1;
    }
    if (any_strong) {
        const bool st = on & (bs == 4);
        const bool small = ad < ((alpha >> 2) + 2);
        const bool sp = st & ap & small, sq = st & aq & small;
        const int s = p0 + q0;
        const int sp0 = (p2 + 2 * (p1 + s) + q1 + 4) >> 3, sp1 = (p2 + p1 + s + 2) >> 2, sp2 = (2 * p3 + 3 * p2 + p1 + s + 4) >> 3;
        const int sq0 = (q2 + 2 * (q1 + s) + p1 + 4) >> 3, sq1 = (q2 + q1 + s + 2) >> 2, sq2 = (2 * q3 + 3 * q2 + q1 + s + 4) >> 3;
        const int wp0 = (2 * p1 + p0 + q1 + 2) >> 2, wq0 = (2 * q1 + q0 + p1 + 2) >> 2;
        n0 = sp ? sp0 : (st ? wp0 : n0); n1 = sp ? sp1 : n1; n2 = sp ? sp2 : n2;
        m0 = sq ? sq0 : (st ? wq0 : m0); m1 = sq ? sq1 : m1; m2 = sq ? sq2 : m2;
    }
    v[1] = n2; v[2] = n1; v[3] = n0; v[4] = m0; v[5] = m1; v[6] = m2;
}

__device__ __forceinline__ bool rec_intra(const h264b200_mb_t &m) { return m.mb_class != H264B200_MB_INTER || (m.flags & H264B200_MBF_DBK_AS_INTRA); }

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

/* geometry of one row walk */
struct K4Row {
    const PicJob *job; const h264b200_mb_t *rowrec;
    uint8_t *Yrow, *Crow;          /* first luma / Cb sample of the macroblock row */
    int W, CW, wm, row; size_t csize;
};

/* loads that depend on no other warp: records and the macroblock's own samples */
__device__ __forceinline__ void k4_fetch_static(const K4Row &g, int x, int lane, K4Pre &p)
{
    if (lane < 8) p.rec = __ldg(reinterpret_cast<const int4 *>(g.rowrec + x) + lane);
    else if (lane < 16 && g.row > 0) p.rec = __ldg(reinterpret_cast<const int4 *>(g.rowrec + x - g.wm) + (lane - 8));
    const uint8_t *Y = g.Yrow + x * 16 + (size_t)(lane >> 2) * g.W + (lane & 3) * 4;
    p.y0 = __ldcg(reinterpret_cast<const uint32_t *>(Y));
    p.y1 = __ldcg(reinterpret_cast<const uint32_t *>(Y + (size_t)8 * g.W));
    const uint8_t *C = g.Crow + ((lane >> 4) ? g.csize : 0) + (size_t)((lane >> 1) & 7) * g.CW + x * 8 + (lane & 1) * 4;
    p.c = __ldcg(reinterpret_cast<const uint32_t *>(C));
}
/* samples above the macroblock: final once the row above has published >= min(x+2, wm) */
__device__ __forceinline__ void k4_fetch_top(const K4Row &g, int x, int lane, K4Pre &p)
{
    p.top = 0;
    if (g.row == 0) return;
    if (lane < 16) p.top = __ldcg(reinterpret_cast<const uint32_t *>(g.Yrow + x * 16 - (ptrdiff_t)(4 - (lane >> 2)) * g.W + (lane & 3) * 4));
    else if (lane < 24) {
        const int k = lane - 16, pl = k >> 2, r = (k >> 1) & 1, cw = k & 1;
        p.top = __ldcg(reinterpret_cast<const uint32_t *>(g.Crow + (pl ? g.csize : 0) + x * 8 - (ptrdiff_t)(2 - r) * g.CW + cw * 4));
    }
}
/* registers -> shared window (interior columns 4.., rows 4..; top rows 0..3) */
__device__ __forceinline__ void k4_commit(K4Warp &w, int cur, int lane, int row, const K4Pre &p)
{
    if (lane < 8) reinterpret_cast<int4 *>(&w.rec[cur])[lane] = p.rec;
    else if (lane < 16 && row > 0) reinterpret_cast<int4 *>(&w.top)[lane - 8] = p.rec;
    *reinterpret_cast<uint32_t *>(&w.y[4 + (lane >> 2)][4 + (lane & 3) * 4]) = p.y0;
    *reinterpret_cast<uint32_t *>(&w.y[12 + (lane >> 2)][4 + (lane & 3) * 4]) = p.y1;
    *reinterpret_cast<uint32_t *>(&w.c[lane >> 4][4 + ((lane >> 1) & 7)][4 + (lane & 1) * 4]) = p.c;
    if (lane < 16) *reinterpret_cast<uint32_t *>(&w.y[lane >> 2][4 + (lane & 3) * 4]) = p.top;
    else if (lane < 24) { const int k = lane - 16; *reinterpret_cast<uint32_t *>(&w.c[k >> 2][2 + ((k >> 1) & 1)][4 + (k & 1) * 4]) = p.top; }
}

/* filter the macroblock held in the window; returns false when nothing was filtered */
__device__ __forceinline__ bool k4_filter(K4Warp &w, int cur, int lane)
{
    const h264b200_mb_t &q = w.rec[cur], &left = w.rec[cur ^ 1], &top = w.top;
    const int fl = q.dbk_flags;
    const bool f_left = (fl & H264B200_DBK_LEFT) && left.mb_class != H264B200_MB_MISSING;
    const bool f_top = (fl & H264B200_DBK_TOP) && top.mb_class != H264B200_MB_MISSING;
    const bool f_inner = fl & H264B200_DBK_INNER;

    /* ---- boundary strengths: lane = dir*16 + edge*4 + segment ---- */
    unsigned weak_mask, strong_mask;                   /* one bit per (dir, edge, segment): which edges need which filter variant */
    {
        const int dir = lane >> 4, e = (lane >> 2) & 3, k = lane & 3;
        const int rq = dir ? e * 4 + k : k * 4 + e;
        int bsv = 0;
        if (e == 0) {
            if (dir ? f_top : f_left) bsv = dbk_bs(dir ? top : left, dir ? 12 + k : k * 4 + 3, q, rq, true);
        } else if (f_inner) bsv = dbk_bs(q, dir ? rq - 4 : rq - 1, q, rq, false);
        w.bs[dir][e][k] = (uint8_t)bsv;
        weak_mask = __ballot_sync(0xffffffffu, bsv != 0 && bsv < 4); strong_mask = __ballot_sync(0xffffffffu, bsv == 4);
        if (!(weak_mask | strong_mask)) return false;  /* h264bsd_deblocking.c:611 */
        if (lane < 6) {                                /* thresholds: [luma/chroma][left, top, inner] */
            const int ch = lane / 3, which = lane - ch * 3;
            int qp_q = q.qp_dbk, qp_p = which == 0 ? left.qp_dbk : which == 1 ? top.qp_dbk : q.qp_dbk;
            if (ch) {                                  /* both chroma QPs use the CURRENT macroblock's offset (:1489-1515) */
                qp_q = H264_QPC[clip3i(0, 51, qp_q + q.chroma_qp_off)];
                qp_p = H264_QPC[clip3i(0, 51, qp_p + q.chroma_qp_off)];
            }
            const int av = (qp_p + qp_q + 1) >> 1;
            const int ia = clip3i(0, 51, av + q.dbk_off_a), ib = clip3i(0, 51, av + q.dbk_off_b);
            w.thr[ch][which] = (uint32_t)H264_ALPHA[ia] | ((uint32_t)H264_BETA[ib] << 8);
            w.tc0[ch][which] = (uint32_t)H264_TC0[ia][0] | ((uint32_t)H264_TC0[ia][1] << 8) | ((uint32_t)H264_TC0[ia][2] << 16);
        }
    }
    __syncwarp();

    /* ---- vertical edges (dir 0), then horizontal edges (dir 1).  Lanes 0..15 hold one luma line of 20
     * samples in registers, lanes 16..31 one chroma line of 12 (Cb lines, then Cr lines); all edges of the
     * line are filtered in registers, one shared-memory round trip per direction.  Loop step e filters luma
     * edge e (samples v[4e..4e+7]) and, for e < 2, chroma edge e (same registers), whose strength is that of
     * luma edge 2e (h264bsd_deblocking.c:1650-1735). ---- */
    const bool luma = lane < 16;
    const int ch = luma ? 0 : 1, pl = (lane >> 3) & 1, i = luma ? lane : (lane & 7);
#pragma unroll
    for (int dir = 0; dir < 2; dir++) {
        int v[20];
        if (dir == 0) {                                /* a row: word loads */
            const uint32_t *src = luma ? reinterpret_cast<const uint32_t *>(&w.y[4 + i][0]) : reinterpret_cast<const uint32_t *>(&w.c[pl][4 + i][0]);
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const uint32_t wd = (k < 3 || luma) ? src[k] : 0u;
                v[4 * k] = wd & 0xff; v[4 * k + 1] = (wd >> 8) & 0xff; v[4 * k + 2] = (wd >> 16) & 0xff; v[4 * k + 3] = wd >> 24;
            }
        } else {                                       /* a column: byte loads, consecutive lanes hit consecutive bytes */
            const uint8_t *src = luma ? &w.y[0][4 + i] : &w.c[pl][0][4 + i];
            const int pitch = luma ? K4_LP : K4_CP;
#pragma unroll
            for (int k = 0; k < 20; k++) v[k] = (k < 12 || luma) ? src[k * pitch] : 0;
        }
        const uint32_t thr_e0 = w.thr[ch][dir], thr_in = w.thr[ch][2], tc_e0 = w.tc0[ch][dir], tc_in = w.tc0[ch][2];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const unsigned wl = 0xfu << (dir * 16 + e * 4), wc = e < 2 ? 0xfu << (dir * 16 + e * 8) : 0u;
            const unsigned weak = weak_mask & (wl | wc), strong = strong_mask & (wl | wc);
            if (!(weak | strong)) continue;            /* warp-uniform */
            int bsv = luma ? w.bs[dir][e][i >> 2] : w.bs[dir][(2 * e) & 3][i >> 1];
            if (!luma && e >= 2) bsv = 0;
            dbk_edge(v + 4 * e, bsv, e ? thr_in : thr_e0, e ? tc_in : tc_e0, luma, weak != 0, strong != 0);
        }
        if (dir == 0) {
            uint32_t *dst = luma ? reinterpret_cast<uint32_t *>(&w.y[4 + i][0]) : reinterpret_cast<uint32_t *>(&w.c[pl][4 + i][0]);
#pragma unroll
            for (int k = 0; k < 5; k++)
                if (k < 3 || luma) dst[k] = (uint32_t)v[4 * k] | ((uint32_t)v[4 * k + 1] << 8) | ((uint32_t)v[4 * k + 2] << 16) | ((uint32_t)v[4 * k + 3] << 24);
        } else {
            uint8_t *dst = luma ? &w.y[0][4 + i] : &w.c[pl][0][4 + i];
            const int pitch = luma ? K4_LP : K4_CP;
#pragma unroll
            for (int k = 1; k < 19; k++) if (k < 12 || luma) dst[k * pitch] = (uint8_t)v[k];
        }
        __syncwarp();
    }
    return true;
}

/* window -> frame: rows -3..-1 x cols 0..15, rows 0..15 x cols -4..15 (chroma: row -1; rows 0..7 x cols -4..7).
 * Each lane owns up to 3 luma and 2 chroma words; their window / frame offsets are fixed for a row walk. */
struct K4Wb { int ys[3], cs[2]; int yg[3], cg[2]; unsigned flags; };   /* flags: bit k on, bit 8+k top row, bit 16+k left column (k 0..2 luma, 3..4 chroma) */

__device__ __forceinline__ void k4_wb_init(const K4Row &g, int lane, K4Wb &t)
{
    t.flags = 0;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const int i = lane + 32 * k;
        int r = 0, cw = 0;
        if (i < 92) {
            t.flags |= 1u << k;
            if (i < 12) { r = 1 + i / 4; cw = 1 + (i & 3); t.flags |= 1u << (8 + k); }
            else { const int j = i - 12; r = 4 + j / 5; cw = j % 5; if (cw == 0) t.flags |= 1u << (16 + k); }
        }
        t.ys[k] = r * K4_LP + cw * 4; t.yg[k] = (r - 4) * g.W + (cw - 1) * 4;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const int i = lane + 32 * k;
        int pl = 0, r = 0, cw = 0;
        if (i < 52) {
            t.flags |= 1u << (3 + k);
            pl = i / 26; const int j = i - pl * 26;
            if (j < 2) { r = 3; cw = 1 + j; t.flags |= 1u << (11 + k); }
            else { const int q = j - 2; r = 4 + q / 3; cw = q % 3; if (cw == 0) t.flags |= 1u << (19 + k); }
        }
        t.cs[k] = (pl * 12 + r) * K4_CP + cw * 4; t.cg[k] = (pl ? (int)g.csize : 0) + (r - 4) * g.CW + (cw - 1) * 4;
    }
}

__device__ __forceinline__ void k4_writeback(const K4Row &g, K4Warp &w, const K4Wb &t, int x)
{
    uint8_t *Y = g.Yrow + x * 16, *C0 = g.Crow + x * 8;
    const uint8_t *ys = &w.y[0][0], *cs = &w.c[0][0][0];
    unsigned on = t.flags & 0xff;
    if (g.row == 0) on &= ~(t.flags >> 8);
    if (x == 0) on &= ~(t.flags >> 16);
#pragma unroll
    for (int k = 0; k < 3; k++) if ((on >> k) & 1) *reinterpret_cast<uint32_t *>(Y + t.yg[k]) = *reinterpret_cast<const uint32_t *>(ys + t.ys[k]);
#pragma unroll
    for (int k = 0; k < 2; k++) if ((on >> (3 + k)) & 1) *reinterpret_cast<uint32_t *>(C0 + t.cg[k]) = *reinterpret_cast<const uint32_t *>(cs + t.cs[k]);
}

__global__ void __launch_bounds__(K4_WARPS * 32, 6) k4_deblock(Batch b)
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
        K4Row g;
        g.job = &job; g.wm = job.wm; g.row = row; g.W = job.wm * 16; g.CW = g.W >> 1;
        g.csize = (size_t)g.W * (job.hm * 16) >> 2;
        g.rowrec = job.mbs + (size_t)row * g.wm;
        g.Yrow = job.cur + (size_t)row * 16 * g.W;
        g.Crow = job.cur + (size_t)g.W * (job.hm * 16) + (size_t)row * 8 * g.CW;
        int32_t *prog = job.progress + job.hm;           /* second half: the first hm counters belong to K3 */
        const int32_t *above = prog + row - 1;
        int seen = row > 0 ? 0 : 0x7fffffff;
        const int wm = g.wm;
        K4Wb wb;
        k4_wb_init(g, lane, wb);

        const bool tr = b.trace && (t - (uint32_t)row * b.n_jobs) == 0 && lane == 0;
        if (tr) b.trace[256 + row * 4] = gtime();
        /* prologue: macroblock 0 straight into the window */
        K4Pre p;
        k4_fetch_static(g, 0, lane, p);
        wf_wait2(above, min(2, wm), seen, lane);
        k4_fetch_top(g, 0, lane, p);
        __syncwarp();
        k4_commit(w, 0, lane, row, p);
        int cur = 0;
        if (tr) b.trace[256 + row * 4 + 1] = gtime();
        for (int x = 0; x < wm; x++, cur ^= 1) {
            if (tr && x == wm / 2) b.trace[256 + row * 4 + 2] = gtime();
            /* ---- everything macroblock x+1 needs goes in flight now, while x is filtered ---- */
            const bool more = x + 1 < wm;
            bool top_ahead = false;
            int polled = 0;
            if (more) {
                k4_fetch_static(g, x + 1, lane, p);
                top_ahead = seen >= min(x + 3, wm);
                if (top_ahead) k4_fetch_top(g, x + 1, lane, p);
            }
            const bool poll = seen < wm;                   /* refresh `seen` once per macroblock, without waiting for it here */
            if (poll && lane == 0) polled = ld_acquire(above);
            __syncwarp();
            const bool active = w.rec[cur].dbk_flags && w.rec[cur].mb_class != H264B200_MB_MISSING;
            if (active) {
                const int need = min(x + 2, wm);
                if (seen < need) {
                    seen = __shfl_sync(0xffffffffu, polled, 0);
                    wf_wait2(above, need, seen, lane);        /* only a row running right at the wavefront spins here */
                }
                const bool f = k4_filter(w, cur, lane);
                    if (f) k4_writeback(g, w, wb, x);
                }
            /* the release (a memory barrier over the write-back) is paid once per K4_PUBLISH macroblocks */
            if (((x + 1) % K4_PUBLISH) == 0 || !more) wf_publish2(prog + row, x + 1, lane);
            if (more) {
                /* carry the right 4 columns over as the next left halo, then land the prefetched macroblock */
                uint32_t carry;
                if (lane < 16) carry = *reinterpret_cast<const uint32_t *>(&w.y[4 + lane][16]);
                else carry = *reinterpret_cast<const uint32_t *>(&w.c[(lane >> 3) & 1][4 + (lane & 7)][8]);
                __syncwarp();
                if (lane < 16) *reinterpret_cast<uint32_t *>(&w.y[4 + lane][0]) = carry;
                else *reinterpret_cast<uint32_t *>(&w.c[(lane >> 3) & 1][4 + (lane & 7)][0]) = carry;
                if (poll) seen = max(seen, __shfl_sync(0xffffffffu, polled, 0));
                if (!top_ahead) { wf_wait2(above, min(x + 3, wm), seen, lane); k4_fetch_top(g, x + 1, lane, p); }
                k4_commit(w, cur ^ 1, lane, row, p);
            }
            __syncwarp();
        }
        if (tr) b.trace[256 + row * 4 + 3] = gtime();
    }
}
