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
 * a 2:1 wavefront.
 *
 * What bounds this kernel once enough pictures are in flight is instruction
 * issue (ncu, profiles/), so the design minimises warp instructions per
 * macroblock:
 *   - one warp owns a PAIR of macroblock rows of one picture: lanes 0..15 walk
 *     row 2p at macroblock i, lanes 16..31 walk row 2p+1 at macroblock i-2 — the
 *     wavefront lag — in lock step.  Every instruction of the walk serves two
 *     macroblocks, and the hand-over between the two rows costs nothing: row
 *     2p+1 takes the samples above it from a two-slot ring in shared memory that
 *     row 2p fills, not from global memory, and never polls.  Only row 2p waits
 *     on another warp (the pair above), through the progress counter of row
 *     2p-1 (st.release / acquire poll with back-off, ld.global.cg samples);
 *   - the edge filter runs on TWO lines per lane, one per 16-bit half of a
 *     register (k4_simd.cuh: VABSDIFF4, VIMNMX.S16x2, PRMT sign masks, biased
 *     32-bit arithmetic): of the 16 lanes of a row, 8 hold two luma lines each
 *     (20 samples: all four edges of a direction in registers, one shared-memory
 *     round trip per direction) and 8 hold two chroma lines each (Cb, Cr);
 *   - the walk is software pipelined: while macroblock x is filtered, the
 *     record, the 16x16 + 2x8x8 samples and (row 2p) the rows above of x+1 are
 *     already in flight into registers.
 * Pairs are handed out by an atomic ticket (pair-major / picture-minor), so a
 * warp only ever waits on lower tickets = resident or finished warps.
 *
 * Per macroblock and row half the 16 lanes
 *   1. derive the 32 boundary strengths, two per lane;
 *   2. filter vertical edges, then horizontal edges, in the shared window;
 *   3. write the window back (16-byte stores) and publish progress (odd rows).
 * HBM per macroblock: 384 B read + 384 B written + 128 B record (neighbour
 * records and the sample halos are L2 / shared-memory hits).
 */
#pragma once
#include "k_common.cuh"
#include "k4_simd.cuh"

#define K4_WARPS 4
#define K4_LP 20             /* luma window pitch: 5 words, conflict-free for two lines per lane */
#define K4_CP 12
#ifndef K4_PUBLISH
#define K4_PUBLISH 2         /* macroblocks per progress hand-over (3 and 4 measured the same: 2.68 / 2.70 ms against 2.67) */
#endif

/* the window of one row half: the macroblock plus 4 samples to the left and 4 (2 chroma) rows above */
/* Bank layout (found by enumerating paddings against the access patterns of the two passes, the commit and the
 * write-back, all 32 lanes = both halves at once): luma at bank 0, chroma one word later with its planes 40 words
 * apart, the two halves 16 banks apart.  47 shared-memory wavefronts for the 38 access instructions of a step
 * where the unpadded layout needed 96. */
struct K4Chroma { __align__(4) uint8_t r[12][K4_CP]; uint8_t pad[16]; };
struct __align__(16) K4Half {
    h264b200_mb_t rec[2];    /* by macroblock column parity: rec[x & 1] current, the other one left */
    h264b200_mb_t top;
    __align__(4) uint8_t y[20][K4_LP];      /* rows/cols 0..3: samples above / left of the macroblock */
    uint32_t pad0;
    K4Chroma c[2];
    uint8_t bs[2][4][4];     /* [dir][edge][segment] */
    uint32_t thr[2][3];      /* [luma/chroma][left, top, inner]: alpha | beta << 8 */
    uint32_t tc0[2][3];      /* tc0(bS=1) | tc0(bS=2) << 8 | tc0(bS=3) << 16 */
    uint32_t pad1[7];
};
static_assert(sizeof(K4Half) % 128 == 64, "the two row halves must sit 16 banks apart");
struct __align__(16) K4Pair {
    K4Half h[2];
    uint32_t ring[2][32];    /* [macroblock parity]: window rows 16..19 (20 words), chroma rows 10..11 (12 words) of the upper row */
};

/* everything of the NEXT macroblock that can be fetched ahead, one register set per lane */
struct K4Pre { int4 rec; int4 y; int2 c; uint32_t top0, top1; };

/* Records are read as 32-bit words (h264b200_records.h): word 0 = mb_class | qp_y << 8 | qp_c << 16 | qp_dbk << 24,
 * word 1 = chroma_qp_off | dbk_flags << 8 | off_a << 16 | off_b << 24, word 5 = nz_mask | slice_id << 16,
 * word 6 = ref_slot[4], word 11 = dbk_idc | flags << 8, words 16..31 = mv[16]. */
__device__ __forceinline__ bool rec_intra(const uint32_t *m) { return (m[0] & 0xff) != H264B200_MB_INTER || ((m[11] >> 8) & H264B200_MBF_DBK_AS_INTRA); }

/* bS between 4x4 block rp of macroblock p and block rq of macroblock q (raster indices); straight-line code */
__device__ __forceinline__ int dbk_bs(const uint32_t *p, int rp, const uint32_t *q, int rq, bool any_intra, bool mb_edge)
{
    const int bp = (rp & 1) | ((rp & 2) << 1) | ((rp & 4) >> 1) | (rp & 8), bq = (rq & 1) | ((rq & 2) << 1) | ((rq & 4) >> 1) | (rq & 8);
    const bool nz = ((p[5] >> bp) | (q[5] >> bq)) & 1;
    const uint32_t refp = (p[6] >> (8 * ((rp >> 3) * 2 + ((rp & 3) >> 1)))) & 0xff, refq = (q[6] >> (8 * ((rq >> 3) * 2 + ((rq & 3) >> 1)))) & 0xff;
    const uint32_t mp = p[16 + rp], mq = q[16 + rq];
    const int dx = (int)(int16_t)mp - (int)(int16_t)mq, dy = ((int)mp >> 16) - ((int)mq >> 16);
    const bool far = (unsigned)(dx + 3) > 6u || (unsigned)(dy + 3) > 6u || refp != refq;     /* abs(d) >= 4, quarter samples */
    return any_intra ? (mb_edge ? 4 : 3) : nz ? 2 : far ? 1 : 0;
}

/* geometry of one row walk (per row half) */
struct K4Row {
    const h264b200_mb_t *rowrec;
    uint8_t *Yrow, *Crow;          /* first luma / Cb sample of the macroblock row */
    int W, CW, wm, row; size_t csize;
};

/* loads that depend on no other warp: records and the macroblock's own samples.  hl = lane within the half:
 * lanes 0..7 the record, 8..15 the record above; every lane one luma row (16 B) and one chroma row (8 B) */
__device__ __forceinline__ void k4_fetch_static(const K4Row &g, int x, int hl, K4Pre &p)
{
    if (hl < 8) p.rec = __ldg(reinterpret_cast<const int4 *>(g.rowrec + x) + hl);
    else if (g.row > 0) p.rec = __ldg(reinterpret_cast<const int4 *>(g.rowrec + x - g.wm) + (hl - 8));
    p.y = __ldcg(reinterpret_cast<const int4 *>(g.Yrow + x * 16 + (size_t)hl * g.W));
    p.c = __ldcg(reinterpret_cast<const int2 *>(g.Crow + ((hl >> 3) ? g.csize : 0) + (size_t)(hl & 7) * g.CW + x * 8));
}
/* samples above the macroblock from global memory (upper row of a pair): final once the row above has
 * published >= min(x+2, wm).  16 luma words (4 rows), lanes 0..7 also one of the 8 chroma words (2 x 2 rows) */
__device__ __forceinline__ void k4_fetch_top(const K4Row &g, int x, int hl, K4Pre &p)
{
    p.top0 = __ldcg(reinterpret_cast<const uint32_t *>(g.Yrow + x * 16 - (ptrdiff_t)(4 - (hl >> 2)) * g.W + (hl & 3) * 4));
    if (hl < 8) {
        const int pl = hl >> 2, r = (hl >> 1) & 1, cw = hl & 1;
        p.top1 = __ldcg(reinterpret_cast<const uint32_t *>(g.Crow + (pl ? g.csize : 0) + x * 8 - (ptrdiff_t)(2 - r) * g.CW + cw * 4));
    }
}
/* the same samples for the lower row of a pair, out of the ring the upper row filled: columns 0..11 (chroma 0..3)
 * of macroblock x are final in the strip stored after x, columns 12..15 (4..7) in the left halo of the strip after x+1 */
__device__ __forceinline__ void k4_ring_top(const K4Pair &pw, int x, int hl, K4Pre &p)
{
    const int r = hl >> 2, k = hl & 3;
    p.top0 = k < 3 ? pw.ring[x & 1][5 * r + k + 1] : pw.ring[(x + 1) & 1][5 * r];
    if (hl < 8) {
        const int plr = hl >> 1, cw = hl & 1;
        p.top1 = cw == 0 ? pw.ring[x & 1][20 + 3 * plr + 1] : pw.ring[(x + 1) & 1][20 + 3 * plr];
    }
}
/* registers -> shared window (interior columns 4.., rows 4..; top rows 0..3) */
__device__ __forceinline__ void k4_commit(K4Half &w, int x, int hl, int row, const K4Pre &p)
{
    if (hl < 8) reinterpret_cast<int4 *>(&w.rec[x & 1])[hl] = p.rec;
    else if (row > 0) reinterpret_cast<int4 *>(&w.top)[hl - 8] = p.rec;
    uint32_t *dy = reinterpret_cast<uint32_t *>(&w.y[4 + hl][4]);
    dy[0] = (uint32_t)p.y.x; dy[1] = (uint32_t)p.y.y; dy[2] = (uint32_t)p.y.z; dy[3] = (uint32_t)p.y.w;
    uint32_t *dc = reinterpret_cast<uint32_t *>(&w.c[hl >> 3].r[4 + (hl & 7)][4]);
    dc[0] = (uint32_t)p.c.x; dc[1] = (uint32_t)p.c.y;
    if (row > 0) {
        *reinterpret_cast<uint32_t *>(&w.y[hl >> 2][4 + (hl & 3) * 4]) = p.top0;
        if (hl < 8) *reinterpret_cast<uint32_t *>(&w.c[hl >> 2].r[2 + ((hl >> 1) & 1)][4 + (hl & 1) * 4]) = p.top1;
    }
}

/* Filter the macroblocks held in the two windows of the pair.  act: this lane's row half has a macroblock to
 * filter in this step.  Returns whether this lane's macroblock was changed at all (h264bsd_deblocking.c:611).
 * Control flow is warp uniform; whatever differs between the two halves is a per-lane select. */
__device__ __forceinline__ bool k4_filter(K4Half &w, int x, int hl, int half, bool act)
{
    const h264b200_mb_t &q = w.rec[x & 1], &left = w.rec[(x & 1) ^ 1], &top = w.top;
    const uint32_t *qw = reinterpret_cast<const uint32_t *>(&q), *lw = reinterpret_cast<const uint32_t *>(&left), *tw = reinterpret_cast<const uint32_t *>(&top);
    const int fl = act ? (int)((qw[1] >> 8) & 0xff) : 0;
    const bool f_left = (fl & H264B200_DBK_LEFT) && (lw[0] & 0xff) != H264B200_MB_MISSING;
    const bool f_top = (fl & H264B200_DBK_TOP) && (tw[0] & 0xff) != H264B200_MB_MISSING;
    const bool f_inner = fl & H264B200_DBK_INNER;
    const bool q_in = rec_intra(qw), l_in = rec_intra(lw), t_in = rec_intra(tw);

    /* ---- boundary strengths: lane hl = edge*4 + segment, both directions ---- */
    unsigned weak[2], strong[2];                       /* one bit per (half, edge, segment) and direction */
    {
        const int e = hl >> 2, k = hl & 3;
#pragma unroll
        for (int dir = 0; dir < 2; dir++) {
            const int rq = dir ? e * 4 + k : k * 4 + e;
            /* the block on the other side of the edge: in the left / upper macroblock for edge 0 */
            const uint32_t *pw_ = e ? qw : dir ? tw : lw;
            const int rp = e ? (dir ? rq - 4 : rq - 1) : dir ? 12 + k : k * 4 + 3;
            const bool on = e ? f_inner : dir ? f_top : f_left;
            const bool any_intra = q_in | (e ? q_in : dir ? t_in : l_in);
            const int bsv = on ? dbk_bs(pw_, rp, qw, rq, any_intra, e == 0) : 0;
            w.bs[dir][e][k] = (uint8_t)bsv;
            weak[dir] = __ballot_sync(0xffffffffu, bsv != 0 && bsv < 4); strong[dir] = __ballot_sync(0xffffffffu, bsv == 4);
        }
    }
    if (!(weak[0] | weak[1] | strong[0] | strong[1])) return false;
    const bool mine = ((weak[0] | weak[1] | strong[0] | strong[1]) >> (16 * half)) & 0xffffu;
    if (hl < 6 && mine) {                              /* thresholds: [luma/chroma][left, top, inner] */
        const int ch = hl / 3, which = hl - ch * 3;
        int qp_q = q.qp_dbk, qp_p = which == 0 ? left.qp_dbk : which == 1 ? top.qp_dbk : q.qp_dbk;
        if (ch) {                                      /* both chroma QPs use the CURRENT macroblock's offset (:1489-1515) */
            qp_q = H264_QPC[clip3i(0, 51, qp_q + q.chroma_qp_off)];
            qp_p = H264_QPC[clip3i(0, 51, qp_p + q.chroma_qp_off)];
        }
        const int av = (qp_p + qp_q + 1) >> 1;
        const int ia = clip3i(0, 51, av + q.dbk_off_a), ib = clip3i(0, 51, av + q.dbk_off_b);
        w.thr[ch][which] = (uint32_t)H264_ALPHA[ia] | ((uint32_t)H264_BETA[ib] << 8);
        w.tc0[ch][which] = (uint32_t)H264_TC0[ia][0] | ((uint32_t)H264_TC0[ia][1] << 8) | ((uint32_t)H264_TC0[ia][2] << 16);
    }
    __syncwarp();

    /* ---- vertical edges (dir 0), then horizontal edges (dir 1).  Lanes 0..7 of the half hold two luma lines of
     * 20 samples, lanes 8..15 two chroma lines of 12 (Cb: 8..11, Cr: 12..15); all edges of the lines are filtered
     * in registers.  Step e filters luma edge e (samples v[4e..4e+7]) and, for e < 2, chroma edge e, whose
     * strength is that of luma edge 2e (h264bsd_deblocking.c:1650-1735).  The two lines of a lane lie in the same
     * 4-sample (2-sample) segment, so they share bS. ---- */
    const bool luma = hl < 8;
    const int ch = luma ? 0 : 1, pl = (hl >> 2) & 1, i = luma ? hl : (hl & 3);   /* i: line pair index */
    /* the direction loop stays rolled: one copy of the four edge filters in the instruction stream instead of two
     * (the unrolled kernel was 45 KB of SASS and stalled on instruction fetch, more so with more resident warps) */
#pragma unroll 1
    for (int dir = 0; dir < 2; dir++) {
        uint32_t v[20];
        if (dir == 0) {                                /* two rows: word loads */
            const uint32_t *ra = luma ? reinterpret_cast<const uint32_t *>(&w.y[4 + 2 * i][0]) : reinterpret_cast<const uint32_t *>(&w.c[pl].r[4 + 2 * i][0]);
            const int pw = luma ? K4_LP / 4 : K4_CP / 4;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                if (k < 3 || luma) k4s_unpack_rows(ra[k], ra[pw + k], v + 4 * k);
                else { v[4 * k] = v[4 * k + 1] = v[4 * k + 2] = v[4 * k + 3] = 0; }
            }
        } else {                                       /* two columns: 16-bit loads, consecutive lanes hit consecutive halfwords */
            const uint8_t *src = luma ? &w.y[0][4 + 2 * i] : &w.c[pl].r[0][4 + 2 * i];
            const int pitch = luma ? K4_LP : K4_CP;
#pragma unroll
            for (int k = 0; k < 20; k++) v[k] = (k < 12 || luma) ? k4s_unpack_pair(*reinterpret_cast<const uint16_t *>(src + k * pitch)) : 0u;
        }
        const uint32_t thr_e0 = w.thr[ch][dir], thr_in = w.thr[ch][2], tc_e0 = w.tc0[ch][dir], tc_in = w.tc0[ch][2];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            const unsigned wl = 0x000f000fu << (e * 4), wc = e < 2 ? 0x000f000fu << (e * 8) : 0u;
            const unsigned wk = (dir ? weak[1] : weak[0]) & (wl | wc), st = (dir ? strong[1] : strong[0]) & (wl | wc);
            if (!(wk | st)) continue;                  /* warp-uniform */
            int bsv = luma ? w.bs[dir][e][i >> 1] : w.bs[dir][(2 * e) & 3][i];
            if (!luma && e >= 2) bsv = 0;
            dbk_edge2(v + 4 * e, bsv, e ? thr_in : thr_e0, e ? tc_in : tc_e0, luma, wk != 0, st != 0);
        }
        if (mine) {
            if (dir == 0) {
                uint32_t *ra = luma ? reinterpret_cast<uint32_t *>(&w.y[4 + 2 * i][0]) : reinterpret_cast<uint32_t *>(&w.c[pl].r[4 + 2 * i][0]);
                const int pw = luma ? K4_LP / 4 : K4_CP / 4;
#pragma unroll
                for (int k = 0; k < 5; k++) if (k < 3 || luma) k4s_pack_rows(v + 4 * k, &ra[k], &ra[pw + k]);
            } else {
                uint8_t *dst = luma ? &w.y[0][4 + 2 * i] : &w.c[pl].r[0][4 + 2 * i];
                const int pitch = luma ? K4_LP : K4_CP;
#pragma unroll
                for (int k = 1; k < 19; k++) if (k < 12 || luma) *reinterpret_cast<uint16_t *>(dst + k * pitch) = (uint16_t)k4s_pack_pair(v[k]);
            }
        }
        __syncwarp();
    }
    return mine;
}

/* window -> frame: rows -3..-1 x cols 0..15, rows 0..15 x cols -4..15 (chroma: row -1; rows 0..7 x cols -4..7).
 * Every lane of the half owns one luma and one chroma row (a word for the halo, a 16 / 8 byte store for the
 * macroblock) and one word of the rows above. */
__device__ __forceinline__ void k4_writeback(const K4Row &g, const K4Half &w, int x, int hl)
{
    {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(&w.y[4 + hl][0]);
        uint8_t *d = g.Yrow + x * 16 + (size_t)hl * g.W;
        if (x > 0) *reinterpret_cast<uint32_t *>(d - 4) = s[0];
        *reinterpret_cast<int4 *>(d) = make_int4((int)s[1], (int)s[2], (int)s[3], (int)s[4]);
    }
    {
        const int pl = hl >> 3, r = hl & 7;
        const uint32_t *s = reinterpret_cast<const uint32_t *>(&w.c[pl].r[4 + r][0]);
        uint8_t *d = g.Crow + (pl ? g.csize : 0) + (size_t)r * g.CW + x * 8;
        if (x > 0) *reinterpret_cast<uint32_t *>(d - 4) = s[0];
        *reinterpret_cast<int2 *>(d) = make_int2((int)s[1], (int)s[2]);
    }
    if (g.row > 0) {
        if (hl < 12) {
            const int r = hl >> 2, k = hl & 3;         /* window rows 1..3 = frame rows -3..-1 */
            *reinterpret_cast<uint32_t *>(g.Yrow + x * 16 - (ptrdiff_t)(3 - r) * g.W + k * 4) = *reinterpret_cast<const uint32_t *>(&w.y[1 + r][4 + 4 * k]);
        } else {
            const int pl = (hl >> 1) & 1, k = hl & 1;
            *reinterpret_cast<uint32_t *>(g.Crow + (pl ? g.csize : 0) + x * 8 - (ptrdiff_t)g.CW + k * 4) = *reinterpret_cast<const uint32_t *>(&w.c[pl].r[3][4 + 4 * k]);
        }
    }
}

__device__ __forceinline__ void k4_body(const Batch &b)
{
    __shared__ K4Pair sm[K4_WARPS];
    const int lane = threadIdx.x & 31, hl = lane & 15, half = lane >> 4;
    K4Pair &pw = sm[threadIdx.x >> 5];
    K4Half &w = pw.h[half];
    const uint32_t n_pairs = ((uint32_t)b.max_hm + 1) >> 1, n_tasks = (uint32_t)b.n_jobs * n_pairs;
    /* ring source of this lane: word `lane` of the strip = upper window rows 16..19 (5 words each), chroma rows 10..11 (3 words each) */
    const uint32_t *ring_src;
    if (lane < 20) ring_src = reinterpret_cast<const uint32_t *>(&pw.h[0].y[16 + lane / 5][0]) + lane % 5;
    else { const int j = lane - 20, plr = j / 3; ring_src = reinterpret_cast<const uint32_t *>(&pw.h[0].c[plr >> 1].r[10 + (plr & 1)][0]) + j % 3; }

    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(&b.tickets[1], 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tasks) break;
        const int pair = t / b.n_jobs;
        const PicJob &job = b.jobs[t - (uint32_t)pair * b.n_jobs];
        const int row0 = 2 * pair;
        if (row0 >= job.hm || !job.any_deblock) continue;
        const int wm = job.wm;
        K4Row g;
        g.wm = wm; g.row = row0 + half; g.W = wm * 16; g.CW = g.W >> 1;
        g.csize = (size_t)g.W * (job.hm * 16) >> 2;
        const bool valid = g.row < job.hm;               /* an odd picture height leaves the last pair without a lower row */
        const bool lower = job.hm > row0 + 1;            /* warp uniform */
        if (!valid) g.row = row0;                        /* keep the addresses of the idle half inside the picture */
        g.rowrec = job.mbs + (size_t)g.row * wm;
        g.Yrow = job.cur + (size_t)g.row * 16 * g.W;
        g.Crow = job.cur + (size_t)g.W * (job.hm * 16) + (size_t)g.row * 8 * g.CW;
        int32_t *prog = job.progress + job.hm;           /* second half: the first hm counters belong to K3 */
        const int32_t *above = prog + row0 - 1;
        int seen = row0 > 0 ? 0 : 0x7fffffff;

        const bool tr = b.trace && (t - (uint32_t)pair * b.n_jobs) == 0 && hl == 0 && valid;
        if (tr) b.trace[256 + g.row * 4] = gtime();
        /* prologue: macroblock 0 of the upper row straight into its window */
        K4Pre p;
        if (half == 0) k4_fetch_static(g, 0, hl, p);
        if (row0 > 0) {
            wf_wait2(above, min(2, wm), seen, lane);
            if (half == 0) k4_fetch_top(g, 0, hl, p);
        }
        __syncwarp();
        if (half == 0) k4_commit(w, 0, hl, g.row, p);
        __syncwarp();
        if (tr) b.trace[256 + g.row * 4 + 1] = gtime();

        const int steps = lower ? wm + 2 : wm;
        for (int i = 0; i < steps; i++) {
            const int x = i - 2 * half;                  /* this half's macroblock in this step */
            const bool have = valid && x >= 0 && x < wm;
            const bool next = valid && x + 1 >= 0 && x + 1 < wm;
            if (tr && x == wm / 2) b.trace[256 + g.row * 4 + 2] = gtime();
            /* ---- everything macroblock x+1 needs goes in flight now, while x is filtered ---- */
            bool top_ahead = false;
            int polled = 0;
            if (next) k4_fetch_static(g, x + 1, hl, p);
            if (row0 > 0 && i + 1 < wm) {                /* upper row: the rows above come from the pair above */
                top_ahead = seen >= min(i + 3, wm);
                if (top_ahead && half == 0) k4_fetch_top(g, i + 1, hl, p);
            }
            const bool poll = seen < wm;                 /* refresh `seen` once per step, without waiting for it here */
            if (poll && lane == 0) { polled = ld_poll(above); wf_acquired(polled > seen); }   /* an acquire: k_common.cuh; the __syncwarp()s of the step order it before the other lanes' loads */

            const bool act = have && w.rec[x & 1].dbk_flags && w.rec[x & 1].mb_class != H264B200_MB_MISSING;
            const bool f = k4_filter(w, x, hl, half, act);
            /* lower row: publish what the PREVIOUS steps completed (i - 2 macroblocks), once per K4_PUBLISH macroblocks and
             * before this step's write-back: the stores the release has to cover were issued a whole filter ago, so
             * its memory barrier returns at once instead of stalling the warp on stores still in flight */
            if (lower && i > 2 && ((i - 2) % K4_PUBLISH) == 0 && lane == 16) st_release(prog + row0 + 1, i - 2);
            if (f) k4_writeback(g, w, x, hl);
            __syncwarp();
            /* upper row -> ring: the strip of this step (left halo final, macroblock columns 0..12 final) */
            if (lower && i <= wm) pw.ring[i & 1][lane] = *ring_src;
            /* carry the right 4 columns over as the next left halo */
            uint32_t cy = 0, cc = 0;
            if (have) {
                cy = *reinterpret_cast<const uint32_t *>(&w.y[4 + hl][16]);
                cc = *reinterpret_cast<const uint32_t *>(&w.c[hl >> 3].r[4 + (hl & 7)][8]);
            }
            __syncwarp();
            if (have) {
                *reinterpret_cast<uint32_t *>(&w.y[4 + hl][0]) = cy;
                *reinterpret_cast<uint32_t *>(&w.c[hl >> 3].r[4 + (hl & 7)][0]) = cc;
            }
            if (poll) seen = max(seen, __shfl_sync(0xffffffffu, polled, 0));
            /* land the prefetched macroblock */
            if (row0 > 0 && i + 1 < wm && !top_ahead) {
                wf_wait2(above, min(i + 3, wm), seen, lane);
                if (half == 0) k4_fetch_top(g, i + 1, hl, p);
            }
            if (next) {
                if (half == 1) k4_ring_top(pw, x + 1, hl, p);
                k4_commit(w, x + 1, hl, g.row, p);
            }
            __syncwarp();
        }
        if (lower && lane == 16) st_release(prog + row0 + 1, wm);      /* the loop ended on a __syncwarp: every lane's stores are ordered before */
        if (tr) b.trace[256 + g.row * 4 + 3] = gtime();
    }
}

/* Occupancy: 4 CTAs (16 warps) per SM at 127 registers, the whole register file.  Measured on B200 per 256 pictures:
 * 2 / 3 / 4 CTAs per SM at 127 registers (limited with dynamic shared memory) 3.81 / 3.24 / 2.68 ms — more warps
 * help — but capping registers for 5 / 6 CTAs (96 / 80 registers) gives 2.86 / 3.18 ms: the 96-register code is 14 % slower
 * at equal occupancy (3.04 ms at 4 CTAs, no spills: less room for ptxas to overlap the dependent chains), which the
 * fifth CTA (+5 %) does not win back. */
__global__ void __launch_bounds__(K4_WARPS * 32, 4) k4_deblock(Batch b) { k4_body(b); }
