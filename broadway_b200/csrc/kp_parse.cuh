/* kp_parse.cuh — kernel Kp: slice data of MANY pictures parsed on the device, one warp per picture (kp_core.h), and
 * k0_jobs, which hands Kp's per-picture results (class counts, concealment list, status words) to the job table the
 * reconstruction kernels K1..K4 read.
 *
 * Launch shape: persistent CTAs; a warp takes pictures from an atomic ticket until none is left, so a
 * launch over more pictures than resident warps stays balanced.  Per CTA the host parser's look-up tables (25 KB,
 * KpTables) are copied into shared memory once; per warp 2.9 KB of staging (KpStage).  Only lane 0 walks the syntax —
 * the bitstream is serial — so the kernel is bound by instruction fetch and dependent-instruction latency, not by HBM
 * (issue saturates at 2 of 4 warp instructions per clock per SM because the L0 instruction cache
 * misses on every new 128-byte line): what makes it pay is that thousands of pictures (every picture of the look-ahead
 * window of every stream) are in flight at once, which is parallelism the host cores do not have.
 */
#pragma once
#include "k_common.cuh"
#include "kp_core.h"

/* Two launch shapes:
 *   kp_parse<8, 4>   CTAs of 8 warps, 4 per SM: Kp shares every SM with whatever else runs (the synchronous API, small batches);
 *   kp_parse<32, 1>  CTAs of 32 warps x 64 registers = a whole SM's register file: ONE CTA per SM and nothing else fits beside
 *                    it.  A launch of n CTAs therefore takes n SMs for itself and leaves the other SMs entirely to the
 *                    reconstruction kernels — spatial partitioning by resource exhaustion.  Time-sharing the SMs instead
 *                    costs both sides (measured: K4 runs 5x, K2 3x slower next to Kp's warps, Kp 1.6x), 17 % of the
 *                    device's throughput; see DESIGN.md section 5. */
struct KpBatch {
    const KpPic *pics;
    uint32_t n_pics;
    uint32_t *ticket;              /* device counter the warps draw pictures from; holds `ticket_base` when the launch starts */
    uint32_t ticket_base;          /* the host knows where a launch leaves the counter (n_pics + one failed draw per warp): no memset between launches */
    const KpTables *tables;        /* device copy of the host-built tables */
};

#define KP_SMEM_BYTES(warps) (sizeof(KpTables) + (size_t)(warps) * sizeof(KpStage))

template <int WARPS, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) kp_parse(KpBatch b)
{
    extern __shared__ __align__(16) unsigned char kp_smem[];
    KpTables &T = *reinterpret_cast<KpTables *>(kp_smem);
    KpStage *stage = reinterpret_cast<KpStage *>(kp_smem + sizeof(KpTables));
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(b.tables);
        uint4 *dst = reinterpret_cast<uint4 *>(&T);
        for (uint32_t i = threadIdx.x; i < sizeof(KpTables) / 16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    KpStage *st = &stage[threadIdx.x >> 5];
    for (;;) {
        uint32_t t = 0;
        if (lane == 0) t = atomicAdd(b.ticket, 1u) - b.ticket_base;
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= b.n_pics) break;
        const KpPic p = b.pics[t];
        kp_parse_picture(lane, p, st, &T);
    }
}
static_assert(sizeof(KpTables) % 16 == 0 && sizeof(KpStage) % 16 == 0, "shared memory layout of kernel Kp");

/* First kernel of a round: fetch the job table from the pinned host memory it was written to (every cudaHostAlloc'ed
 * byte is device-readable under unified addressing) and zero the round's control words.  This replaces a
 * cudaMemcpyAsync + cudaMemsetAsync on the compute stream: those are copy-engine work, and behind the ~800 MB of
 * frame copy-out of the previous round they made a round's kernels wait until that copy-out had finished (the
 * timeline showed every round starting exactly when the previous copy-out ended). */
__global__ void k0_stage(PicJob *jobs, const PicJob *host_jobs, int n_jobs, int32_t *ctrl, uint32_t ctrl_words)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    const uint32_t job_words = (uint32_t)n_jobs * (uint32_t)(sizeof(PicJob) / 4);
    for (uint32_t w = i; w < job_words; w += stride) reinterpret_cast<uint32_t *>(jobs)[w] = reinterpret_cast<const volatile uint32_t *>(host_jobs)[w];
    for (uint32_t w = i; w < ctrl_words; w += stride) ctrl[w] = 0;
}
static_assert(sizeof(PicJob) % 4 == 0, "PicJob is copied word-wise");

/* After Kp, before K1: copy what the parse found out about each picture into its job, and put the status words behind
 * the frame so that they travel to the host with it. */
__global__ void k0_jobs(PicJob *jobs, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    PicJob &j = jobs[i];
    const KpResult *r = j.kp_res;
    if (!r) return;
    j.n_intra = r->n_intra; j.n_inter = r->n_inter; j.any_deblock = r->any_deblock;
    j.n_conceal = r->n_conceal;
    j.conceal_list = reinterpret_cast<const uint32_t *>(j.coef_in + (size_t)r->conceal_offset * 16);
    *reinterpret_cast<h264b200_picstat_t *>(j.cur + j.stat_off) = r->stat;
}
