/* h264_engine.cu — the CUDA reconstruction engine of libh264b200.so (sm_100a).
 *
 * This is the device half of what the reference does inside
 * h264bsdDecodeMacroblock after the QP update (h264bsd_macroblock_layer.c:
 * 1099-1129: ProcessResidual, h264bsdIntraPrediction, h264bsdInterPrediction)
 * and in h264bsdFilterPicture (h264bsd_deblocking.c:574-639), re-organised for
 * a B200: the host parser (h264_slice.c) hands over WHOLE PICTURES as
 * macroblock records + coefficient slots (include/h264b200_records.h); the
 * engine copies them to HBM and runs four kernel families over a BATCH of
 * pictures (one per attached decoder instance) per launch:
 *     K1 k1_transform   dequant + inverse transforms          (k1_transform.cuh)
 *     K2 k2_inter       motion compensation + residual add    (k2_inter.cuh)
 *     K3 k3_intra       intra prediction wavefront            (k3_intra.cuh)
 *        k3c_conceal    spatial concealment of lost MBs       (k3c_conceal.cuh; only pictures that lost slices)
 *     K4 k4_deblock     deblocking wavefront                  (k4_deblock.cuh)
 * then copies each finished frame — or, on request, its cropped RGBA version (K5,
 * k5_rgba.cuh) — into a pinned host mirror, which is the pointer
 * h264bsdNextOutputPicture returns (Decoder.c:113-147 layout).
 *
 * HBM layout per decoder instance: n_slots frames back to back, each planar
 * I420, MB aligned (Y 16wm x 16hm, Cb, Cr; pitch = width) — exactly the
 * reference's output layout, so D2H needs no repacking.  Records and slots of a
 * picture live in a ring of NBUF input buffers (pinned host + device twin).
 *
 * DEVICE-PARSE instances (H264B200_ENGINE_DEVICE_PARSE) hand over slice NAL units instead of records
 * (include/h264b200_slices.h); kernel Kp (kp_core.h, kp_parse.cuh) parses the slice data of every queued
 * picture of every instance in one launch on its own stream — pictures of one stream are independent at
 * that level, so the look-ahead window of all streams is the parallelism — and writes the same records
 * and slots into HBM.  Pictures wait in a FIFO per instance; a scheduling thread (h264b200DecodeStreams) polls
 * h264b200EngineDrive, which launches Kp whenever enough pictures are queued and SMs of Kp's share are free, and ONE
 * reconstruction round at a time (the oldest picture of each instance whose Kp launch has finished), up to three
 * rounds deep; h264b200EngineAdvance / h264b200EngineSubmit are the lock-step forms of the same two steps.
 *
 * Streams: s_h2d (records / slice blocks in) -> s_parse[0..15] (Kp launches, overlapping) -> s_comp (K0 staging,
 * K1..K4) -> s_d2h (frames out), chained with events, so the copy-in of what comes next and the copy-out of what is
 * finished overlap the kernels; nothing a round or a Kp launch needs goes through a copy engine in front of it
 * (k0_stage, persistent ticket counters), because behind ~800 MB of frame copy-out it would wait.  There is no CPU reconstruction path in
 * this library: if CUDA is unusable h264_default_backend() returns NULL and
 * h264bsdDecode reports H264BSD_MEMALLOC_ERROR (reason on stderr).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <atomic>
#include <deque>
#include <mutex>
#include <vector>
#include "h264b200.h"
#include "h264b200_batch.h"
#include "h264_internal.h"
#include "k1_transform.cuh"
#include "k2_inter.cuh"
#include "k3_intra.cuh"
#include "k3c_conceal.cuh"
#include "k4_deblock.cuh"
#include "k5_rgba.cuh"
#include "kp_parse.cuh"

#define NBUF 3                 /* input buffers in flight per instance (host-parse; device-parse: look-ahead depth + 2) */
#define RING_EXTRA 4           /* device-parse input ring = look-ahead + this: the picture being scanned and up to DRIVE_ROUNDS launched ones whose kernels still read their buffers */
#define NSCR 8                 /* batch scratch sets in flight per engine */
#define NPAR 16                /* Kp launches in flight per engine: one scratch set and one CUDA stream each, so that they overlap —
                                  a launch over a quarter of the look-ahead window does not fill the SMs on its own */
#define STAT_TAIL 128          /* bytes behind every frame: h264b200_picstat_t of the picture (device-parse), copied out with it */
#define CTRL_HEAD 16           /* int32 words before the progress counters: [0] K3 ticket, [1] K4 ticket */

#define CUDA_TRY(call, fail) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    fprintf(stderr, "h264b200: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(e__), __FILE__, __LINE__); fail; } } while (0)

struct Inst;

struct PicBuf {
    h264_pic_input_t in;           /* host-parse: in.mbs / in.coef = ONE pinned block, records first, coefficient slots behind them;
                                      device-parse: in.block = pinned {picture header, slices} (include/h264b200_slices.h) */
    h264b200_mb_t *d_mbs;          /* host-parse: device twin of the pinned block (one cudaMemcpyAsync per picture);
                                      device-parse: where kernel Kp writes the records ... */
    int16_t *d_coef;               /* ... and the coefficient slots (worst-case capacity, KP_COEF_CAP) */
    uint32_t d_coef_cap;           /* slots */
    cudaEvent_t done;              /* (not owned) event of the round whose kernels read this buffer */
    int state;                     /* 0 free, 1 being filled by the parser, 2 queued, 3 launched */
    int own_host, own_dev;         /* the pinned / device memory of this buffer is its own allocation (not a slice of the instance's) */
    Inst *inst;
    /* device-parse */
    uint8_t *d_block; uint32_t d_block_cap; int own_dblock;
    KpMbCtx *d_ctx; KpResult *d_res;
    cudaEvent_t parsed;            /* (not owned) event of the Kp launch that parses this picture */
    uint32_t parse_seq;            /* which Kp launch; 0: not launched yet */
    int parse_slot;                /* index of the launch's ParseScratch (its `done` event is re-recorded by later launches of the slot) */
    int block_on_device;           /* the block was copied to d_block when the picture was submitted (be_pic_submit) */
    uint32_t gate_gen;             /* generation of the frame slot's host mirror that must have been released before this picture is launched */
    int tape_parse, tape_last_round;   /* retained runs: tape index of the Kp launch that fills / of the last round that read this buffer */
};

struct Inst {
    h264b200_engine *e;
    uint32_t wm, hm, n_mbs, n_slots;
    size_t frame_bytes;
    size_t frame_stride;                      /* frame_bytes + STAT_TAIL */
    uint8_t *d_frames, *h_frames;
    cudaEvent_t slot_ready[H264_MAX_SLOTS];   /* (not owned) copy-out event of the batch that last wrote the slot's mirror */
    uint8_t slot_flags[H264_MAX_SLOTS];       /* bit 1: a copy-out into the slot's mirror has been issued (slot_ready is valid) */
    uint8_t slot_scr[H264_MAX_SLOTS];         /* which scratch set's d2h_done slot_ready is ... */
    uint32_t slot_rseq[H264_MAX_SLOTS];       /* ... and the round it was recorded for: the event is re-recorded when the set serves a later round, and then
                                                 says nothing about this slot any more — except that its own copy-out ended long ago (slot_copy_pending) */
    uint32_t slot_qgen[H264_MAX_SLOTS];       /* pictures handed over (queued) into the slot so far */
    uint32_t slot_lgen[H264_MAX_SLOTS];       /* generation of the last LAUNCHED picture of the slot */
    uint32_t slot_popped[H264_MAX_SLOTS];     /* newest generation of the slot handed out through frame_host_async */
    std::atomic<uint32_t> slot_released[H264_MAX_SLOTS];   /* newest generation the caller is done with (h264b200PictureRelease) */
    PicBuf *bufs; int n_bufs;
    int next_buf;
    int batched;
    int lane_pref;                            /* reconstruction lane of this instance's pictures when rounds are split */
    int last_lane; cudaEvent_t last_done;     /* lane and kernels-finished event of the instance's last launched picture (ordering across lanes) */
    int dev_parse;                            /* slice data parsed by kernel Kp */
    std::deque<PicBuf *> *fifo;               /* pictures handed over, oldest first; one per round is launched (engine mutex) */
    std::atomic<uint32_t> n_pending;          /* == fifo->size(), readable without the mutex */
    uint8_t *h_blocks, *d_blocks, *d_parse;   /* device-parse: the instance-wide allocations the buffers are slices of */
    /* optional output formatting (K5): cropped RGBA instead of the I420 frame */
    int out_format; int cl, ct, cw, ch;
    uint8_t *d_rgba, *h_rgba; size_t rgba_bytes;
    uint8_t *d_pre;                           /* H264B200_ENGINE_TAP_PREDEBLOCK: the last launched picture before K4 */
};

struct BatchPlan { bool k1, k2, k3, k3c, k4, k0; uint32_t total_mbs; int max_hm; int n_jobs; };

/* Retained runs (H264B200_ENGINE_RETAIN) keep a TAPE of what was launched — Kp launches and reconstruction rounds, in
 * host order, with their inputs resident in HBM — so that h264b200EngineReplay can re-run the device work alone. */
struct Retained {
    int kind;                      /* 0: reconstruction round, 1: Kp launch */
    std::vector<void *> owned;     /* device allocations of this entry */
    cudaEvent_t ev;                /* recorded after the entry's work, every time it runs */
    uint32_t n_pics;
    /* round */
    std::vector<PicJob> jobs;      /* host copy */
    PicJob *d_jobs;
    Batch batch;
    size_t ctrl_words;
    BatchPlan pl;
    std::vector<int> wait_parse;   /* tape indices of the Kp launches that produce this round's records */
    unsigned long long *d_bytes;   /* algorithmic bytes per kernel family, counted on the device after the live run */
    uint64_t bytes[4]; bool bytes_read;
    /* Kp launch */
    KpBatch kp; int stream;
    int wait_round;                /* tape index of the last round that read a parse buffer this launch overwrites, -1: none */
    uint64_t kp_in_bytes, kp_rec_bytes;
    std::vector<int> rounds;       /* tape indices of the rounds that consume this launch (for its slot bytes) */
};

struct Scratch {
    PicJob *h_jobs, *d_jobs;
    uint32_t cap_jobs;
    int32_t *d_ctrl;
    size_t cap_ctrl;               /* words */
    cudaEvent_t done;              /* kernels of the batch finished */
    cudaEvent_t d2h_done;          /* copy-out of the batch finished */
    bool used;
    uint32_t seq;                  /* the round this set currently serves (engine round_seq) */
    uint32_t done_seq;             /* the last round of this set whose copy-out the scheduling thread has seen finished */
};

struct ParseScratch {              /* one Kp launch */
    KpPic *h_pics, *d_pics;
    uint32_t cap;
    uint32_t *d_ticket;
    uint32_t ticket_val;           /* where the last launch of this slot left d_ticket */
    cudaEvent_t done;              /* the launch has finished */
    bool used;
    uint32_t ctas;                 /* exclusive mode: SMs the launch owns while it runs */
    uint32_t seq;                  /* parse_seq of the launch */
    bool finished;                 /* the launch has been seen finished (h264b200EngineDrive) */
};

struct CopyList {
    std::vector<void *> dst, src; std::vector<size_t> size;
    void clear() { dst.clear(); src.clear(); size.clear(); }
    void add(void *d, const void *s_, size_t n) { if (n) { dst.push_back(d); src.push_back(const_cast<void *>(s_)); size.push_back(n); } }
};

struct h264b200_engine {
    int device, sm_count;
    /* H264B200_TIMELINE=file: device-side start / end of every Kp launch, reconstruction round and copy-out of the live run,
     * written as CSV when the engine is destroyed (what nsys would show; there is no nsys in the image) */
    struct TlEntry { int kind; uint32_t n; double host_ms; cudaEvent_t a, b; };
    std::vector<TlEntry> tl; const char *tl_path; cudaEvent_t tl_base; double tl_host0;
    uint32_t kp_sms;               /* > 0: SMs a Kp launch may take for itself (kp_parse<32, 1>); 0: Kp shares the SMs (kp_parse<8, 4>) */
    int kp_on_comp;                /* H264B200_KP_ON_COMP=1: Kp launches go to the reconstruction stream (serialised with K1..K4) instead of overlapping them */
    uint32_t wf_cap;               /* CTAs per SM the wavefront kernels K3 / K4 are launched with at most (tickets hand out the rows); H264B200_WF_CAP, default 16 */
    cudaStream_t s_h2d, s_comp, s_d2h, s_parse[NPAR];
    cudaEvent_t ev_h2d, ev_comp, ev_rep0, ev_rep1, ev_gate;
    cudaStream_t s_comp2; cudaEvent_t ev_comp2;   /* second reconstruction lane (H264B200_SPLIT_ROUNDS=1, experiment): a round goes out as two halves over
                                                     disjoint sets of instances on two streams, so that one half's wavefront tails overlap the other half's K1 / K2 */
    int split_rounds; uint32_t next_lane;
    std::mutex mu;
    std::vector<Inst *> insts;
    std::vector<Inst *> zombies;   /* shut-down instances whose frame pools retained batches still name */
    std::vector<Inst *> pool;      /* shut-down instances kept for reuse: pinned + device allocation is slow */
    std::vector<PicBuf *> tmp_parse, tmp_round;
    Scratch scr[NSCR];
    int next_scr;
    ParseScratch pscr[NPAR];
    int next_pscr;
    uint32_t parse_seq;
    bool copy_at_submit;           /* slices are uploaded by the thread that scanned them, when the picture is submitted (default; H264B200_COPY_AT_SUBMIT=0:
                                      by the Kp launch, hundreds of ~230 KB copies at once — measured 14 ms per launch next to the copy-out traffic,
                                      and the launch and the round behind it waited for them) */
    CopyList cl;                   /* scratch list of the launch being built (engine mutex) */
    double drv_locked_ms, drv_copy_ms, drv_locked_max; uint64_t drv_polls, drv_launches;
    uint64_t drv_miss[4];          /* per launched round: heads not taken because unparsed / in a running Kp launch / held back by an unreleased output; [3] rounds */   /* H264B200_TIMELINE: host time of the scheduling steps */
    std::atomic<uint32_t> copyouts_deferred;   /* rounds whose copy-out h264b200EngineDrive still has to issue (outside the mutex) */
    uint32_t round_seq;            /* rounds launched so far */
    double last_drive_ms;          /* host clock of the last h264b200EngineDrive (0: never): while a scheduling thread is polling, picture states come from what IT saw */
    uint32_t n_unparsed;           /* device-parse pictures queued and not yet handed to Kp (engine mutex) */
    KpTables *d_tables;
    uint32_t window, parse_threshold;
    uint32_t n_inst_hint; size_t inst_budget;   /* h264b200EngineSetStreams: instances to expect, device bytes each may spend on look-ahead buffers */
    uint32_t eff_window;           /* the look-ahead the instances created so far can actually hold (<= window) */
    uint32_t *d_err, *h_err;
    unsigned long long *d_trace; int trace_left;
    h264b200_stats_t st;
    uint32_t flags;                /* H264B200_ENGINE_* */
    std::vector<Retained *> retained;
    /* per-kernel timing of replays */
    std::vector<cudaEvent_t> tev;  /* rounds: 5 events, Kp launches: 2 events per timed tape entry */
    std::vector<int> tev_entry;
    double k_ms[5]; uint64_t k_bytes[5]; uint64_t k_launches[5];
    h264_backend_t be;
};

/* ----------------------------------------------------------------- helpers */
static void set_device(h264b200_engine *e) { cudaSetDevice(e->device); }

static size_t parse_bytes_per_buf(uint32_t n_mbs)
{
    /* records | contexts | result | coefficient slots (worst case) */
    return (size_t)n_mbs * (sizeof(h264b200_mb_t) + sizeof(KpMbCtx)) + 256 + (size_t)KP_COEF_CAP(n_mbs) * 32;
}
static uint32_t block_cap0(uint32_t n_mbs) { return (n_mbs * 48u + 4096u + 255u) & ~255u; }

static int picbuf_alloc_host(PicBuf *p, Inst *in, uint32_t coef_cap)
{
    memset(p, 0, sizeof *p);
    p->inst = in; p->tape_parse = p->tape_last_round = -1;
    const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
    CUDA_TRY(cudaHostAlloc((void **)&p->in.mbs, rec_bytes + (size_t)coef_cap * 32, cudaHostAllocDefault), return -1);
    p->in.coef = (int16_t *)((uint8_t *)p->in.mbs + rec_bytes);
    CUDA_TRY(cudaMalloc((void **)&p->d_mbs, rec_bytes + (size_t)coef_cap * 32), return -1);
    p->d_coef = (int16_t *)((uint8_t *)p->d_mbs + rec_bytes);
    p->in.coef_cap = coef_cap; p->d_coef_cap = coef_cap;
    p->own_host = p->own_dev = 1;
    p->in.priv = p;
    return 0;
}
static void picbuf_free(PicBuf *p)
{
    if (p->own_host && p->in.mbs) cudaFreeHost(p->in.mbs);
    if (p->own_host && p->in.block) cudaFreeHost(p->in.block);
    if (p->own_dev && p->d_mbs) cudaFree(p->d_mbs);
    if (p->own_dblock && p->d_block) cudaFree(p->d_block);
    memset(p, 0, sizeof *p);
}

/* Is the copy-out last issued into `slot` possibly still running?  slot_ready is the d2h_done event of a scratch set; once
 * that set has moved on to a later round the event belongs to THAT round — waiting for it (or polling it) would tie this
 * slot to a copy-out several rounds younger, which is what serialised every round behind the previous round's copy-out.
 * A set is reused only after its previous copy-out has finished (launch_round), so "moved on" means "arrived". */
static bool slot_copy_pending(const h264b200_engine *e, const Inst *in, int slot)
{
    if (!(in->slot_flags[slot] & 2)) return false;
    return __atomic_load_n(&e->scr[in->slot_scr[slot]].seq, __ATOMIC_ACQUIRE) == in->slot_rseq[slot];
}

/* ------------------------------------------------------------ kernel launch */
/* algorithmic bytes of a batch per kernel family, as SURVEY.md 8(d) defines them; runs once per retained round, outside
 * any timed region (the records of a device-parsed picture only exist on the device) */
__global__ void k_count_bytes(Batch b, unsigned long long *out)
{
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= b.total_mbs) return;
    const PicJob &job = b.jobs[find_job(b, g)];
    const h264b200_mb_t *m = job.mbs + (g - job.mb_base);
    const uint32_t nb = __popc(m->resid_mask & 0xffffffu), nd = __popc(m->resid_mask >> 24);
    if (m->mb_class == H264B200_MB_MISSING) return;
    unsigned long long k1 = (unsigned long long)(nb + nd) * 64, k2 = 0, k3 = 0, k4 = 0;
    if (m->mb_class == H264B200_MB_INTER) k2 = 896 + nb * 32;
    else k3 = 576 + (nb + (m->mb_class == H264B200_MB_IPCM ? 12 : 0)) * 32;
    if (m->dbk_flags) k4 = 896;
    for (int o = 16; o; o >>= 1) {
        k1 += __shfl_down_sync(0xffffffffu, k1, o); k2 += __shfl_down_sync(0xffffffffu, k2, o);
        k3 += __shfl_down_sync(0xffffffffu, k3, o); k4 += __shfl_down_sync(0xffffffffu, k4, o);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, k1); atomicAdd(out + 1, k2); atomicAdd(out + 2, k3); atomicAdd(out + 3, k4); }
}

static void launch_kernels(h264b200_engine *e, const Batch &b, const BatchPlan &pl, cudaEvent_t *tev, cudaStream_t s)
{
    if (pl.k0) { k0_jobs<<<(pl.n_jobs + 127) / 128, 128, 0, s>>>(const_cast<PicJob *>(b.jobs), pl.n_jobs); e->st.kernel_launches++; }
    if (tev) cudaEventRecord(tev[0], s);
    if (pl.k1) { uint32_t blocks = (pl.total_mbs * 8 + 255) / 256; k1_transform<<<blocks, 256, 0, s>>>(b); e->st.kernel_launches++; }   /* 8 lanes per macroblock */
    if (tev) cudaEventRecord(tev[1], s);
    if (pl.k2) { k2_inter<<<(pl.total_mbs * 16 + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, s>>>(b); e->st.kernel_launches++; }   /* 16 threads per macroblock */
    if (tev) cudaEventRecord(tev[2], s);
    uint32_t n_tasks = (uint32_t)pl.n_jobs * (uint32_t)pl.max_hm;
    if (pl.k3) {
        uint32_t blocks = (n_tasks + K3_WARPS - 1) / K3_WARPS, cap = (uint32_t)e->sm_count * e->wf_cap;
        k3_intra<<<blocks < cap ? blocks : cap, K3_WARPS * 32, 0, s>>>(b); e->st.kernel_launches++;
    }
    if (pl.k3c) { k3c_conceal<<<pl.n_jobs, 32, 0, s>>>(b); e->st.kernel_launches++; }   /* lost slices only: one warp per picture */
    if (tev) cudaEventRecord(tev[3], s);
    if (pl.k4) {                   /* one warp per PAIR of macroblock rows */
        uint32_t n_pairs = (uint32_t)pl.n_jobs * (((uint32_t)pl.max_hm + 1) / 2);
        uint32_t blocks = (n_pairs + K4_WARPS - 1) / K4_WARPS, cap = (uint32_t)e->sm_count * e->wf_cap;
        k4_deblock<<<blocks < cap ? blocks : cap, K4_WARPS * 32, 0, s>>>(b);
        e->st.kernel_launches++;
    }
    if (tev) cudaEventRecord(tev[4], s);
}

/* Kp launch: exclusive mode (kp_sms > 0, the default of batched device-parse engines) = CTAs that fill an SM each, at most
 * kp_sms of them, so the launch owns those SMs and the rest of the device stays with K1..K4 (kp_parse.cuh); shared mode =
 * 8-warp CTAs on every SM. */
static uint32_t kp_ctas(const h264b200_engine *e, uint32_t n_pics) { uint32_t b = (n_pics + 31) / 32; return b > e->kp_sms ? e->kp_sms : b; }
static void kp_launch(h264b200_engine *e, const KpBatch &kb, cudaStream_t s)
{
    if (e->kp_sms > 0) {
        kp_parse<32, 1><<<kp_ctas(e, kb.n_pics), 1024, KP_SMEM_BYTES(32), s>>>(kb);
    } else {
        uint32_t blocks = (kb.n_pics + 7) / 8, cap = (uint32_t)e->sm_count * 4;
        kp_parse<8, 4><<<blocks < cap ? blocks : cap, 256, KP_SMEM_BYTES(8), s>>>(kb);
    }
}

static double host_ms_now() { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return 1e3 * (double)t.tv_sec + 1e-6 * (double)t.tv_nsec; }
static void tl_begin(h264b200_engine *e, int kind, uint32_t n, cudaStream_t s)
{
    if (!e->tl_path || e->tl.size() >= 20000) return;
    h264b200_engine::TlEntry t; t.kind = kind; t.n = n; t.host_ms = host_ms_now() - e->tl_host0;
    cudaEventCreate(&t.a); cudaEventCreate(&t.b);
    cudaEventRecord(t.a, s);
    e->tl.push_back(t);
}
static void tl_end(h264b200_engine *e, cudaStream_t s) { if (e->tl_path && !e->tl.empty() && e->tl.size() < 20000) cudaEventRecord(e->tl.back().b, s); }
static void tl_dump(h264b200_engine *e)
{
    if (!e->tl_path) return;
    FILE *f = fopen(e->tl_path, "a");
    if (e->drv_polls) fprintf(stderr, "h264b200 scheduling thread: %llu steps, %llu launched something: %.1f ms under the mutex (max %.2f), %.1f ms issuing copy-outs\n",
                              (unsigned long long)e->drv_polls, (unsigned long long)e->drv_launches, e->drv_locked_ms, e->drv_locked_max, e->drv_copy_ms);
    if (e->drv_miss[3]) fprintf(stderr, "h264b200 rounds: %llu; streams left out per round on average: %.1f head not yet handed to Kp, %.1f in a running Kp launch, %.1f held back by an unreleased output\n",
                                (unsigned long long)e->drv_miss[3], (double)e->drv_miss[0] / e->drv_miss[3], (double)e->drv_miss[1] / e->drv_miss[3], (double)e->drv_miss[2] / e->drv_miss[3]);
    if (f) {
        fprintf(f, "kind,pictures,host_launch_ms,gpu_start_ms,gpu_end_ms\n");
        for (auto &t : e->tl) {
            float a = 0, b = 0;
            if (cudaEventElapsedTime(&a, e->tl_base, t.a) != cudaSuccess || cudaEventElapsedTime(&b, e->tl_base, t.b) != cudaSuccess) continue;
            fprintf(f, "%s,%u,%.3f,%.3f,%.3f\n", t.kind == 0 ? "round" : t.kind == 1 ? "kp" : "d2h", t.n, t.host_ms, a, b);
        }
        fclose(f);
    }
    for (auto &t : e->tl) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    e->tl.clear();
}

/* issue the copies of a list, one cudaMemcpyAsync each */
static int copy_list(CopyList &c, cudaMemcpyKind kind, cudaStream_t s)
{
    for (size_t i = 0; i < c.dst.size(); i++) CUDA_TRY(cudaMemcpyAsync(c.dst[i], c.src[i], c.size[i], kind, s), return -1);
    return 0;
}

/* ------------------------------------------------------------------ Kp launch */
/* engine mutex held.  Copy the blocks of the given queued pictures to the device and parse them all in one launch. */
static int launch_parse(h264b200_engine *e, std::vector<PicBuf *> &list)
{
    const uint32_t n = (uint32_t)list.size();
    const bool retain = (e->flags & H264B200_ENGINE_RETAIN) != 0;
    ParseScratch &ps = e->pscr[e->next_pscr];
    const int pslot = e->next_pscr;
    const int stream = e->kp_on_comp ? -1 : e->next_pscr;
    e->next_pscr = (e->next_pscr + 1) % NPAR;
    if (ps.used && !ps.finished) cudaEventSynchronize(ps.done);
    if (ps.cap < n) {
        if (ps.h_pics) cudaFreeHost(ps.h_pics);
        if (ps.d_pics) cudaFree(ps.d_pics);
        /* room for the largest launch the engine makes, once: growing it later means cudaFree / cudaMalloc, which wait for
         * everything on the device — with Kp launches of 0.2 s in flight that stalled the scheduling thread (and the
         * mutex) for up to a second whenever a launch was a few pictures larger than the last one of its slot */
        ps.cap = n + 64;
        if (ps.cap < e->kp_sms * 32u + 64u) ps.cap = e->kp_sms * 32u + 64u;
        if (ps.cap < 1024) ps.cap = 1024;
        CUDA_TRY(cudaHostAlloc((void **)&ps.h_pics, ps.cap * sizeof(KpPic), cudaHostAllocDefault), return -1);
        CUDA_TRY(cudaMalloc((void **)&ps.d_pics, ps.cap * sizeof(KpPic)), return -1);
    }
    Retained *ret = nullptr;
    KpPic *d_pics = ps.d_pics; uint32_t *d_ticket = ps.d_ticket;
    if (retain) {
        ret = new Retained();
        ret->kind = 1; ret->stream = stream; ret->wait_round = -1; ret->n_pics = n; ret->d_jobs = nullptr; ret->d_bytes = nullptr;
        ret->kp_in_bytes = ret->kp_rec_bytes = 0; ret->bytes_read = false;
        CUDA_TRY(cudaEventCreateWithFlags(&ret->ev, cudaEventDisableTiming), return -1);
        CUDA_TRY(cudaMalloc((void **)&d_pics, n * sizeof(KpPic)), return -1);
        ret->owned.push_back(d_pics);
        CUDA_TRY(cudaMalloc((void **)&d_ticket, 64), return -1);
        ret->owned.push_back(d_ticket);
    }
    e->parse_seq++;
    if (!e->parse_seq) e->parse_seq = 1;
    uint64_t in_bytes = 0;
    e->cl.clear();
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = list[i]; Inst *in = p->inst; h264_pic_input_t *pic = &p->in;
        uint8_t *d_block = p->d_block;
        if (retain) {              /* the block stays resident for the replay */
            void *a = nullptr;
            CUDA_TRY(cudaMalloc(&a, pic->block_used), return -1);
            ret->owned.push_back(a);
            d_block = (uint8_t *)a;
            if (p->tape_last_round > ret->wait_round) ret->wait_round = p->tape_last_round;
            p->tape_parse = (int)e->retained.size();
            ret->kp_in_bytes += pic->block_used; ret->kp_rec_bytes += (uint64_t)in->n_mbs * sizeof(h264b200_mb_t);
        } else if (p->d_block_cap < pic->block_used) {
            /* no earlier launch reads this buffer's block any more (pic_begin waited for `done`) */
            if (p->own_dblock) cudaFree(p->d_block);
            p->d_block_cap = pic->block_cap; p->own_dblock = 1;
            CUDA_TRY(cudaMalloc((void **)&p->d_block, p->d_block_cap), return -1);
            d_block = p->d_block;
        }
        if (!p->block_on_device || retain) e->cl.add(d_block, pic->block, pic->block_used);
        in_bytes += pic->block_used;
        KpPic &kp = ps.h_pics[i];
        kp.block = d_block; kp.mbs = p->d_mbs; kp.coef = p->d_coef; kp.ctx = p->d_ctx; kp.coef_cap = p->d_coef_cap; kp.pad = 0; kp.res = p->d_res;
        p->parsed = retain ? ret->ev : ps.done;
        p->parse_seq = e->parse_seq; p->parse_slot = pslot;
    }
    if (copy_list(e->cl, cudaMemcpyHostToDevice, e->s_h2d)) return -1;
    e->st.h2d_bytes += in_bytes;
    e->n_unparsed = e->n_unparsed > n ? e->n_unparsed - n : 0;
    cudaStream_t s = stream < 0 ? e->s_comp : e->s_parse[stream];
    cudaEventRecord(e->ev_h2d, e->s_h2d);
    cudaStreamWaitEvent(s, e->ev_h2d, 0);
    KpBatch kb; kb.pics = d_pics; kb.n_pics = n; kb.ticket = d_ticket; kb.tables = e->d_tables; kb.ticket_base = 0;
    if (retain) {                  /* the replay needs the table on the device and starts every run from a zeroed counter */
        cudaMemcpyAsync(d_pics, ps.h_pics, n * sizeof(KpPic), cudaMemcpyHostToDevice, s);
        cudaMemsetAsync(d_ticket, 0, 64, s);
    } else {
        /* no copy-engine work in front of the kernel (it would queue behind the frame copy-out): the warps read their
         * picture descriptors straight from the pinned table, and the counter continues where the slot's last launch left it */
        kb.pics = ps.h_pics;
        kb.ticket_base = ps.ticket_val;
        const uint32_t warps = e->kp_sms ? kp_ctas(e, n) * 32u : (((n + 7) / 8 < (uint32_t)e->sm_count * 4 ? (n + 7) / 8 : (uint32_t)e->sm_count * 4) * 8u);
        ps.ticket_val += n + warps;
    }
    tl_begin(e, 1, n, s);
    kp_launch(e, kb, s);
    tl_end(e, s);
    e->st.kernel_launches++; e->st.kp_launches++; e->st.kp_pictures += n;
    cudaEventRecord(ps.done, s); ps.used = true; ps.ctas = e->kp_sms ? kp_ctas(e, n) : 0; ps.seq = e->parse_seq; ps.finished = false;
    if (retain) {
        cudaEventRecord(ret->ev, s);
        ret->kp = kb;
        e->retained.push_back(ret);
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) { fprintf(stderr, "h264b200: Kp launch failed: %s\n", cudaGetErrorString(le)); return -1; }
    return 0;
}

/* ------------------------------------------------------------------ round launch */
/* engine mutex held.  Reconstruct the given pictures (at most one per instance) as one batch. */
/* The copy-out half of a round: per picture one copy of the finished frame (and the status words behind it) into its pinned
 * mirror, then the event that says they have arrived, then — only then — the pictures are published as launched to those
 * who poll without the mutex (be_frame_state): slot_ready names an event that has been recorded for THIS round.
 * Touches no engine state, so the free-running schedule runs it outside the engine mutex: hundreds of copy calls per round
 * would otherwise keep every submitting worker waiting. */
struct CopyOut { std::vector<PicBuf *> list; cudaEvent_t d2h_done; };
static void copy_out_issue(h264b200_engine *e, CopyOut &co)
{
    const uint32_t n = (uint32_t)co.list.size();
    tl_begin(e, 2, n, e->s_d2h);
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = co.list[i]; Inst *in = p->inst; const int slot = p->in.cur_slot;
        uint8_t *d_frame = in->d_frames + (size_t)slot * in->frame_stride, *h_frame = in->h_frames + (size_t)slot * in->frame_stride;
        if (in->out_format == H264B200_OUT_RGBA && in->d_rgba) {
            /* K5 runs on the copy-out stream: it only reads the finished frame */
            RgbaJob rj; rj.frame = d_frame; rj.out = in->d_rgba + (size_t)slot * in->rgba_bytes;
            rj.W = (int)in->wm * 16; rj.H = (int)in->hm * 16; rj.cl = in->cl; rj.ct = in->ct; rj.cw = in->cw; rj.ch = in->ch;
            const int items = ((rj.cw + 3) / 4) * ((rj.ch + 1) / 2);
            k5_rgba<<<(items + 255) / 256, 256, 0, e->s_d2h>>>(rj);
            if (!(e->flags & H264B200_ENGINE_NO_D2H)) {
                cudaMemcpyAsync(in->h_rgba + (size_t)slot * in->rgba_bytes, rj.out, in->rgba_bytes, cudaMemcpyDeviceToHost, e->s_d2h);
                cudaMemcpyAsync(h_frame + in->frame_bytes, d_frame + in->frame_bytes, sizeof(h264b200_picstat_t), cudaMemcpyDeviceToHost, e->s_d2h);
            }
        } else if (!(e->flags & H264B200_ENGINE_NO_D2H)) {
            /* the frame and the status words behind it: one copy */
            cudaMemcpyAsync(h_frame, d_frame, in->frame_bytes + sizeof(h264b200_picstat_t), cudaMemcpyDeviceToHost, e->s_d2h);
        }
    }
    cudaMemcpyAsync(e->h_err, e->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_d2h);
    tl_end(e, e->s_d2h);
    cudaEventRecord(co.d2h_done, e->s_d2h);
    for (uint32_t i = 0; i < n; i++) {
        Inst *in = co.list[i]->inst; const int slot = co.list[i]->in.cur_slot;
        __atomic_store_n(&in->slot_lgen[slot], in->slot_lgen[slot] + 1, __ATOMIC_RELEASE);
    }
}

static uint32_t launch_round(h264b200_engine *e, std::vector<PicBuf *> &list, CopyOut *defer = nullptr, int lane = 0)
{
    cudaStream_t cs = lane ? e->s_comp2 : e->s_comp;
    cudaEvent_t evc = lane ? e->ev_comp2 : e->ev_comp;
    uint32_t n = (uint32_t)list.size();
    if (!n) return 0;
    const bool retain = (e->flags & H264B200_ENGINE_RETAIN) != 0;
    Retained *ret = nullptr;
    Scratch &sc = e->scr[e->next_scr];
    e->next_scr = (e->next_scr + 1) % NSCR;
    if (sc.used) { cudaEventSynchronize(sc.done); cudaEventSynchronize(sc.d2h_done); }   /* in the free-running schedule both ended long ago */
    const uint32_t rseq = ++e->round_seq;
    __atomic_store_n(&sc.seq, rseq, __ATOMIC_RELEASE);      /* from here on the set's events speak for this round only (slot_copy_pending) */
    const int scr_index = (int)(&sc - e->scr);
    size_t ctrl_words = CTRL_HEAD;
    for (PicBuf *p : list) ctrl_words += 2 * (size_t)p->inst->hm;
    if (sc.cap_jobs < n) {
        if (sc.h_jobs) cudaFreeHost(sc.h_jobs);
        if (sc.d_jobs) cudaFree(sc.d_jobs);
        sc.h_jobs = nullptr; sc.d_jobs = nullptr;
        sc.cap_jobs = n + 16;
        if (sc.cap_jobs < (uint32_t)e->insts.size() + 16) sc.cap_jobs = (uint32_t)e->insts.size() + 16;      /* once: growing means cudaFree, which waits for the whole device */
        CUDA_TRY(cudaHostAlloc((void **)&sc.h_jobs, sc.cap_jobs * sizeof(PicJob), cudaHostAllocDefault), { sc.cap_jobs = 0; goto fail; });
        CUDA_TRY(cudaMalloc((void **)&sc.d_jobs, sc.cap_jobs * sizeof(PicJob)), { sc.cap_jobs = 0; goto fail; });
    }
    if (sc.cap_ctrl < ctrl_words) {
        if (sc.d_ctrl) cudaFree(sc.d_ctrl);
        sc.d_ctrl = nullptr;
        sc.cap_ctrl = ctrl_words + 1024;
        { size_t all = CTRL_HEAD + 1024; for (Inst *in : e->insts) all += 2 * (size_t)in->hm; if (sc.cap_ctrl < all) sc.cap_ctrl = all; }
        CUDA_TRY(cudaMalloc((void **)&sc.d_ctrl, sc.cap_ctrl * sizeof(int32_t)), { sc.cap_ctrl = 0; goto fail; });
    }
    {
    PicJob *d_jobs = sc.d_jobs; int32_t *d_ctrl = sc.d_ctrl;
    if (retain) {                  /* a retained round owns its job table and control area */
        ret = new Retained();
        ret->kind = 0; ret->n_pics = n; ret->bytes_read = false; ret->wait_round = -1; ret->d_bytes = nullptr;
        memset(ret->bytes, 0, sizeof ret->bytes);
        CUDA_TRY(cudaEventCreateWithFlags(&ret->ev, cudaEventDisableTiming), goto fail);
        CUDA_TRY(cudaMalloc((void **)&ret->d_jobs, n * sizeof(PicJob)), goto fail);
        ret->owned.push_back(ret->d_jobs);
        d_jobs = ret->d_jobs;
        CUDA_TRY(cudaMalloc((void **)&d_ctrl, ctrl_words * sizeof(int32_t)), goto fail);
        ret->owned.push_back(d_ctrl);
        CUDA_TRY(cudaMalloc((void **)&ret->d_bytes, 4 * sizeof(unsigned long long)), goto fail);
        ret->owned.push_back(ret->d_bytes);
    }

    BatchPlan pl; memset(&pl, 0, sizeof pl);
    uint32_t mb_base = 0; size_t prog_off = CTRL_HEAD;
    uint32_t last_parse_seq = 0;
    bool uploaded = false;
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = list[i]; Inst *in = p->inst; h264_pic_input_t *pic = &p->in;
        h264b200_mb_t *d_mbs = p->d_mbs; int16_t *d_coef_in = p->d_coef, *d_coef = p->d_coef;
        PicJob &j = sc.h_jobs[i];
        memset(&j, 0, sizeof j);
        if (in->dev_parse) {
            /* records and slots are where kernel Kp left them; its launch must have finished */
            if (p->parse_seq != last_parse_seq) {
                /* the slot's event may have been re-recorded by a LATER launch — then this picture's launch finished long ago
                 * (a slot is reused only after its launch ended) and waiting on the event would wait for the wrong launch */
                if (retain || e->pscr[p->parse_slot].seq == p->parse_seq) cudaStreamWaitEvent(cs, p->parsed, 0);
                last_parse_seq = p->parse_seq;
            }
            j.kp_res = p->d_res;
            pl.k0 = pl.k1 = pl.k3 = pl.k3c = pl.k4 = true;
            if (pic->has_p_slice) pl.k2 = true;
            if (retain) {
                bool seen = false;
                for (int w : ret->wait_parse) if (w == p->tape_parse) seen = true;
                if (!seen && p->tape_parse >= 0) { ret->wait_parse.push_back(p->tape_parse); e->retained[(size_t)p->tape_parse]->rounds.push_back((int)e->retained.size()); }
                p->tape_last_round = (int)e->retained.size();
            }
        } else {
            size_t coef_bytes = (size_t)pic->coef_used * 32;
            const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
            if (retain) {                                          /* [records | levels] and a separate residual buffer, kept */
                void *a = nullptr, *c = nullptr;
                CUDA_TRY(cudaMalloc(&a, rec_bytes + coef_bytes + 32), goto fail);
                ret->owned.push_back(a);
                CUDA_TRY(cudaMalloc(&c, coef_bytes + 32), goto fail);
                ret->owned.push_back(c);
                d_mbs = (h264b200_mb_t *)a; d_coef_in = (int16_t *)((uint8_t *)a + rec_bytes); d_coef = (int16_t *)c;
            } else if (p->d_coef_cap < pic->coef_used) {          /* the host side grew: follow */
                /* no earlier batch reads this ring slot any more (pic_begin waited for `done`), so it can be replaced */
                cudaFree(p->d_mbs);
                p->d_mbs = nullptr;
                p->d_coef_cap = pic->coef_cap;
                CUDA_TRY(cudaMalloc((void **)&p->d_mbs, rec_bytes + (size_t)p->d_coef_cap * 32), goto fail);
                p->d_coef = (int16_t *)((uint8_t *)p->d_mbs + rec_bytes);
                d_mbs = p->d_mbs; d_coef_in = d_coef = p->d_coef;
            }
            /* records and coefficient slots are adjacent on both sides: one copy */
            CUDA_TRY(cudaMemcpyAsync(d_mbs, pic->mbs, rec_bytes + coef_bytes, cudaMemcpyHostToDevice, e->s_h2d), goto fail);
            uploaded = true;
            e->st.h2d_bytes += rec_bytes + coef_bytes;
            j.n_intra = pic->n_intra; j.n_inter = pic->n_inter; j.any_deblock = pic->any_deblock;
            j.n_conceal = pic->n_conceal;
            j.conceal_list = reinterpret_cast<const uint32_t *>(d_coef_in + (size_t)pic->conceal_offset * 16);
            if (pic->n_conceal) pl.k3c = true;
            if (pic->coef_used) pl.k1 = true;
            if (pic->n_inter) pl.k2 = true;
            if (pic->n_intra) pl.k3 = true;
            if (pic->any_deblock) pl.k4 = true;
        }
        j.mbs = d_mbs; j.coef_in = d_coef_in; j.coef = d_coef;
        j.cur = in->d_frames + (size_t)pic->cur_slot * in->frame_stride;
        j.frames = in->d_frames; j.frame_bytes = (uint32_t)in->frame_stride;
        j.stat_off = (uint32_t)in->frame_bytes;
        j.wm = (int32_t)in->wm; j.hm = (int32_t)in->hm;
        j.progress = d_ctrl + prog_off; prog_off += 2 * (size_t)in->hm;
        j.mb_base = mb_base; mb_base += in->n_mbs;
        if ((int)in->hm > pl.max_hm) pl.max_hm = (int)in->hm;
        /* the frame being written may still be on its way to the host from an earlier batch */
        if (slot_copy_pending(e, in, pic->cur_slot)) cudaStreamWaitEvent(cs, in->slot_ready[pic->cur_slot], 0);
        /* the instance's previous picture ran on the other lane: its kernels first (a re-recorded event only waits longer) */
        if (in->last_done && in->last_lane != lane) cudaStreamWaitEvent(cs, in->last_done, 0);
        in->last_lane = lane; in->last_done = sc.done;
    }
    pl.total_mbs = mb_base; pl.n_jobs = (int)n;
    if (e->flags & H264B200_ENGINE_NO_RECON) pl.k1 = pl.k2 = pl.k3 = pl.k3c = pl.k4 = false;

    Batch b;
    b.jobs = d_jobs; b.n_jobs = (int32_t)n; b.max_hm = pl.max_hm; b.total_mbs = mb_base;
    b.uniform_mbs = list[0]->inst->n_mbs;
    for (PicBuf *p : list) if (p->inst->n_mbs != b.uniform_mbs) b.uniform_mbs = 0;
    b.tickets = (uint32_t *)d_ctrl; b.error_flags = e->d_err; b.trace = e->trace_left > 0 ? e->d_trace : nullptr;

    if (uploaded) {                /* records of host-parsed pictures; device-parsed ones have nothing in the upload stream a round depends on */
        cudaEventRecord(e->ev_h2d, e->s_h2d);
        cudaStreamWaitEvent(cs, e->ev_h2d, 0);
    }
    k0_stage<<<32, 256, 0, cs>>>(d_jobs, sc.h_jobs, (int)n, d_ctrl, (uint32_t)ctrl_words);   /* job table + zeroed control words, without the copy engines */
    e->st.kernel_launches++;
    tl_begin(e, 0, n, cs);
    if (e->flags & H264B200_ENGINE_TAP_PREDEBLOCK) {
        /* parity aid: K0..K3, the pictures copied aside, then K4 */
        BatchPlan a = pl, c; memset(&c, 0, sizeof c);
        a.k4 = false;
        c.k4 = pl.k4; c.total_mbs = pl.total_mbs; c.max_hm = pl.max_hm; c.n_jobs = pl.n_jobs;
        launch_kernels(e, b, a, nullptr, cs);
        for (uint32_t i = 0; i < n; i++) {
            Inst *in = list[i]->inst;
            if (!in->d_pre && cudaMalloc((void **)&in->d_pre, in->frame_bytes) != cudaSuccess) { in->d_pre = nullptr; continue; }
            cudaMemcpyAsync(in->d_pre, in->d_frames + (size_t)list[i]->in.cur_slot * in->frame_stride, in->frame_bytes, cudaMemcpyDeviceToDevice, cs);
        }
        launch_kernels(e, b, c, nullptr, cs);
    } else
    launch_kernels(e, b, pl, nullptr, cs);
    tl_end(e, cs);
    if (retain) {
        cudaMemsetAsync(ret->d_bytes, 0, 4 * sizeof(unsigned long long), cs);
        k_count_bytes<<<(mb_base + 255) / 256, 256, 0, cs>>>(b, ret->d_bytes);
    }
    if (b.trace) {                 /* debug: dump the wavefront timing of job 0 of this batch */
        const int rows = sc.h_jobs[0].hm < 512 ? sc.h_jobs[0].hm : 512;
        std::vector<unsigned long long> h(256 + 4 * 512);
        cudaStreamSynchronize(cs);
        cudaMemcpy(h.data(), e->d_trace, h.size() * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "h264b200 trace: batch of %u, job0 %dx%d MBs; per row: start, first-mb, mid, end (us from row 0 start)\n", n, sc.h_jobs[0].wm, sc.h_jobs[0].hm);
        for (int r = 0; r < rows; r++) { const unsigned long long *q = &h[256 + r * 4], t0 = h[256];
            fprintf(stderr, "  row %2d: %8.1f %8.1f %8.1f %8.1f\n", r, (q[0] - t0) / 1e3, (q[1] - t0) / 1e3, (q[2] - t0) / 1e3, (q[3] - t0) / 1e3); }
        cudaMemset(e->d_trace, 0, h.size() * 8);
        e->trace_left--;
    }
    cudaEventRecord(sc.done, cs); sc.used = true;
    cudaEventRecord(evc, cs);
    cudaStreamWaitEvent(e->s_d2h, evc, 0);
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = list[i]; Inst *in = p->inst; const int slot = p->in.cur_slot;
        p->done = sc.done;
        p->state = 3;
        in->slot_ready[slot] = sc.d2h_done;
        in->slot_scr[slot] = (uint8_t)scr_index; in->slot_rseq[slot] = rseq;
        in->slot_flags[slot] = 2;
        if (in->out_format == H264B200_OUT_RGBA && in->d_rgba) e->st.kernel_launches++;      /* K5, launched with the copy-out */
        if (!(e->flags & H264B200_ENGINE_NO_D2H)) e->st.d2h_bytes += (in->out_format == H264B200_OUT_RGBA && in->d_rgba) ? in->rgba_bytes : in->frame_bytes;
    }
    if (defer) { defer->list = list; defer->d2h_done = sc.d2h_done; e->copyouts_deferred.fetch_add(1); }      /* h264b200EngineDrive issues the copies after it has let go of the mutex */
    else { CopyOut co; co.list = list; co.d2h_done = sc.d2h_done; copy_out_issue(e, co); }
    e->st.pictures += n; e->st.batches++;
    if (retain) {
        cudaEventRecord(ret->ev, cs);
        ret->jobs.assign(sc.h_jobs, sc.h_jobs + n);
        ret->batch = b; ret->batch.trace = nullptr; ret->ctrl_words = ctrl_words;
        ret->pl = pl;
        e->retained.push_back(ret);
        ret = nullptr;
    }
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) fprintf(stderr, "h264b200: kernel launch failed: %s\n", cudaGetErrorString(le));
    return n;
    }
fail:
    /* an allocation or copy failed before anything of this round was launched: give the buffers back and remember the
     * failure per instance (frame_host then returns NULL for it) instead of leaving them queued */
    if (ret) { for (void *q : ret->owned) cudaFree(q); delete ret; }
    for (PicBuf *p : list) { p->state = 0; p->inst->slot_lgen[p->in.cur_slot]++; p->inst->slot_flags[p->in.cur_slot] |= 4; }
    return 0;
}

/* ------------------------------------------------------------------ scheduling */
static bool gated(const PicBuf *p)
{
    const Inst *in = p->inst;
    return (int32_t)(in->slot_released[p->in.cur_slot].load(std::memory_order_acquire) - p->gate_gen) < 0;
}

/* engine mutex held.  One scheduling step: Kp over the queued device-parse pictures when it pays (or must happen), then
 * one reconstruction round over the oldest queued picture of every instance. */
static uint32_t advance_locked(h264b200_engine *e, bool force)
{
    set_device(e);
    std::vector<PicBuf *> &pl = e->tmp_parse; pl.clear();
    bool head_unparsed = false;
    for (Inst *in : e->insts) {
        if (!in->dev_parse || in->fifo->empty()) continue;
        bool first = true;
        for (PicBuf *p : *in->fifo) {
            if (!p->parse_seq) { pl.push_back(p); if (first && !gated(p)) head_unparsed = true; }
            first = false;
        }
    }
    /* a launch over `parse_threshold` pictures (a quarter of the look-ahead of every stream) as soon as that many wait
     * and one of the NPAR launch slots is free: up to NPAR launches overlap on the device, and the reconstruction
     * rounds of pictures parsed earlier run beside them */
    const ParseScratch &nps = e->pscr[e->next_pscr];
    bool slot_free = !nps.used || cudaEventQuery(nps.done) == cudaSuccess;
    if (slot_free && e->kp_sms && pl.size() >= e->parse_threshold) {
        /* exclusive launches own their SMs: together they must stay within the SMs given to Kp, or the reconstruction
         * rounds would find no SM until a launch ends */
        uint32_t busy = 0;
        for (int k = 0; k < NPAR; k++) if (e->pscr[k].used && e->pscr[k].ctas && cudaEventQuery(e->pscr[k].done) != cudaSuccess) busy += e->pscr[k].ctas;
        if (busy + kp_ctas(e, (uint32_t)pl.size()) > e->kp_sms) slot_free = false;
    }
    if (!pl.empty() && (force || head_unparsed || (pl.size() >= e->parse_threshold && slot_free))) {
        if (launch_parse(e, pl)) {
            for (PicBuf *p : pl) p->inst->slot_flags[p->in.cur_slot] |= 4;
        }
    }
    std::vector<PicBuf *> &rl = e->tmp_round; rl.clear();
    for (Inst *in : e->insts) {
        if (in->fifo->empty()) continue;
        PicBuf *p = in->fifo->front();
        if (in->dev_parse && !p->parse_seq) continue;
        if (gated(p)) continue;
        rl.push_back(p);
    }
    if (rl.empty()) return 0;
    for (PicBuf *p : rl) { p->inst->fifo->pop_front(); p->inst->n_pending.fetch_sub(1, std::memory_order_release); }
    launch_round(e, rl);
    return (uint32_t)rl.size();
}
static uint32_t advance_all_locked(h264b200_engine *e)
{
    uint32_t total = 0, n;
    while ((n = advance_locked(e, true)) != 0) total += n;
    return total;
}

/* engine mutex held.  One step of the FREE-RUNNING schedule (h264b200EngineDrive): what a dedicated scheduling thread
 * calls a few thousand times per second while worker threads submit pictures and collect outputs on their own.  A call
 * that finds nothing to launch costs a handful of event queries.
 *   Kp: as soon as `parse_threshold` unparsed pictures are queued and SMs of Kp's share are free, a launch over as many
 *       pictures as fit those SMs — the oldest unparsed picture of every instance first, then the second oldest ..., so
 *       that a truncated launch still serves every stream — in multiples of 32 (every CTA of the launch full).  Fewer than the threshold are only
 *       launched when no Kp launch is running at all and the caller reports that its workers are idle (`idle`): the
 *       look-ahead windows are full or the streams are ending, so waiting would not bring more.
 *   round: the oldest queued picture of every instance whose Kp launch has FINISHED (a round then starts at once and its
 *       buffers come back soon), once at least 7 of 8 instances that have anything queued are ready — under the same
 *       "nothing else will happen" condition any — while fewer than DRIVE_ROUNDS rounds are on the device (launched,
 *       copy-out not finished): enough to keep kernels and copy-out busy back to back.
 * *kp_pics = pictures handed to Kp by this call.  Returns the pictures of the round launched (0: none). */
#define DRIVE_ROUNDS 3
static uint32_t drive_locked(h264b200_engine *e, int idle, uint32_t *kp_pics, CopyOut *defer, CopyOut *defer2)
{
    set_device(e);
    if (kp_pics) *kp_pics = 0;
    /* Kp launches still running (a slot once seen finished is not queried again) */
    uint32_t running[NPAR], busy_ctas = 0; int n_running = 0;
    for (int k = 0; k < NPAR; k++) {
        ParseScratch &ps = e->pscr[k];
        if (!ps.used || ps.finished) continue;
        if (cudaEventQuery(ps.done) == cudaSuccess) { ps.finished = true; continue; }
        running[n_running++] = ps.seq; busy_ctas += ps.ctas;
    }
    /* ---- Kp ---- */
    const uint32_t unparsed = e->n_unparsed;
    if (unparsed && (unparsed >= e->parse_threshold || (idle && n_running == 0))) {
        const ParseScratch &nps = e->pscr[e->next_pscr];
        bool go = !nps.used || nps.finished;
        uint32_t take = unparsed > 8192 ? 8192 : unparsed;
        if (go && e->kp_sms) {
            const uint32_t free_ctas = busy_ctas < e->kp_sms ? e->kp_sms - busy_ctas : 0;
            const uint32_t min_ctas = (e->parse_threshold + 31) / 32;     /* a launch worth its latency */
            if (free_ctas >= (take + 31) / 32) {
                /* every CTA full: a launch of 551 pictures would own 18 SMs for what 17.2 can do (the rest joins the next launch) */
                if (take >= e->parse_threshold && take >= 64 && !idle) take -= take % 32;
            }
            else if (free_ctas >= min_ctas || (n_running == 0 && free_ctas)) take = free_ctas * 32;
            else go = false;
        }
        if (go && take) {
            std::vector<PicBuf *> &pl = e->tmp_parse; pl.clear();
            for (uint32_t level = 0; pl.size() < take; level++) {
                bool any = false;
                for (Inst *in : e->insts) {
                    if (!in->dev_parse) continue;
                    uint32_t k = 0;
                    for (PicBuf *p : *in->fifo) if (!p->parse_seq) { if (k == level) { if (pl.size() < take) pl.push_back(p); any = true; break; } k++; }
                }
                if (!any) break;
            }
            const uint32_t n = (uint32_t)pl.size();
            if (n) {
                if (launch_parse(e, pl)) { for (PicBuf *p : pl) p->inst->slot_flags[p->in.cur_slot] |= 4; }
                else {
                    if (kp_pics) *kp_pics = n;
                    running[n_running < NPAR ? n_running++ : NPAR - 1] = e->parse_seq;
                }
            }
        }
    }
    /* ---- round ---- */
    uint32_t in_flight = 0;
    for (int k = 0; k < NSCR; k++) {
        Scratch &q = e->scr[k];
        if (!q.used || q.done_seq == q.seq) continue;
        if (cudaEventQuery(q.d2h_done) != cudaSuccess) in_flight++;
        else __atomic_store_n(&q.done_seq, q.seq, __ATOMIC_RELEASE);      /* what be_frame_state reports to the polling workers */
    }
    if (in_flight >= (e->split_rounds ? 2u * DRIVE_ROUNDS : (uint32_t)DRIVE_ROUNDS)) return 0;
    std::vector<PicBuf *> &rl = e->tmp_round; rl.clear();
    uint32_t nonempty = 0, miss[3] = {0, 0, 0};
    for (Inst *in : e->insts) {
        if (in->fifo->empty()) continue;
        nonempty++;
        PicBuf *p = in->fifo->front();
        if (in->dev_parse) {
            if (!p->parse_seq) { miss[0]++; continue; }
            bool busy = false;
            for (int k = 0; k < n_running; k++) if (running[k] == p->parse_seq) busy = true;
            if (busy) { miss[1]++; continue; }
        }
        if (gated(p)) { miss[2]++; continue; }
        rl.push_back(p);
    }
    if (rl.empty()) return 0;
    if ((uint32_t)rl.size() * 8 < nonempty * 7 && !(idle && n_running == 0 && in_flight == 0)) return 0;
    for (PicBuf *p : rl) { p->inst->fifo->pop_front(); p->inst->n_pending.fetch_sub(1, std::memory_order_release); }
    e->drv_miss[0] += miss[0]; e->drv_miss[1] += miss[1]; e->drv_miss[2] += miss[2]; e->drv_miss[3]++;
    if (e->split_rounds && defer2 && rl.size() >= 64 && !(e->flags & (H264B200_ENGINE_RETAIN | H264B200_ENGINE_TAP_PREDEBLOCK))) {
        /* experiment: the round as two halves over disjoint instances on two streams */
        std::vector<PicBuf *> a, b2;
        for (PicBuf *p : rl) (p->inst->lane_pref ? b2 : a).push_back(p);
        uint32_t n = 0;
        if (!a.empty()) n += launch_round(e, a, defer, 0);
        if (!b2.empty()) n += launch_round(e, b2, defer2, 1);
        return n;
    }
    return launch_round(e, rl, defer);
}

/* ------------------------------------------------------- backend callbacks */
static void inst_free(Inst *in);
/* engine mutex held: the look-ahead the runner may use is what the smallest instance can hold */
static void note_window(h264b200_engine *e, const Inst *in)
{
    if (!in->dev_parse || (uint32_t)(in->n_bufs - RING_EXTRA) >= e->eff_window) return;
    e->eff_window = (uint32_t)(in->n_bufs - RING_EXTRA);
    const uint32_t thr = (e->n_inst_hint ? e->n_inst_hint : 1) * (e->eff_window >= 4 ? e->eff_window / 2 : 1);
    if (thr < e->parse_threshold) e->parse_threshold = thr;
}
static void *be_inst_create_ex(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n_slots, int host_parse)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx;
    if (n_slots > H264_MAX_SLOTS) return NULL;
    set_device(e);
    /* host_parse: an instance of a device-parse engine that brings its own records (h264b200SetHostParse); its pictures
     * join the same reconstruction rounds, they just have nothing for kernel Kp to do */
    const int dev_parse = (e->flags & H264B200_ENGINE_DEVICE_PARSE) != 0 && !host_parse;
    int n_bufs = dev_parse ? (int)e->window + RING_EXTRA : host_parse ? NBUF + 1 : NBUF;   /* host share of a device-parse run: one picture queued ahead of the round in flight */
    if (dev_parse && e->inst_budget) {
        /* the look-ahead window is a wish: 4K pictures cost 35 MB of worst-case parse output each, and hundreds of
         * instances must fit the device together */
        const size_t per_buf = parse_bytes_per_buf(wm * hm) + block_cap0(wm * hm);
        const size_t fit = e->inst_budget / per_buf;
        if ((size_t)n_bufs > fit) n_bufs = fit < RING_EXTRA + 2 ? RING_EXTRA + 2 : (int)fit;
    }
    {   /* reuse a pooled instance of the same geometry */
        std::lock_guard<std::mutex> lk(e->mu);
        for (size_t i = 0; i < e->pool.size(); i++) {
            Inst *c = e->pool[i];
            if (c->wm == wm && c->hm == hm && c->n_slots == n_slots && c->dev_parse == dev_parse && (dev_parse ? c->n_bufs >= RING_EXTRA + 1 : c->n_bufs == n_bufs)) {
                note_window(e, c);
                e->pool.erase(e->pool.begin() + i);
                memset(c->slot_flags, 0, sizeof c->slot_flags); memset(c->slot_qgen, 0, sizeof c->slot_qgen); memset(c->slot_lgen, 0, sizeof c->slot_lgen);
                memset(c->slot_popped, 0, sizeof c->slot_popped);
                for (int k = 0; k < H264_MAX_SLOTS; k++) c->slot_released[k].store(0);
                c->next_buf = 0; c->out_format = H264B200_OUT_I420;
                c->batched = (e->flags & H264B200_ENGINE_BATCHED) != 0;
                for (int k = 0; k < c->n_bufs; k++) { c->bufs[k].state = 0; c->bufs[k].parse_seq = 0; c->bufs[k].tape_parse = c->bufs[k].tape_last_round = -1; }
                /* a new stream must not start on the previous stream's samples (they would show through MISSING
                 * macroblocks and through frame_host of a slot nothing was decoded into) */
                cudaMemsetAsync(c->d_frames, 0, c->frame_stride * c->n_slots, e->s_comp);
                memset(c->h_frames, 0, c->frame_stride * c->n_slots);
                if (c->d_rgba) { cudaMemsetAsync(c->d_rgba, 0, c->rgba_bytes * c->n_slots, e->s_comp); memset(c->h_rgba, 0, c->rgba_bytes * c->n_slots); }
                if (e->split_rounds) cudaStreamSynchronize(e->s_comp);      /* the instance's first picture may run on the other lane */
                c->last_done = nullptr; c->last_lane = 0; c->lane_pref = (int)(e->next_lane++ & 1);
                e->insts.push_back(c);
                return c;
            }
        }
    }
    Inst *in = new Inst();
    in->e = e; in->wm = wm; in->hm = hm; in->n_mbs = wm * hm; in->n_slots = n_slots;
    in->frame_bytes = (size_t)in->n_mbs * 384;
    in->frame_stride = in->frame_bytes + STAT_TAIL;
    in->batched = (e->flags & H264B200_ENGINE_BATCHED) != 0;
    in->dev_parse = dev_parse; in->n_bufs = n_bufs;
    in->last_done = nullptr; in->last_lane = 0; in->lane_pref = 0;
    in->fifo = new std::deque<PicBuf *>();
    in->bufs = (PicBuf *)calloc((size_t)n_bufs, sizeof(PicBuf));
    if (!in->bufs) { inst_free(in); return NULL; }
    CUDA_TRY(cudaMalloc((void **)&in->d_frames, in->frame_stride * n_slots), { inst_free(in); return NULL; });
    CUDA_TRY(cudaMemset(in->d_frames, 0, in->frame_stride * n_slots), { inst_free(in); return NULL; });
    CUDA_TRY(cudaHostAlloc((void **)&in->h_frames, in->frame_stride * n_slots, cudaHostAllocDefault), { inst_free(in); return NULL; });
    memset(in->h_frames, 0, in->frame_stride * n_slots);
    if (!dev_parse) {
        for (int i = 0; i < n_bufs; i++) if (picbuf_alloc_host(&in->bufs[i], in, in->n_mbs * 10 + 64)) { inst_free(in); return NULL; }   /* frees what was allocated so far */
    } else {
        /* three allocations per instance: pinned blocks, their device twins, and what kernel Kp writes */
        const size_t pb = parse_bytes_per_buf(in->n_mbs), cap0 = block_cap0(in->n_mbs);
        const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t), ctx_bytes = (size_t)in->n_mbs * sizeof(KpMbCtx);
        CUDA_TRY(cudaHostAlloc((void **)&in->h_blocks, cap0 * n_bufs, cudaHostAllocDefault), { inst_free(in); return NULL; });
        CUDA_TRY(cudaMalloc((void **)&in->d_blocks, cap0 * n_bufs), { inst_free(in); return NULL; });
        CUDA_TRY(cudaMalloc((void **)&in->d_parse, pb * n_bufs), { inst_free(in); return NULL; });
        for (int i = 0; i < n_bufs; i++) {
            PicBuf *p = &in->bufs[i];
            uint8_t *base = in->d_parse + pb * i;
            p->inst = in; p->tape_parse = p->tape_last_round = -1;
            p->in.block = in->h_blocks + cap0 * i; p->in.block_cap = (uint32_t)cap0;
            p->d_block = in->d_blocks + cap0 * i; p->d_block_cap = (uint32_t)cap0;
            p->d_mbs = (h264b200_mb_t *)base;
            p->d_ctx = (KpMbCtx *)(base + rec_bytes);
            p->d_res = (KpResult *)(base + rec_bytes + ctx_bytes);
            p->d_coef = (int16_t *)(base + rec_bytes + ctx_bytes + 256);
            p->d_coef_cap = KP_COEF_CAP(in->n_mbs);
            p->in.priv = p;
        }
    }
    std::lock_guard<std::mutex> lk(e->mu);
    note_window(e, in);
    in->lane_pref = (int)(e->next_lane++ & 1);
    e->insts.push_back(in);
    return in;
}

static void *be_inst_create(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n_slots) { return be_inst_create_ex(be, wm, hm, n_slots, 0); }

static void inst_free(Inst *in)
{
    if (in->bufs) { for (int i = 0; i < in->n_bufs; i++) picbuf_free(&in->bufs[i]); free(in->bufs); }
    if (in->h_blocks) cudaFreeHost(in->h_blocks);
    if (in->d_blocks) cudaFree(in->d_blocks);
    if (in->d_parse) cudaFree(in->d_parse);
    if (in->d_frames) cudaFree(in->d_frames);
    if (in->h_frames) cudaFreeHost(in->h_frames);
    if (in->d_rgba) { cudaFree(in->d_rgba); cudaFreeHost(in->h_rgba); }
    if (in->d_pre) cudaFree(in->d_pre);
    delete in->fifo;
    delete in;
}

static void be_inst_destroy(h264_backend_t *be, void *inst)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    bool keep;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        while (!in->fifo->empty()) {
            if (!advance_all_locked(e) && !in->fifo->empty()) {
                /* held back by an output nobody will release any more */
                for (int k = 0; k < H264_MAX_SLOTS; k++) in->slot_released[k].store(in->slot_popped[k]);
                if (!advance_all_locked(e)) break;
            }
        }
        while (!in->fifo->empty()) { if (in->dev_parse && !in->fifo->front()->parse_seq && e->n_unparsed) e->n_unparsed--; in->fifo->front()->state = 0; in->fifo->pop_front(); }
        in->n_pending.store(0);
        keep = !e->retained.empty();           /* retained batches name this instance's frame pool and parse buffers */
        if (keep) e->zombies.push_back(in);
        for (size_t i = 0; i < e->insts.size(); i++) if (e->insts[i] == in) { e->insts.erase(e->insts.begin() + i); break; }
    }
    set_device(e);
    /* a round of this instance may have been launched by the scheduling thread with its copy-out still to be issued: the
     * stream synchronisation below must see those copies, or they would land in a pooled instance's new life */
    while (e->copyouts_deferred.load() != 0) { struct timespec ts = {0, 50000}; nanosleep(&ts, NULL); }
    for (int k = 0; k < NPAR; k++) cudaStreamSynchronize(e->s_parse[k]);
    cudaStreamSynchronize(e->s_comp); cudaStreamSynchronize(e->s_comp2); cudaStreamSynchronize(e->s_d2h);
    if (!keep) { std::lock_guard<std::mutex> lk(e->mu); e->pool.push_back(in); }
}

static h264_pic_input_t *be_pic_begin(h264_backend_t *be, void *inst)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    PicBuf *p = &in->bufs[in->next_buf];
    in->next_buf = (in->next_buf + 1) % in->n_bufs;
    if (p->state == 2) {           /* the look-ahead ran a whole ring ahead of the launches */
        std::lock_guard<std::mutex> lk(e->mu);
        while (p->state == 2) {
            if (!advance_all_locked(e) && p->state == 2) {
                for (int k = 0; k < H264_MAX_SLOTS; k++) in->slot_released[k].store(in->slot_popped[k]);
                if (!advance_all_locked(e)) return NULL;
            }
        }
    }
    if (p->state == 3) { set_device(e); cudaEventSynchronize(p->done); }
    p->state = 1; p->parse_seq = 0; p->block_on_device = 0;
    return &p->in;
}

static int be_coef_grow(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_slots)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
    uint32_t cap = pic->coef_cap * 2 > min_slots ? pic->coef_cap * 2 : min_slots;
    uint8_t *n = NULL;
    if (in->dev_parse) return -1;
    set_device(e);
    CUDA_TRY(cudaHostAlloc((void **)&n, rec_bytes + (size_t)cap * 32, cudaHostAllocDefault), return -1);
    memcpy(n, pic->mbs, rec_bytes + (size_t)pic->coef_used * 32);           /* records written so far and their slots */
    cudaFreeHost(pic->mbs);
    pic->mbs = (h264b200_mb_t *)n; pic->coef = (int16_t *)(n + rec_bytes); pic->coef_cap = cap;
    return 0;
}

/* device-parse: room for a longer block (the pinned side; the device twin follows at the Kp launch) */
static int be_block_grow(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_bytes)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    PicBuf *p = (PicBuf *)pic->priv;
    if (!in->dev_parse) return -1;
    uint32_t cap = pic->block_cap * 2 > min_bytes ? pic->block_cap * 2 : min_bytes;
    cap = (cap + 4095u) & ~4095u;
    uint8_t *n = NULL;
    set_device(e);
    CUDA_TRY(cudaHostAlloc((void **)&n, cap, cudaHostAllocDefault), return -1);
    if (pic->block && pic->block_used) memcpy(n, pic->block, pic->block_used);
    if (p->own_host && pic->block) cudaFreeHost(pic->block);
    pic->block = n; pic->block_cap = cap; p->own_host = 1;
    return 0;
}

static int be_pic_submit(h264_backend_t *be, void *inst, h264_pic_input_t *pic)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    PicBuf *p = (PicBuf *)pic->priv;
    if (e->copy_at_submit && in->dev_parse && pic->block && !(e->flags & H264B200_ENGINE_RETAIN) && p->d_block_cap >= pic->block_used) {
        /* the slices travel now, from the thread that scanned them, before the engine mutex is taken: a Kp launch then
         * finds its input on the device (the copy precedes, in s_h2d, the event the launch will wait for) and the
         * scheduling thread is spared one copy call per picture */
        set_device(e);
        if (cudaMemcpyAsync(p->d_block, pic->block, pic->block_used, cudaMemcpyHostToDevice, e->s_h2d) == cudaSuccess) p->block_on_device = 1;
    }
    std::lock_guard<std::mutex> lk(e->mu);
    if (in->dev_parse) e->n_unparsed++;
    p->state = 2;
    in->slot_qgen[pic->cur_slot]++;
    /* whatever the caller still holds of this frame slot (h264b200NextOutputPictureAsync) must be released before the
     * picture may overwrite the slot's host mirror; pops of a slot always precede its re-use in host order */
    p->gate_gen = in->slot_popped[pic->cur_slot];
    in->fifo->push_back(p);
    in->n_pending.fetch_add(1, std::memory_order_release);
    if (!in->batched) advance_all_locked(e);
    return 0;
}

/* Wait until generation `gen` of `slot` is in its host mirror.  Returns 0; 1 when a LATER picture has already been
 * launched into the slot (its mirror may be overwritten: the caller waited too long); 2 when the picture is still queued
 * (only without may_advance); -1 on a CUDA error. */
static int wait_slot(h264b200_engine *e, Inst *in, int slot, uint32_t gen, bool may_advance)
{
    if ((int32_t)(in->slot_lgen[slot] - gen) < 0) {
        if (!may_advance) return 2;
        std::lock_guard<std::mutex> lk(e->mu);
        while ((int32_t)(in->slot_lgen[slot] - gen) < 0) {
            if (!advance_locked(e, true)) {
                /* nothing can be launched: the picture is held back by an unreleased output of this very caller */
                for (int k = 0; k < H264_MAX_SLOTS; k++) in->slot_released[k].store(in->slot_popped[k]);
                if (!advance_locked(e, true)) break;
            }
        }
    }
    if (in->slot_lgen[slot] != gen) return (int32_t)(in->slot_lgen[slot] - gen) < 0 ? 2 : 1;
    if (in->slot_flags[slot] & 4) return -1;
    if (slot_copy_pending(e, in, slot)) {
        set_device(e);
        cudaError_t er = cudaEventSynchronize(in->slot_ready[slot]);
        if (er != cudaSuccess) { fprintf(stderr, "h264b200: reconstruction failed: %s\n", cudaGetErrorString(er)); return -1; }
    }
    return 0;
}

static uint8_t *slot_mirror(Inst *in, int slot)
{
    if (in->out_format == H264B200_OUT_RGBA && in->h_rgba) return in->h_rgba + (size_t)slot * in->rgba_bytes;
    return in->h_frames + (size_t)slot * in->frame_stride;
}

static uint8_t *be_frame_host(h264_backend_t *be, void *inst, int slot, uint32_t *error_flags)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (slot < 0 || slot >= (int)in->n_slots) return NULL;
    if (wait_slot(e, in, slot, in->slot_qgen[slot], true) < 0) return NULL;
    if (error_flags) *error_flags = *e->h_err;
    return slot_mirror(in, slot);
}

/* non-blocking: where the newest picture of `slot` will be, and its generation (for frame_wait) */
static uint8_t *be_frame_host_async(h264_backend_t *be, void *inst, int slot, uint32_t *gen)
{
    Inst *in = (Inst *)inst; (void)be;
    if (slot < 0 || slot >= (int)in->n_slots) return NULL;
    if (gen) *gen = in->slot_qgen[slot];
    in->slot_popped[slot] = in->slot_qgen[slot];
    return slot_mirror(in, slot);
}

/* the ticket carries 24 bits of the generation: take the value nearest below the current one */
static uint32_t full_gen(const Inst *in, int slot, uint32_t gen)
{
    uint32_t full = (in->slot_qgen[slot] & 0xff000000u) | (gen & 0xffffffu);
    if (full > in->slot_qgen[slot]) full -= 1u << 24;
    return full;
}

static int be_frame_wait(h264_backend_t *be, void *inst, int slot, uint32_t gen, uint32_t *error_flags)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (slot < 0 || slot >= (int)in->n_slots) return -1;
    /* a batched device-parse instance is driven by h264b200EngineAdvance: never launch from here */
    int rc = wait_slot(e, in, slot, full_gen(in, slot, gen), !(in->dev_parse && in->batched));
    if (error_flags) *error_flags = *e->h_err;
    return rc;
}

static int be_frame_state(h264_backend_t *be, void *inst, int slot, uint32_t gen)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (slot < 0 || slot >= (int)in->n_slots) return -1;
    /* No mutex: workers poll this for every stream while the scheduling thread holds the mutex to launch.  The launching
     * thread publishes slot_ready / slot_flags BEFORE it advances slot_lgen (release); a later picture cannot be launched
     * into the slot before the caller released this one, so what is read behind the acquire belongs to generation g. */
    const uint32_t g = full_gen(in, slot, gen);
    const uint32_t lg = __atomic_load_n(&in->slot_lgen[slot], __ATOMIC_ACQUIRE);
    if ((int32_t)(lg - g) < 0) return 2;
    if (lg != g) return 0;                              /* a later picture was launched into the slot: h264b200PictureWait reports that */
    if (in->slot_flags[slot] & 4) return -1;
    if (slot_copy_pending(e, in, slot)) {
        /* While a scheduling thread drives the engine it polls the copy-out events anyway and publishes what it sees;
         * fifteen workers asking the runtime about every picture of every stream on every sweep (850 000 cudaEventQuery
         * per second) slowed every other CUDA call of the process down, the launches included. */
        if (__atomic_load_n(&e->scr[in->slot_scr[slot]].done_seq, __ATOMIC_ACQUIRE) == in->slot_rseq[slot]) return 0;
        if (e->last_drive_ms > 0 && host_ms_now() - e->last_drive_ms < 20.0) return 1;
        set_device(e);
        return cudaEventQuery(in->slot_ready[slot]) == cudaSuccess ? 0 : 1;
    }
    return 0;
}

static void be_frame_release(h264_backend_t *be, void *inst, int slot, uint32_t gen)
{
    Inst *in = (Inst *)inst; (void)be;
    if (slot < 0 || slot >= (int)in->n_slots) return;
    const uint32_t g = full_gen(in, slot, gen);
    if ((int32_t)(g - in->slot_released[slot].load(std::memory_order_relaxed)) > 0) in->slot_released[slot].store(g, std::memory_order_release);
}

static int be_frame_status(h264_backend_t *be, void *inst, int slot, h264b200_picstat_t *out)
{
    Inst *in = (Inst *)inst; (void)be;
    if (slot < 0 || slot >= (int)in->n_slots || !out) return -1;
    memcpy(out, in->h_frames + (size_t)slot * in->frame_stride + in->frame_bytes, sizeof *out);
    return 0;
}

static uint32_t be_inst_pending(h264_backend_t *be, void *inst) { (void)be; return ((Inst *)inst)->n_pending.load(std::memory_order_acquire); }

/* output format of an instance: 0 ok.  RGBA buffers are allocated on first use. */
static int be_set_output(h264_backend_t *be, void *inst, int format, int cl, int ct, int cw, int ch)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (format != H264B200_OUT_I420 && format != H264B200_OUT_RGBA) return -1;
    if (format == H264B200_OUT_RGBA) {
        if (cw <= 0 || ch <= 0 || cl < 0 || ct < 0 || cl + cw > (int)in->wm * 16 || ct + ch > (int)in->hm * 16 || (cl & 1) || (ct & 1)) return -1;
        const size_t bytes = (size_t)cw * ch * 4;
        set_device(e);
        if (in->rgba_bytes != bytes) {
            if (in->d_rgba) { cudaStreamSynchronize(e->s_d2h); cudaFree(in->d_rgba); cudaFreeHost(in->h_rgba); in->d_rgba = in->h_rgba = NULL; }
            CUDA_TRY(cudaMalloc((void **)&in->d_rgba, bytes * in->n_slots), return -1);
            CUDA_TRY(cudaHostAlloc((void **)&in->h_rgba, bytes * in->n_slots, cudaHostAllocDefault), return -1);
            memset(in->h_rgba, 0, bytes * in->n_slots);
            in->rgba_bytes = bytes;
        }
        in->cl = cl; in->ct = ct; in->cw = cw; in->ch = ch;
    }
    in->out_format = format;
    return 0;
}

static void be_destroy(h264_backend_t *be) { (void)be; }

/* ------------------------------------------------------------- engine API */
extern "C" int h264b200Probe(char *msg, size_t cap)
{
    int n = 0;
    cudaError_t er = cudaGetDeviceCount(&n);
    if (er != cudaSuccess || n == 0) {
        if (msg && cap) snprintf(msg, cap, "no usable CUDA device: %s", er != cudaSuccess ? cudaGetErrorString(er) : "device count is 0");
        return -1;
    }
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    if (msg && cap) snprintf(msg, cap, "%d device(s); device 0: %s sm_%d%d, %d SMs", n, p.name, p.major, p.minor, p.multiProcessorCount);
    if (p.major != 10) { if (msg && cap) snprintf(msg, cap, "device 0 is sm_%d%d; this library is built for sm_100a only", p.major, p.minor); return -2; }
    return 0;
}

extern "C" h264b200_engine_t *h264b200EngineCreateEx(int device, uint32_t flags)
{
    char msg[256];
    if (h264b200Probe(msg, sizeof msg)) { fprintf(stderr, "h264b200: %s\n", msg); return NULL; }
    if (device < 0) { if (cudaGetDevice(&device) != cudaSuccess) device = 0; }
    h264b200_engine *e = new h264b200_engine();
    e->device = device; e->flags = flags; e->next_scr = 0; e->next_pscr = 0; e->parse_seq = 0; e->n_unparsed = 0; e->round_seq = 0; e->last_drive_ms = 0; e->copyouts_deferred.store(0);
    e->drv_locked_ms = e->drv_copy_ms = e->drv_locked_max = 0; e->drv_polls = e->drv_launches = 0; memset(e->drv_miss, 0, sizeof e->drv_miss);
    { const char *c = getenv("H264B200_COPY_AT_SUBMIT"); e->copy_at_submit = !(c && atoi(c) == 0); }
    e->window = 1; e->eff_window = 1; e->parse_threshold = 1; e->n_inst_hint = 0; e->inst_budget = 0;
    memset(&e->st, 0, sizeof e->st); memset(e->scr, 0, sizeof e->scr); memset(e->pscr, 0, sizeof e->pscr);
    memset(e->k_ms, 0, sizeof e->k_ms); memset(e->k_bytes, 0, sizeof e->k_bytes); memset(e->k_launches, 0, sizeof e->k_launches);
    CUDA_TRY(cudaSetDevice(device), { delete e; return NULL; });
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device), { delete e; return NULL; });
    e->sm_count = p.multiProcessorCount;
    e->wf_cap = 16;
    { const char *c = getenv("H264B200_KP_ON_COMP"); e->kp_on_comp = c && atoi(c) > 0; }
    {   /* exclusive Kp launches for batched device-parse engines: three quarters of the SMs (112 of 148: the share of Kp in the
         * device work of the default workload, 16.5 of 22 ms per 256 pictures); H264B200_KP_SMS overrides, 0 = shared mode */
        const char *c = getenv("H264B200_KP_SMS");
        e->kp_sms = (flags & H264B200_ENGINE_BATCHED) && (flags & H264B200_ENGINE_DEVICE_PARSE) ? (uint32_t)(e->sm_count * 3 / 4 + 1) : 0;
        if (c) e->kp_sms = (uint32_t)atoi(c) < (uint32_t)e->sm_count ? (uint32_t)atoi(c) : (uint32_t)e->sm_count;
        CUDA_TRY(cudaFuncSetAttribute(kp_parse<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KP_SMEM_BYTES(32)), { delete e; return NULL; });
        CUDA_TRY(cudaFuncSetAttribute(kp_parse<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)KP_SMEM_BYTES(8)), { delete e; return NULL; });
    }
    e->tl_path = getenv("H264B200_TIMELINE");
    if (e->tl_path && !*e->tl_path) e->tl_path = nullptr;
    if (e->tl_path) { cudaEventCreate(&e->tl_base); cudaEventRecord(e->tl_base, 0); e->tl_host0 = host_ms_now(); }
    { const char *c = getenv("H264B200_WF_CAP"); if (c && atoi(c) > 0 && atoi(c) <= 64) e->wf_cap = (uint32_t)atoi(c); }
    CUDA_TRY(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking), { delete e; return NULL; });
    {   /* The reconstruction stream gets the highest priority, the Kp streams the lowest: a Kp launch lives for ~0.2 s in
         * persistent CTAs, a reconstruction round for a few milliseconds in thousands of short ones; whenever an SM has room
         * the block scheduler should give it to the round that a stream's next picture (and the host's next output) waits for.
         * H264B200_PRIO=0 creates every stream at the default priority. */
        int lo = 0, hi = 0;
        const char *pe = getenv("H264B200_PRIO");
        const bool prio = !(pe && atoi(pe) == 0);
        cudaDeviceGetStreamPriorityRange(&lo, &hi);         /* lo: numerically largest = least urgent */
        CUDA_TRY(cudaStreamCreateWithPriority(&e->s_comp, cudaStreamNonBlocking, prio ? hi : 0), { delete e; return NULL; });
        CUDA_TRY(cudaStreamCreateWithPriority(&e->s_d2h, cudaStreamNonBlocking, prio ? hi : 0), { delete e; return NULL; });
        CUDA_TRY(cudaStreamCreateWithPriority(&e->s_comp2, cudaStreamNonBlocking, prio ? hi : 0), { delete e; return NULL; });
        for (int k = 0; k < NPAR; k++) CUDA_TRY(cudaStreamCreateWithPriority(&e->s_parse[k], cudaStreamNonBlocking, prio ? lo : 0), { delete e; return NULL; });
    }
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_h2d, cudaEventDisableTiming), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_comp, cudaEventDisableTiming), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_comp2, cudaEventDisableTiming), { delete e; return NULL; });
    { const char *c = getenv("H264B200_SPLIT_ROUNDS"); e->split_rounds = c && atoi(c) > 0; e->next_lane = 0; }
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_gate, cudaEventDisableTiming), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreate(&e->ev_rep0), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreate(&e->ev_rep1), { delete e; return NULL; });
    for (int i = 0; i < NSCR; i++) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->scr[i].done, cudaEventDisableTiming), { delete e; return NULL; });
        CUDA_TRY(cudaEventCreateWithFlags(&e->scr[i].d2h_done, cudaEventDisableTiming), { delete e; return NULL; });
    }
    for (int i = 0; i < NPAR; i++) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->pscr[i].done, cudaEventDisableTiming), { delete e; return NULL; });
        CUDA_TRY(cudaMalloc((void **)&e->pscr[i].d_ticket, 64), { delete e; return NULL; });
        CUDA_TRY(cudaMemset(e->pscr[i].d_ticket, 0, 64), { delete e; return NULL; });
    }
    {   /* the host parser's code tables, for kernel Kp */
        KpTables *t = (KpTables *)malloc(sizeof(KpTables));
        if (!t) { delete e; return NULL; }
        h264_kp_fill_tables(t);
        CUDA_TRY(cudaMalloc((void **)&e->d_tables, sizeof(KpTables)), { free(t); delete e; return NULL; });
        CUDA_TRY(cudaMemcpy(e->d_tables, t, sizeof(KpTables), cudaMemcpyHostToDevice), { free(t); delete e; return NULL; });
        free(t);
    }
    CUDA_TRY(cudaMalloc((void **)&e->d_err, 64), { delete e; return NULL; });
    CUDA_TRY(cudaMemset(e->d_err, 0, 64), { delete e; return NULL; });
    CUDA_TRY(cudaHostAlloc((void **)&e->h_err, 64, cudaHostAllocDefault), { delete e; return NULL; });
    *e->h_err = 0;
    e->d_trace = nullptr; e->trace_left = 0;
    if (getenv("H264B200_TRACE")) {
        e->trace_left = atoi(getenv("H264B200_TRACE"));
        CUDA_TRY(cudaMalloc((void **)&e->d_trace, (256 + 4 * 512) * 8), { delete e; return NULL; });
        cudaMemset(e->d_trace, 0, (256 + 4 * 512) * 8);
    }
    memset(&e->be, 0, sizeof e->be);
    e->be.inst_create = be_inst_create; e->be.inst_destroy = be_inst_destroy; e->be.pic_begin = be_pic_begin;
    e->be.coef_grow = be_coef_grow; e->be.pic_submit = be_pic_submit; e->be.frame_host = be_frame_host;
    e->be.frame_host_async = be_frame_host_async;
    e->be.frame_wait = be_frame_wait;
    e->be.set_output = be_set_output;
    e->be.block_grow = be_block_grow; e->be.frame_status = be_frame_status; e->be.frame_release = be_frame_release;
    e->be.inst_pending = be_inst_pending;
    e->be.frame_state = be_frame_state; e->be.inst_create_ex = be_inst_create_ex;
    e->be.parse_mode = (flags & H264B200_ENGINE_DEVICE_PARSE) != 0;
    e->be.destroy = be_destroy; e->be.ctx = e;
    return e;
}
extern "C" h264b200_engine_t *h264b200EngineCreate(int device) { return h264b200EngineCreateEx(device, H264B200_ENGINE_BATCHED); }

static void free_retained(h264b200_engine *e)
{
    for (Retained *r : e->retained) { for (void *p : r->owned) cudaFree(p); cudaEventDestroy(r->ev); delete r; }
    e->retained.clear();
    for (Inst *z : e->zombies) inst_free(z);
    e->zombies.clear();
    for (Inst *in : e->insts) for (int k = 0; k < in->n_bufs; k++) in->bufs[k].tape_parse = in->bufs[k].tape_last_round = -1;
    for (cudaEvent_t ev : e->tev) cudaEventDestroy(ev);
    e->tev.clear(); e->tev_entry.clear();
}

extern "C" void h264b200EngineDestroy(h264b200_engine_t *e)
{
    if (!e) return;
    set_device(e);
    cudaStreamSynchronize(e->s_h2d); for (int k = 0; k < NPAR; k++) cudaStreamSynchronize(e->s_parse[k]);
    cudaStreamSynchronize(e->s_comp); cudaStreamSynchronize(e->s_comp2); cudaStreamSynchronize(e->s_d2h);
    tl_dump(e);
    free_retained(e);
    for (Inst *p : e->pool) inst_free(p);
    e->pool.clear();
    for (int i = 0; i < NSCR; i++) {
        Scratch &s = e->scr[i];
        if (s.h_jobs) cudaFreeHost(s.h_jobs);
        if (s.d_jobs) cudaFree(s.d_jobs);
        if (s.d_ctrl) cudaFree(s.d_ctrl);
        cudaEventDestroy(s.done); cudaEventDestroy(s.d2h_done);
    }
    for (int i = 0; i < NPAR; i++) {
        ParseScratch &s = e->pscr[i];
        if (s.h_pics) cudaFreeHost(s.h_pics);
        if (s.d_pics) cudaFree(s.d_pics);
        if (s.d_ticket) cudaFree(s.d_ticket);
        cudaEventDestroy(s.done);
    }
    cudaFree(e->d_tables);
    cudaFree(e->d_err); cudaFreeHost(e->h_err);
    if (e->d_trace) cudaFree(e->d_trace);
    cudaEventDestroy(e->ev_h2d); cudaEventDestroy(e->ev_comp); cudaEventDestroy(e->ev_rep0); cudaEventDestroy(e->ev_rep1); cudaEventDestroy(e->ev_gate);
    cudaStreamDestroy(e->s_h2d); cudaStreamDestroy(e->s_comp); cudaStreamDestroy(e->s_comp2); cudaStreamDestroy(e->s_d2h); cudaEventDestroy(e->ev_comp2);
    for (int k = 0; k < NPAR; k++) cudaStreamDestroy(e->s_parse[k]);
    delete e;
}

extern "C" u32 h264_decoder_create(storage_t *pStorage, u32 noOutputReordering, h264_backend_t *be);
extern "C" u32 h264b200InitOnEngine(storage_t *pStorage, u32 noOutputReordering, h264b200_engine_t *e)
{
    if (!e) return HANTRO_NOK;
    return h264_decoder_create(pStorage, noOutputReordering, &e->be);
}

extern "C" u32 h264b200EngineSubmit(h264b200_engine_t *e)
{
    if (!e) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    return advance_all_locked(e);
}
extern "C" u32 h264b200EngineAdvance(h264b200_engine_t *e)
{
    if (!e) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    return advance_locked(e, false);
}
extern "C" u32 h264b200EngineDrive(h264b200_engine_t *e, int idle, u32 *kp_pictures)
{
    if (!e) return 0;
    CopyOut co, co2; co.d2h_done = nullptr; co2.d2h_done = nullptr;
    u32 n, kp = 0;
    const double t0 = host_ms_now();
    e->last_drive_ms = t0;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        n = drive_locked(e, idle, &kp, &co, &co2);
    }
    const double t1 = e->tl_path ? host_ms_now() : 0;
    if (!co.list.empty()) { copy_out_issue(e, co); e->copyouts_deferred.fetch_sub(1); }
    if (!co2.list.empty()) { copy_out_issue(e, co2); e->copyouts_deferred.fetch_sub(1); }
    if (e->tl_path) {
        const double t2 = host_ms_now();
        e->drv_polls++;
        if (n || kp) { e->drv_launches++; e->drv_locked_ms += t1 - t0; if (t1 - t0 > e->drv_locked_max) e->drv_locked_max = t1 - t0; e->drv_copy_ms += t2 - t1; }
    }
    if (kp_pictures) *kp_pictures = kp;
    return n;
}
extern "C" void h264b200EngineSetWindow(h264b200_engine_t *e, uint32_t depth, uint32_t parse_threshold)
{
    if (!e) return;
    std::lock_guard<std::mutex> lk(e->mu);
    e->window = depth ? (depth > 64 ? 64 : depth) : 1;
    e->eff_window = e->window;
    e->parse_threshold = parse_threshold ? parse_threshold : 1;
}
extern "C" uint32_t h264b200EngineWindow(h264b200_engine_t *e) { if (!e) return 0; std::lock_guard<std::mutex> lk(e->mu); return e->eff_window; }
/* How many instances are going to share the engine: each may spend an equal part of half of the device memory on its
 * look-ahead buffers (the rest stays for frame pools and scratch); instances created afterwards shrink their window to that. */
extern "C" void h264b200EngineSetStreams(h264b200_engine_t *e, uint32_t n_streams)
{
    if (!e || !n_streams) return;
    set_device(e);
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return;
    std::lock_guard<std::mutex> lk(e->mu);
    e->n_inst_hint = n_streams;
    e->inst_budget = (size_t)((double)tot * 0.5) / n_streams;      /* of the TOTAL: pooled instances of an earlier run already hold their share */
}
/* pictures one Kp launch parses at full rate: one per warp of the SMs it owns (0: no such limit) */
extern "C" uint32_t h264b200EngineParseSlots(h264b200_engine_t *e) { return e ? e->kp_sms * 32u : 0; }

extern "C" void h264b200EngineSync(h264b200_engine_t *e)
{
    if (!e) return;
    set_device(e);
    cudaError_t a = cudaStreamSynchronize(e->s_h2d), p0 = cudaSuccess;
    for (int k = 0; k < NPAR; k++) { cudaError_t pk = cudaStreamSynchronize(e->s_parse[k]); if (pk != cudaSuccess) p0 = pk; }
    cudaError_t b = cudaStreamSynchronize(e->s_comp), b2 = cudaStreamSynchronize(e->s_comp2), c = cudaStreamSynchronize(e->s_d2h);
    if (b == cudaSuccess) b = b2;
    if (a != cudaSuccess || b != cudaSuccess || c != cudaSuccess || p0 != cudaSuccess)
        fprintf(stderr, "h264b200: engine sync failed: %s\n", cudaGetErrorString(a != cudaSuccess ? a : p0 != cudaSuccess ? p0 : b != cudaSuccess ? b : c));
}

extern "C" void h264b200EngineStats(h264b200_engine_t *e, h264b200_stats_t *out) { if (e && out) { std::lock_guard<std::mutex> lk(e->mu); *out = e->st; } }
extern "C" u32 h264b200EngineErrorFlags(h264b200_engine_t *e) { return e ? *e->h_err : 0; }
extern "C" void h264b200EngineSetFlags(h264b200_engine_t *e, uint32_t flags)
{
    if (!e) return;
    std::lock_guard<std::mutex> lk(e->mu);
    e->flags = flags; e->be.parse_mode = (flags & H264B200_ENGINE_DEVICE_PARSE) != 0;
}
extern "C" uint32_t h264b200EngineFlags(h264b200_engine_t *e) { if (!e) return 0; std::lock_guard<std::mutex> lk(e->mu); return e->flags; }

/* -------------------------------------------------------- resident replay */
extern "C" void h264b200EngineDropRetained(h264b200_engine_t *e)
{
    if (!e) return;
    h264b200EngineSync(e);
    std::lock_guard<std::mutex> lk(e->mu);
    free_retained(e);
}

/* Re-run the tape: every Kp launch and every reconstruction round of the retained run, in the order and with the
 * dependencies of the live run (a round waits for the Kp launches that produce its records; a Kp launch waits for the
 * last round that read a parse buffer it overwrites), inputs resident in HBM, no host<->device copies. */
extern "C" u32 h264b200EngineReplay(h264b200_engine_t *e, u32 reps, int time_kernels)
{
    if (!e) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    set_device(e);
    u32 pics = 0;
    cudaEventRecord(e->ev_rep0, e->s_comp);
    for (u32 rep = 0; rep < reps; rep++) {
        /* nothing of this repetition starts before the previous one (or the caller's earlier work) has finished */
        cudaEventRecord(e->ev_gate, e->s_comp);
        for (int k = 0; k < NPAR; k++) cudaStreamWaitEvent(e->s_parse[k], e->ev_gate, 0);
        for (size_t bi = 0; bi < e->retained.size(); bi++) {
            Retained *r = e->retained[bi];
            cudaEvent_t *tev = nullptr;
            if (time_kernels) {
                const int nev = r->kind == 0 ? 5 : 2;
                size_t base = e->tev.size();
                for (int k = 0; k < nev; k++) { cudaEvent_t ev; cudaEventCreate(&ev); e->tev.push_back(ev); }
                e->tev_entry.push_back((int)bi);
                tev = &e->tev[base];
            }
            if (r->kind == 1) {
                cudaStream_t s = r->stream < 0 ? e->s_comp : e->s_parse[r->stream];
                if (r->wait_round >= 0) cudaStreamWaitEvent(s, e->retained[(size_t)r->wait_round]->ev, 0);
                cudaMemsetAsync(r->kp.ticket, 0, 64, s);
                if (tev) cudaEventRecord(tev[0], s);
                kp_launch(e, r->kp, s);
                if (tev) cudaEventRecord(tev[1], s);
                cudaEventRecord(r->ev, s);
                e->st.kernel_launches++; e->st.kp_launches++; e->st.kp_pictures += r->kp.n_pics;
            } else {
                for (int w : r->wait_parse) cudaStreamWaitEvent(e->s_comp, e->retained[(size_t)w]->ev, 0);
                /* the round's own control area (tickets + wavefront progress) and job table are part of its retained allocation */
                cudaMemsetAsync(r->batch.tickets, 0, r->ctrl_words * sizeof(int32_t), e->s_comp);
                launch_kernels(e, r->batch, r->pl, tev, e->s_comp);
                cudaEventRecord(r->ev, e->s_comp);
                pics += r->n_pics;
                e->st.batches++;
            }
        }
    }
    cudaEventRecord(e->ev_rep1, e->s_comp);
    e->st.pictures += pics;
    return pics;
}

/* Device time of the last h264b200EngineReplay call: CUDA events on the compute stream around all of its launches. */
extern "C" double h264b200EngineReplayMs(h264b200_engine_t *e)
{
    float ms = 0;
    if (!e) return -1.0;
    set_device(e);
    if (cudaEventSynchronize(e->ev_rep1) != cudaSuccess || cudaEventElapsedTime(&ms, e->ev_rep0, e->ev_rep1) != cudaSuccess) return -1.0;
    return (double)ms;
}

/* Fold the event pairs of timed replays into the per-kernel totals (after a sync).  Families: K1, K2, K3, K4, Kp. */
extern "C" void h264b200EngineKernelTimes(h264b200_engine_t *e, h264b200_kernel_times_t *out, int reset)
{
    if (!e || !out) return;
    h264b200EngineSync(e);
    std::lock_guard<std::mutex> lk(e->mu);
    set_device(e);
    for (Retained *r : e->retained) if (r->kind == 0 && !r->bytes_read) {
        unsigned long long h[4] = {0, 0, 0, 0};
        cudaMemcpy(h, r->d_bytes, sizeof h, cudaMemcpyDeviceToHost);
        for (int k = 0; k < 4; k++) r->bytes[k] = h[k];
        r->bytes_read = true;
    }
    size_t pos = 0;
    for (size_t i = 0; i < e->tev_entry.size(); i++) {
        Retained *r = e->retained[(size_t)e->tev_entry[i]];
        if (r->kind == 0) {
            const bool on[4] = {r->pl.k1, r->pl.k2, r->pl.k3 || r->pl.k3c, r->pl.k4};
            for (int k = 0; k < 4; k++) if (on[k]) {
                float ms = 0; cudaEventElapsedTime(&ms, e->tev[pos + k], e->tev[pos + k + 1]);
                e->k_ms[k] += ms; e->k_bytes[k] += r->bytes[k]; e->k_launches[k]++;
            }
            pos += 5;
        } else {
            /* Kp: the blocks read + the records and coefficient slots written (= half of K1's bytes of the rounds it feeds) */
            float ms = 0; cudaEventElapsedTime(&ms, e->tev[pos], e->tev[pos + 1]);
            uint64_t slots = 0;
            for (int w : r->rounds) slots += e->retained[(size_t)w]->bytes[0] / 2;
            e->k_ms[4] += ms; e->k_bytes[4] += r->kp_in_bytes + r->kp_rec_bytes + slots; e->k_launches[4]++;
            pos += 2;
        }
    }
    for (cudaEvent_t ev : e->tev) cudaEventDestroy(ev);
    e->tev.clear(); e->tev_entry.clear();
    for (int k = 0; k < 5; k++) { out->ms[k] = e->k_ms[k]; out->bytes[k] = e->k_bytes[k]; out->launches[k] = e->k_launches[k]; }
    if (reset) { memset(e->k_ms, 0, sizeof e->k_ms); memset(e->k_bytes, 0, sizeof e->k_bytes); memset(e->k_launches, 0, sizeof e->k_launches); }
}

/* Compare every frame slot on the device with its pinned host mirror (which holds what the
 * normal decode delivered): returns the number of slots that differ.  Used after a replay. */
extern "C" u32 h264b200EngineCheckResident(h264b200_engine_t *e)
{
    if (!e) return 0xffffffffu;
    h264b200EngineSync(e);
    std::lock_guard<std::mutex> lk(e->mu);
    set_device(e);
    u32 bad = 0;
    std::vector<Inst *> all(e->insts);
    all.insert(all.end(), e->zombies.begin(), e->zombies.end());
    for (Inst *in : all) {
        if (in->out_format != H264B200_OUT_I420) continue;      /* the I420 mirror of an RGBA instance is not filled */
        std::vector<uint8_t> tmp(in->frame_bytes);
        for (uint32_t s = 0; s < in->n_slots; s++) {
            if (cudaMemcpy(tmp.data(), in->d_frames + (size_t)s * in->frame_stride, in->frame_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 0xffffffffu;
            if (memcmp(tmp.data(), in->h_frames + (size_t)s * in->frame_stride, in->frame_bytes)) bad++;
        }
    }
    return bad;
}

/* Debug / parity: records, coefficient slots and results of the device-parse picture most recently LAUNCHED for the
 * instance behind `inst`... (see h264b200DebugFetchParse in h264_decoder.c) */
extern "C" int h264b200_engine_fetch_parse(h264_backend_t *be, void *inst, int back, h264b200_mb_t *mbs, int16_t *coef, uint32_t coef_cap, KpResult *res)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (!in->dev_parse || back < 0 || back >= in->n_bufs) return -1;
    h264b200EngineSync(e);
    set_device(e);
    const PicBuf *p = &in->bufs[(in->next_buf + in->n_bufs - 1 - back) % in->n_bufs];
    if (p->state != 3) return -2;
    KpResult r;
    if (cudaMemcpy(&r, p->d_res, sizeof r, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    if (res) *res = r;
    if (mbs && cudaMemcpy(mbs, p->d_mbs, (size_t)in->n_mbs * sizeof(h264b200_mb_t), cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    if (coef) {
        const uint32_t n = r.coef_used < coef_cap ? r.coef_used : coef_cap;
        if (n && cudaMemcpy(coef, p->d_coef, (size_t)n * 32, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    }
    return 0;
}

extern "C" long h264b200DebugFetchPredeblock(storage_t *pStorage, uint8_t *out, size_t cap)
{
    h264_decoder_t *d = pStorage ? (h264_decoder_t *)pStorage->impl : NULL;
    if (!d || !d->be || !d->be_inst || d->be->inst_create != be_inst_create || !out) return -1;
    h264b200_engine *e = (h264b200_engine *)d->be->ctx; Inst *in = (Inst *)d->be_inst;
    if (!(e->flags & H264B200_ENGINE_TAP_PREDEBLOCK) || !in->d_pre || cap < in->frame_bytes) return -2;
    h264b200EngineSync(e);
    set_device(e);
    if (cudaMemcpy(out, in->d_pre, in->frame_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return -3;
    return (long)in->frame_bytes;
}

extern "C" int h264b200DebugFetchParse(storage_t *pStorage, int back, void *mbs, int16_t *coef, uint32_t coef_cap, uint32_t *res)
{
    h264_decoder_t *d = pStorage ? (h264_decoder_t *)pStorage->impl : NULL;
    if (!d || !d->be || !d->be_inst || !d->device_parse || d->be->inst_create != be_inst_create) return -1;
    return h264b200_engine_fetch_parse(d->be, d->be_inst, back, (h264b200_mb_t *)mbs, coef, coef_cap, (KpResult *)res);
}

/* ------------------------------------------------------- default backend */
static h264b200_engine *g_default_engine;
static std::mutex g_default_mu;

extern "C" h264_backend_t *h264_default_backend(void)
{
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_engine) {
        int dev = -1;
        const char *s = getenv("H264B200_DEVICE"), *pm = getenv("H264B200_PARSE");
        if (s && *s) dev = atoi(s);
        /* the synchronous single-instance API parses on the host by default (one picture at a time gives kernel Kp
         * nothing to run in parallel); H264B200_PARSE=device selects the device parser anyway */
        g_default_engine = h264b200EngineCreateEx(dev, (pm && !strcmp(pm, "device")) ? H264B200_ENGINE_DEVICE_PARSE : 0);
        if (!g_default_engine) {
            fprintf(stderr, "h264b200: no CUDA engine: this library has no CPU reconstruction path\n");
            return NULL;
        }
    }
    return &g_default_engine->be;
}
