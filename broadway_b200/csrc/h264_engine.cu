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
 * Streams: s_h2d (records/slots in) -> s_comp (K1..K4) -> s_d2h (frames out),
 * chained with events, so the copy-in of batch n+1 and the copy-out of batch
 * n-1 overlap the kernels of batch n.  There is no CPU reconstruction path in
 * this library: if CUDA is unusable h264_default_backend() returns NULL and
 * h264bsdDecode reports H264BSD_MEMALLOC_ERROR (reason on stderr).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
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

#define NBUF 3                 /* input buffers in flight per instance */
#define NSCR 4                 /* batch scratch sets in flight per engine */
#define CTRL_HEAD 16           /* int32 words before the progress counters: [0] K3 ticket, [1] K4 ticket */

#define CUDA_TRY(call, fail) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
    fprintf(stderr, "h264b200: %s -> %s (%s:%d)\n", #call, cudaGetErrorString(e__), __FILE__, __LINE__); fail; } } while (0)

struct Inst;

struct PicBuf {
    h264_pic_input_t in;           /* in.mbs / in.coef: ONE pinned block, records first, coefficient slots behind them */
    h264b200_mb_t *d_mbs;          /* device twin of the block: one cudaMemcpyAsync per picture */
    int16_t *d_coef;
    uint32_t d_coef_cap;           /* slots */
    cudaEvent_t done;              /* (not owned) event of the batch whose kernels read this buffer */
    int state;                     /* 0 free, 1 being filled by the parser, 2 queued, 3 launched */
    Inst *inst;
};

struct Inst {
    h264b200_engine *e;
    uint32_t wm, hm, n_mbs, n_slots;
    size_t frame_bytes;
    uint8_t *d_frames, *h_frames;
    cudaEvent_t slot_ready[H264_MAX_SLOTS];   /* (not owned) copy-out event of the batch that last wrote the slot's mirror */
    uint8_t slot_flags[H264_MAX_SLOTS];       /* bit 1: a copy-out into the slot's mirror has been issued (slot_ready is valid) */
    uint32_t slot_qgen[H264_MAX_SLOTS];       /* pictures handed over (queued) into the slot so far */
    uint32_t slot_lgen[H264_MAX_SLOTS];       /* generation of the last LAUNCHED picture of the slot */
    PicBuf bufs[NBUF];
    int next_buf;
    int batched;
    int queued;                               /* pictures of this instance waiting in the engine queue */
    /* optional output formatting (K5): cropped RGBA instead of the I420 frame */
    int out_format; int cl, ct, cw, ch;
    uint8_t *d_rgba, *h_rgba; size_t rgba_bytes;
};

struct Retained {                  /* one batch kept resident for replay */
    std::vector<PicJob> jobs;      /* host copy */
    PicJob *d_jobs;
    std::vector<void *> owned;     /* device allocations of this batch */
    Batch batch;
    size_t ctrl_words;
    bool k1, k2, k3, k4;
    uint64_t bytes[4];             /* algorithmic bytes per kernel family */
    uint32_t n_pics;
};

struct Scratch {
    PicJob *h_jobs, *d_jobs;
    uint32_t cap_jobs;
    int32_t *d_ctrl;
    size_t cap_ctrl;               /* words */
    cudaEvent_t done;              /* kernels of the batch finished */
    cudaEvent_t d2h_done;          /* copy-out of the batch finished */
    bool used;
};

struct h264b200_engine {
    int device, sm_count;
    cudaStream_t s_h2d, s_comp, s_d2h;
    cudaEvent_t ev_h2d, ev_comp, ev_rep0, ev_rep1;
    std::mutex mu;
    std::vector<PicBuf *> queue;
    std::vector<Inst *> insts;
    std::vector<Inst *> zombies;   /* shut-down instances whose frame pools retained batches still name */
    std::vector<Inst *> pool;      /* shut-down instances kept for reuse: pinned + device allocation is slow */
    Scratch scr[NSCR];
    int next_scr;
    uint32_t *d_err, *h_err;
    unsigned long long *d_trace; int trace_left;
    h264b200_stats_t st;
    uint32_t flags;                /* H264B200_ENGINE_* */
    std::vector<Retained *> retained;
    /* per-kernel timing of replays */
    std::vector<cudaEvent_t> tev;  /* 5 events per replayed batch */
    std::vector<int> tev_batch;
    double k_ms[4]; uint64_t k_bytes[4]; uint64_t k_launches[4];
    h264_backend_t be;
};

/* ----------------------------------------------------------------- helpers */
static void set_device(h264b200_engine *e) { cudaSetDevice(e->device); }

static int picbuf_alloc(PicBuf *p, Inst *in, uint32_t coef_cap)
{
    memset(p, 0, sizeof *p);
    p->inst = in;
    const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
    CUDA_TRY(cudaHostAlloc((void **)&p->in.mbs, rec_bytes + (size_t)coef_cap * 32, cudaHostAllocDefault), return -1);
    p->in.coef = (int16_t *)((uint8_t *)p->in.mbs + rec_bytes);
    CUDA_TRY(cudaMalloc((void **)&p->d_mbs, rec_bytes + (size_t)coef_cap * 32), return -1);
    p->d_coef = (int16_t *)((uint8_t *)p->d_mbs + rec_bytes);
    p->in.coef_cap = coef_cap; p->d_coef_cap = coef_cap;
    p->in.priv = p;
    return 0;
}
static void picbuf_free(PicBuf *p)
{
    if (p->in.mbs) cudaFreeHost(p->in.mbs);
    if (p->d_mbs) cudaFree(p->d_mbs);
    memset(p, 0, sizeof *p);
}

/* algorithmic bytes of one picture per kernel family, as SURVEY.md 8(d) defines them */
static void count_bytes(const h264_pic_input_t *pic, uint32_t n_mbs, uint64_t out[4])
{
    uint64_t inter = 0, intra = 0, blocks = 0, dc = 0, blk_inter = 0, blk_intra = 0, dbk = 0;
    for (uint32_t i = 0; i < n_mbs; i++) {
        const h264b200_mb_t *m = &pic->mbs[i];
        uint32_t nb = (uint32_t)__builtin_popcount(m->resid_mask & 0xffffffu), nd = (uint32_t)__builtin_popcount(m->resid_mask >> 24);
        if (m->mb_class == H264B200_MB_MISSING) continue;
        if (m->mb_class == H264B200_MB_INTER) { inter++; blk_inter += nb; } else { intra++; blk_intra += nb; if (m->mb_class == H264B200_MB_IPCM) blk_intra += 12; }
        blocks += nb; dc += nd;
        if (m->dbk_flags) dbk++;
    }
    out[0] = blocks * 64 + dc * 64;                         /* K1: 32 B in + 32 B out per coded block / DC block */
    out[1] = inter * (384 + 384 + 128) + blk_inter * 32;    /* K2 */
    out[2] = intra * (384 + 64 + 128) + blk_intra * 32;     /* K3 */
    out[3] = dbk * (384 + 384 + 128);                       /* K4 */
}

/* ------------------------------------------------------------ kernel launch */
struct BatchPlan { bool k1, k2, k3, k3c, k4; uint32_t total_mbs; int max_hm; int n_jobs; };

static void launch_kernels(h264b200_engine *e, const Batch &b, const BatchPlan &pl, cudaEvent_t *tev)
{
    cudaStream_t s = e->s_comp;
    if (tev) cudaEventRecord(tev[0], s);
    if (pl.k1) { uint32_t blocks = (pl.total_mbs * 8 + 255) / 256; k1_transform<<<blocks, 256, 0, s>>>(b); e->st.kernel_launches++; }   /* 8 lanes per macroblock */
    if (tev) cudaEventRecord(tev[1], s);
    if (pl.k2) { k2_inter<<<(pl.total_mbs * 16 + K2_THREADS - 1) / K2_THREADS, K2_THREADS, 0, s>>>(b); e->st.kernel_launches++; }   /* 16 threads per macroblock */
    if (tev) cudaEventRecord(tev[2], s);
    uint32_t n_tasks = (uint32_t)pl.n_jobs * (uint32_t)pl.max_hm;
    if (pl.k3) {
        uint32_t blocks = (n_tasks + K3_WARPS - 1) / K3_WARPS, cap = (uint32_t)e->sm_count * 16;
        k3_intra<<<blocks < cap ? blocks : cap, K3_WARPS * 32, 0, s>>>(b); e->st.kernel_launches++;
    }
    if (pl.k3c) { k3c_conceal<<<pl.n_jobs, 32, 0, s>>>(b); e->st.kernel_launches++; }   /* lost slices only: one warp per picture */
    if (tev) cudaEventRecord(tev[3], s);
    if (pl.k4) {                   /* one warp per PAIR of macroblock rows */
        uint32_t n_pairs = (uint32_t)pl.n_jobs * (((uint32_t)pl.max_hm + 1) / 2);
        uint32_t blocks = (n_pairs + K4_WARPS - 1) / K4_WARPS, cap = (uint32_t)e->sm_count * 16;
        k4_deblock<<<blocks < cap ? blocks : cap, K4_WARPS * 32, 0, s>>>(b);
        e->st.kernel_launches++;
    }
    if (tev) cudaEventRecord(tev[4], s);
}

/* ------------------------------------------------------------------ submit */
/* engine mutex held.  Launch every queued picture as one batch. */
static uint32_t submit_locked(h264b200_engine *e)
{
    uint32_t n = (uint32_t)e->queue.size();
    if (!n) return 0;
    set_device(e);
    const bool retain = (e->flags & H264B200_ENGINE_RETAIN) != 0;
    Retained *ret = retain ? new Retained() : nullptr;
    Scratch &sc = e->scr[e->next_scr];
    e->next_scr = (e->next_scr + 1) % NSCR;
    if (sc.used) cudaEventSynchronize(sc.done);
    size_t ctrl_words = CTRL_HEAD;
    for (PicBuf *p : e->queue) ctrl_words += 2 * (size_t)p->inst->hm;
    if (sc.cap_jobs < n) {
        if (sc.h_jobs) cudaFreeHost(sc.h_jobs);
        if (sc.d_jobs) cudaFree(sc.d_jobs);
        sc.cap_jobs = n + 16;
        CUDA_TRY(cudaHostAlloc((void **)&sc.h_jobs, sc.cap_jobs * sizeof(PicJob), cudaHostAllocDefault), return 0);
        CUDA_TRY(cudaMalloc((void **)&sc.d_jobs, sc.cap_jobs * sizeof(PicJob)), return 0);
    }
    if (sc.cap_ctrl < ctrl_words) {
        if (sc.d_ctrl) cudaFree(sc.d_ctrl);
        sc.cap_ctrl = ctrl_words + 1024;
        CUDA_TRY(cudaMalloc((void **)&sc.d_ctrl, sc.cap_ctrl * sizeof(int32_t)), return 0);
    }
    PicJob *d_jobs = sc.d_jobs; int32_t *d_ctrl = sc.d_ctrl;
    if (retain) {                  /* a retained batch owns its job table and control area */
        CUDA_TRY(cudaMalloc((void **)&ret->d_jobs, n * sizeof(PicJob)), return 0);
        ret->owned.push_back(ret->d_jobs);
        d_jobs = ret->d_jobs;
        CUDA_TRY(cudaMalloc((void **)&d_ctrl, ctrl_words * sizeof(int32_t)), return 0);
        ret->owned.push_back(d_ctrl);
    }

    BatchPlan pl; memset(&pl, 0, sizeof pl);
    uint32_t mb_base = 0; size_t prog_off = CTRL_HEAD;
    uint64_t bytes[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = e->queue[i]; Inst *in = p->inst; h264_pic_input_t *pic = &p->in;
        h264b200_mb_t *d_mbs = p->d_mbs; int16_t *d_coef_in = p->d_coef, *d_coef = p->d_coef;
        size_t coef_bytes = (size_t)pic->coef_used * 32;
        const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
        if (retain) {                                          /* [records | levels] and a separate residual buffer, kept */
            void *a = nullptr, *c = nullptr;
            CUDA_TRY(cudaMalloc(&a, rec_bytes + coef_bytes + 32), return 0);
            CUDA_TRY(cudaMalloc(&c, coef_bytes + 32), return 0);
            ret->owned.push_back(a); ret->owned.push_back(c);
            d_mbs = (h264b200_mb_t *)a; d_coef_in = (int16_t *)((uint8_t *)a + rec_bytes); d_coef = (int16_t *)c;
        } else if (p->d_coef_cap < pic->coef_used) {          /* the host side grew: follow */
            /* no earlier batch reads this ring slot any more (pic_begin waited for `done`), so it can be replaced */
            cudaFree(p->d_mbs);
            p->d_coef_cap = pic->coef_cap;
            CUDA_TRY(cudaMalloc((void **)&p->d_mbs, rec_bytes + (size_t)p->d_coef_cap * 32), return 0);
            p->d_coef = (int16_t *)((uint8_t *)p->d_mbs + rec_bytes);
            d_mbs = p->d_mbs; d_coef_in = d_coef = p->d_coef;
        }
        /* records and coefficient slots are adjacent on both sides: one copy */
        CUDA_TRY(cudaMemcpyAsync(d_mbs, pic->mbs, rec_bytes + coef_bytes, cudaMemcpyHostToDevice, e->s_h2d), return 0);
        e->st.h2d_bytes += rec_bytes + coef_bytes;

        PicJob &j = sc.h_jobs[i];
        j.mbs = d_mbs; j.coef_in = d_coef_in; j.coef = d_coef;
        j.cur = in->d_frames + (size_t)pic->cur_slot * in->frame_bytes;
        j.frames = in->d_frames; j.frame_bytes = (uint32_t)in->frame_bytes;
        j.wm = (int32_t)in->wm; j.hm = (int32_t)in->hm;
        j.progress = d_ctrl + prog_off; prog_off += 2 * (size_t)in->hm;
        j.n_intra = pic->n_intra; j.n_inter = pic->n_inter; j.any_deblock = pic->any_deblock;
        j.mb_base = mb_base; mb_base += in->n_mbs;
        j.n_conceal = pic->n_conceal;
        j.conceal_list = reinterpret_cast<const uint32_t *>(d_coef_in + (size_t)pic->conceal_offset * 16);
        if (pic->n_conceal) pl.k3c = true;
        if (pic->coef_used) pl.k1 = true;
        if (pic->n_inter) pl.k2 = true;
        if (pic->n_intra) pl.k3 = true;
        if (pic->any_deblock) pl.k4 = true;
        if ((int)in->hm > pl.max_hm) pl.max_hm = (int)in->hm;
        /* the frame being written may still be on its way to the host from an earlier batch */
        if (in->slot_flags[pic->cur_slot] & 2) cudaStreamWaitEvent(e->s_comp, in->slot_ready[pic->cur_slot], 0);
        if (retain) { uint64_t bb[4]; count_bytes(pic, in->n_mbs, bb); for (int k = 0; k < 4; k++) bytes[k] += bb[k]; }
    }
    pl.total_mbs = mb_base; pl.n_jobs = (int)n;

    Batch b;
    b.jobs = d_jobs; b.n_jobs = (int32_t)n; b.max_hm = pl.max_hm; b.total_mbs = mb_base;
    b.uniform_mbs = e->queue[0]->inst->n_mbs;
    for (PicBuf *p : e->queue) if (p->inst->n_mbs != b.uniform_mbs) b.uniform_mbs = 0;
    b.tickets = (uint32_t *)d_ctrl; b.error_flags = e->d_err; b.trace = e->trace_left > 0 ? e->d_trace : nullptr;

    cudaEventRecord(e->ev_h2d, e->s_h2d);
    cudaStreamWaitEvent(e->s_comp, e->ev_h2d, 0);
    cudaMemcpyAsync(d_jobs, sc.h_jobs, n * sizeof(PicJob), cudaMemcpyHostToDevice, e->s_comp);
    cudaMemsetAsync(d_ctrl, 0, ctrl_words * sizeof(int32_t), e->s_comp);
    launch_kernels(e, b, pl, nullptr);
    if (b.trace) {                 /* debug: dump the wavefront timing of job 0 of this batch */
        std::vector<unsigned long long> h(256 + 4 * 512);
        cudaStreamSynchronize(e->s_comp);
        cudaMemcpy(h.data(), e->d_trace, h.size() * 8, cudaMemcpyDeviceToHost);
        fprintf(stderr, "h264b200 trace: batch of %u, job0 %dx%d MBs; per row: start, first-mb, mid, end (us from row 0 start)\n", n, sc.h_jobs[0].wm, sc.h_jobs[0].hm);
        for (int r = 0; r < sc.h_jobs[0].hm; r++) { const unsigned long long *q = &h[256 + r * 4], t0 = h[256];
            fprintf(stderr, "  row %2d: %8.1f %8.1f %8.1f %8.1f\n", r, (q[0] - t0) / 1e3, (q[1] - t0) / 1e3, (q[2] - t0) / 1e3, (q[3] - t0) / 1e3); }
        cudaMemset(e->d_trace, 0, h.size() * 8);
        e->trace_left--;
    }
    cudaEventRecord(sc.done, e->s_comp); sc.used = true;
    cudaEventRecord(e->ev_comp, e->s_comp);
    cudaStreamWaitEvent(e->s_d2h, e->ev_comp, 0);
    for (uint32_t i = 0; i < n; i++) {
        PicBuf *p = e->queue[i]; Inst *in = p->inst; int slot = p->in.cur_slot;
        p->done = sc.done;
        p->state = 3;
        if (in->out_format == H264B200_OUT_RGBA && in->d_rgba) {
            /* K5 runs on the copy-out stream: it only reads the finished frame */
            RgbaJob rj; rj.frame = in->d_frames + (size_t)slot * in->frame_bytes; rj.out = in->d_rgba + (size_t)slot * in->rgba_bytes;
            rj.W = (int)in->wm * 16; rj.H = (int)in->hm * 16; rj.cl = in->cl; rj.ct = in->ct; rj.cw = in->cw; rj.ch = in->ch;
            const int items = ((rj.cw + 3) / 4) * ((rj.ch + 1) / 2);
            k5_rgba<<<(items + 255) / 256, 256, 0, e->s_d2h>>>(rj); e->st.kernel_launches++;
            if (!(e->flags & H264B200_ENGINE_NO_D2H)) {
                cudaMemcpyAsync(in->h_rgba + (size_t)slot * in->rgba_bytes, rj.out, in->rgba_bytes, cudaMemcpyDeviceToHost, e->s_d2h);
                e->st.d2h_bytes += in->rgba_bytes;
            }
        } else if (!(e->flags & H264B200_ENGINE_NO_D2H)) {
            cudaMemcpyAsync(in->h_frames + (size_t)slot * in->frame_bytes, in->d_frames + (size_t)slot * in->frame_bytes,
                            in->frame_bytes, cudaMemcpyDeviceToHost, e->s_d2h);
            e->st.d2h_bytes += in->frame_bytes;
        }
        in->slot_ready[slot] = sc.d2h_done;
        in->slot_flags[slot] = 2; in->slot_lgen[slot] = in->slot_qgen[slot];
        in->queued--;
    }
    cudaMemcpyAsync(e->h_err, e->d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, e->s_d2h);
    cudaEventRecord(sc.d2h_done, e->s_d2h);
    e->st.pictures += n; e->st.batches++;
    if (retain) {
        ret->jobs.assign(sc.h_jobs, sc.h_jobs + n);
        ret->batch = b; ret->batch.trace = nullptr; ret->ctrl_words = ctrl_words;
        ret->k1 = pl.k1; ret->k2 = pl.k2; ret->k3 = pl.k3; ret->k4 = pl.k4;
        for (int k = 0; k < 4; k++) ret->bytes[k] = bytes[k];
        ret->n_pics = n;
        e->retained.push_back(ret);
    }
    e->queue.clear();
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) fprintf(stderr, "h264b200: kernel launch failed: %s\n", cudaGetErrorString(le));
    return n;
}

/* ------------------------------------------------------- backend callbacks */
static void inst_free(Inst *in);
static void *be_inst_create(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n_slots)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx;
    if (n_slots > H264_MAX_SLOTS) return NULL;
    set_device(e);
    {   /* reuse a pooled instance of the same geometry */
        std::lock_guard<std::mutex> lk(e->mu);
        for (size_t i = 0; i < e->pool.size(); i++) {
            Inst *c = e->pool[i];
            if (c->wm == wm && c->hm == hm && c->n_slots == n_slots) {
                e->pool.erase(e->pool.begin() + i);
                memset(c->slot_flags, 0, sizeof c->slot_flags); memset(c->slot_qgen, 0, sizeof c->slot_qgen); memset(c->slot_lgen, 0, sizeof c->slot_lgen);
                c->next_buf = 0; c->queued = 0; c->out_format = H264B200_OUT_I420;
                c->batched = (e->flags & H264B200_ENGINE_BATCHED) != 0;
                for (int k = 0; k < NBUF; k++) c->bufs[k].state = 0;
                e->insts.push_back(c);
                return c;
            }
        }
    }
    Inst *in = (Inst *)calloc(1, sizeof *in);
    if (!in) return NULL;
    in->e = e; in->wm = wm; in->hm = hm; in->n_mbs = wm * hm; in->n_slots = n_slots;
    in->frame_bytes = (size_t)in->n_mbs * 384;
    in->batched = (e->flags & H264B200_ENGINE_BATCHED) != 0;
    CUDA_TRY(cudaMalloc((void **)&in->d_frames, in->frame_bytes * n_slots), { free(in); return NULL; });
    CUDA_TRY(cudaMemset(in->d_frames, 0, in->frame_bytes * n_slots), { cudaFree(in->d_frames); free(in); return NULL; });
    CUDA_TRY(cudaHostAlloc((void **)&in->h_frames, in->frame_bytes * n_slots, cudaHostAllocDefault), { cudaFree(in->d_frames); free(in); return NULL; });
    memset(in->h_frames, 0, in->frame_bytes * n_slots);
    for (int i = 0; i < NBUF; i++) if (picbuf_alloc(&in->bufs[i], in, in->n_mbs * 10 + 64)) { inst_free(in); return NULL; }   /* frees what was allocated so far */
    std::lock_guard<std::mutex> lk(e->mu);
    e->insts.push_back(in);
    return in;
}

static void inst_free(Inst *in)
{
    for (int i = 0; i < NBUF; i++) picbuf_free(&in->bufs[i]);
    cudaFree(in->d_frames); cudaFreeHost(in->h_frames);
    if (in->d_rgba) { cudaFree(in->d_rgba); cudaFreeHost(in->h_rgba); }
    free(in);
}

static void be_inst_destroy(h264_backend_t *be, void *inst)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    bool keep;
    {
        std::lock_guard<std::mutex> lk(e->mu);
        if (in->queued) submit_locked(e);
        keep = !e->retained.empty();           /* retained batches name this instance's frame pool */
        if (keep) e->zombies.push_back(in);
        else for (size_t i = 0; i < e->insts.size(); i++) if (e->insts[i] == in) { e->insts.erase(e->insts.begin() + i); break; }
    }
    set_device(e);
    cudaStreamSynchronize(e->s_comp); cudaStreamSynchronize(e->s_d2h);
    if (!keep) { std::lock_guard<std::mutex> lk(e->mu); e->pool.push_back(in); }
}

static h264_pic_input_t *be_pic_begin(h264_backend_t *be, void *inst)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    PicBuf *p = &in->bufs[in->next_buf];
    in->next_buf = (in->next_buf + 1) % NBUF;
    if (p->state == 2) { std::lock_guard<std::mutex> lk(e->mu); submit_locked(e); }
    if (p->state == 3) { set_device(e); cudaEventSynchronize(p->done); }
    p->state = 1;
    return &p->in;
}

static int be_coef_grow(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_slots)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    const size_t rec_bytes = (size_t)in->n_mbs * sizeof(h264b200_mb_t);
    uint32_t cap = pic->coef_cap * 2 > min_slots ? pic->coef_cap * 2 : min_slots;
    uint8_t *n = NULL;
    set_device(e);
    CUDA_TRY(cudaHostAlloc((void **)&n, rec_bytes + (size_t)cap * 32, cudaHostAllocDefault), return -1);
    memcpy(n, pic->mbs, rec_bytes + (size_t)pic->coef_used * 32);           /* records written so far and their slots */
    cudaFreeHost(pic->mbs);
    pic->mbs = (h264b200_mb_t *)n; pic->coef = (int16_t *)(n + rec_bytes); pic->coef_cap = cap;
    return 0;
}

static int be_pic_submit(h264_backend_t *be, void *inst, h264_pic_input_t *pic)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    PicBuf *p = (PicBuf *)pic->priv;
    std::lock_guard<std::mutex> lk(e->mu);
    if (in->queued) submit_locked(e);          /* consecutive pictures of one stream depend on each other */
    p->state = 2;
    in->slot_qgen[pic->cur_slot]++;
    in->queued++;
    e->queue.push_back(p);
    if (!in->batched) submit_locked(e);
    return 0;
}

/* Wait until generation `gen` of `slot` is in its host mirror.  Returns 0, or 1 when a LATER picture has already been
 * launched into the slot (its mirror may be overwritten: the caller waited too long), or -1 on a CUDA error. */
static int wait_slot(h264b200_engine *e, Inst *in, int slot, uint32_t gen)
{
    if ((int32_t)(in->slot_lgen[slot] - gen) < 0) { std::lock_guard<std::mutex> lk(e->mu); submit_locked(e); }
    if (in->slot_lgen[slot] != gen) return 1;
    if (in->slot_flags[slot] & 2) {
        set_device(e);
        cudaError_t er = cudaEventSynchronize(in->slot_ready[slot]);
        if (er != cudaSuccess) { fprintf(stderr, "h264b200: reconstruction failed: %s\n", cudaGetErrorString(er)); return -1; }
    }
    return 0;
}

static uint8_t *slot_mirror(Inst *in, int slot)
{
    if (in->out_format == H264B200_OUT_RGBA && in->h_rgba) return in->h_rgba + (size_t)slot * in->rgba_bytes;
    return in->h_frames + (size_t)slot * in->frame_bytes;
}

static uint8_t *be_frame_host(h264_backend_t *be, void *inst, int slot, uint32_t *error_flags)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (slot < 0 || slot >= (int)in->n_slots) return NULL;
    if (wait_slot(e, in, slot, in->slot_qgen[slot]) < 0) return NULL;
    if (error_flags) *error_flags = *e->h_err;
    return slot_mirror(in, slot);
}

/* non-blocking: where the newest picture of `slot` will be, and its generation (for frame_wait) */
static uint8_t *be_frame_host_async(h264_backend_t *be, void *inst, int slot, uint32_t *gen)
{
    Inst *in = (Inst *)inst; (void)be;
    if (slot < 0 || slot >= (int)in->n_slots) return NULL;
    if (gen) *gen = in->slot_qgen[slot];
    return slot_mirror(in, slot);
}

static int be_frame_wait(h264_backend_t *be, void *inst, int slot, uint32_t gen, uint32_t *error_flags)
{
    h264b200_engine *e = (h264b200_engine *)be->ctx; Inst *in = (Inst *)inst;
    if (slot < 0 || slot >= (int)in->n_slots) return -1;
    /* the ticket carries 24 bits of the generation: take the value nearest below the current one */
    uint32_t full = (in->slot_qgen[slot] & 0xff000000u) | (gen & 0xffffffu);
    if (full > in->slot_qgen[slot]) full -= 1u << 24;
    int rc = wait_slot(e, in, slot, full);
    if (error_flags) *error_flags = *e->h_err;
    return rc;
}

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
    e->device = device; e->flags = flags; e->next_scr = 0;
    memset(&e->st, 0, sizeof e->st); memset(e->scr, 0, sizeof e->scr);
    memset(e->k_ms, 0, sizeof e->k_ms); memset(e->k_bytes, 0, sizeof e->k_bytes); memset(e->k_launches, 0, sizeof e->k_launches);
    CUDA_TRY(cudaSetDevice(device), { delete e; return NULL; });
    cudaDeviceProp p;
    CUDA_TRY(cudaGetDeviceProperties(&p, device), { delete e; return NULL; });
    e->sm_count = p.multiProcessorCount;
    CUDA_TRY(cudaStreamCreateWithFlags(&e->s_h2d, cudaStreamNonBlocking), { delete e; return NULL; });
    CUDA_TRY(cudaStreamCreateWithFlags(&e->s_comp, cudaStreamNonBlocking), { delete e; return NULL; });
    CUDA_TRY(cudaStreamCreateWithFlags(&e->s_d2h, cudaStreamNonBlocking), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_h2d, cudaEventDisableTiming), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreateWithFlags(&e->ev_comp, cudaEventDisableTiming), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreate(&e->ev_rep0), { delete e; return NULL; });
    CUDA_TRY(cudaEventCreate(&e->ev_rep1), { delete e; return NULL; });
    for (int i = 0; i < NSCR; i++) {
        CUDA_TRY(cudaEventCreateWithFlags(&e->scr[i].done, cudaEventDisableTiming), { delete e; return NULL; });
        CUDA_TRY(cudaEventCreateWithFlags(&e->scr[i].d2h_done, cudaEventDisableTiming), { delete e; return NULL; });
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
    e->be.inst_create = be_inst_create; e->be.inst_destroy = be_inst_destroy; e->be.pic_begin = be_pic_begin;
    e->be.coef_grow = be_coef_grow; e->be.pic_submit = be_pic_submit; e->be.frame_host = be_frame_host;
    e->be.frame_host_async = be_frame_host_async;
    e->be.frame_wait = be_frame_wait;
    e->be.set_output = be_set_output;
    e->be.destroy = be_destroy; e->be.ctx = e;
    return e;
}
extern "C" h264b200_engine_t *h264b200EngineCreate(int device) { return h264b200EngineCreateEx(device, H264B200_ENGINE_BATCHED); }

static void free_retained(h264b200_engine *e)
{
    for (Retained *r : e->retained) { for (void *p : r->owned) cudaFree(p); delete r; }
    e->retained.clear();
    for (Inst *z : e->zombies) {
        for (size_t i = 0; i < e->insts.size(); i++) if (e->insts[i] == z) { e->insts.erase(e->insts.begin() + i); break; }
        inst_free(z);
    }
    e->zombies.clear();
}

extern "C" void h264b200EngineDestroy(h264b200_engine_t *e)
{
    if (!e) return;
    set_device(e);
    cudaStreamSynchronize(e->s_h2d); cudaStreamSynchronize(e->s_comp); cudaStreamSynchronize(e->s_d2h);
    free_retained(e);
    for (Inst *p : e->pool) inst_free(p);
    e->pool.clear();
    for (cudaEvent_t ev : e->tev) cudaEventDestroy(ev);
    for (int i = 0; i < NSCR; i++) {
        Scratch &s = e->scr[i];
        if (s.h_jobs) cudaFreeHost(s.h_jobs);
        if (s.d_jobs) cudaFree(s.d_jobs);
        if (s.d_ctrl) cudaFree(s.d_ctrl);
        cudaEventDestroy(s.done); cudaEventDestroy(s.d2h_done);
    }
    cudaFree(e->d_err); cudaFreeHost(e->h_err);
    cudaEventDestroy(e->ev_h2d); cudaEventDestroy(e->ev_comp); cudaEventDestroy(e->ev_rep0); cudaEventDestroy(e->ev_rep1);
    cudaStreamDestroy(e->s_h2d); cudaStreamDestroy(e->s_comp); cudaStreamDestroy(e->s_d2h);
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
    return submit_locked(e);
}

extern "C" void h264b200EngineSync(h264b200_engine_t *e)
{
    if (!e) return;
    set_device(e);
    cudaError_t a = cudaStreamSynchronize(e->s_h2d), b = cudaStreamSynchronize(e->s_comp), c = cudaStreamSynchronize(e->s_d2h);
    if (a != cudaSuccess || b != cudaSuccess || c != cudaSuccess)
        fprintf(stderr, "h264b200: engine sync failed: %s\n", cudaGetErrorString(a != cudaSuccess ? a : b != cudaSuccess ? b : c));
}

extern "C" void h264b200EngineStats(h264b200_engine_t *e, h264b200_stats_t *out) { if (e && out) { std::lock_guard<std::mutex> lk(e->mu); *out = e->st; } }
extern "C" u32 h264b200EngineErrorFlags(h264b200_engine_t *e) { return e ? *e->h_err : 0; }
extern "C" void h264b200EngineSetFlags(h264b200_engine_t *e, uint32_t flags) { if (e) { std::lock_guard<std::mutex> lk(e->mu); e->flags = flags; } }
extern "C" uint32_t h264b200EngineFlags(h264b200_engine_t *e) { if (!e) return 0; std::lock_guard<std::mutex> lk(e->mu); return e->flags; }

/* -------------------------------------------------------- resident replay */
extern "C" void h264b200EngineDropRetained(h264b200_engine_t *e)
{
    if (!e) return;
    h264b200EngineSync(e);
    std::lock_guard<std::mutex> lk(e->mu);
    free_retained(e);
}

extern "C" u32 h264b200EngineReplay(h264b200_engine_t *e, u32 reps, int time_kernels)
{
    if (!e) return 0;
    std::lock_guard<std::mutex> lk(e->mu);
    set_device(e);
    size_t need = 0;
    for (Retained *r : e->retained) if (r->ctrl_words > need) need = r->ctrl_words;
    u32 pics = 0;
    cudaEventRecord(e->ev_rep0, e->s_comp);
    for (u32 rep = 0; rep < reps; rep++) for (size_t bi = 0; bi < e->retained.size(); bi++) {
        Retained *r = e->retained[bi];
        BatchPlan pl; pl.k1 = r->k1; pl.k2 = r->k2; pl.k3 = r->k3; pl.k4 = r->k4;
        pl.k3c = false; for (const PicJob &pj : r->jobs) if (pj.n_conceal) pl.k3c = true;
        pl.total_mbs = r->batch.total_mbs; pl.max_hm = r->batch.max_hm; pl.n_jobs = r->batch.n_jobs;
        cudaEvent_t *tev = nullptr;
        if (time_kernels) {
            size_t base = e->tev.size();
            for (int k = 0; k < 5; k++) { cudaEvent_t ev; cudaEventCreate(&ev); e->tev.push_back(ev); }
            e->tev_batch.push_back((int)bi);
            tev = &e->tev[base];
        }
        /* the batch's own control area (tickets + wavefront progress) is part of its retained allocation
         * (r->batch.tickets points at it); batches are serialised on s_comp */
        cudaMemsetAsync(r->batch.tickets, 0, r->ctrl_words * sizeof(int32_t), e->s_comp);
        launch_kernels(e, r->batch, pl, tev);
        pics += r->n_pics;
    }
    cudaEventRecord(e->ev_rep1, e->s_comp);
    e->st.pictures += pics; e->st.batches += (uint64_t)reps * e->retained.size();
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

/* Fold the event pairs of timed replays into the per-kernel totals (after a sync). */
extern "C" void h264b200EngineKernelTimes(h264b200_engine_t *e, h264b200_kernel_times_t *out, int reset)
{
    if (!e || !out) return;
    h264b200EngineSync(e);
    std::lock_guard<std::mutex> lk(e->mu);
    for (size_t i = 0; i < e->tev_batch.size(); i++) {
        Retained *r = e->retained[(size_t)e->tev_batch[i]];
        const bool on[4] = {r->k1, r->k2, r->k3, r->k4};
        for (int k = 0; k < 4; k++) if (on[k]) {
            float ms = 0; cudaEventElapsedTime(&ms, e->tev[5 * i + k], e->tev[5 * i + k + 1]);
            e->k_ms[k] += ms; e->k_bytes[k] += r->bytes[k]; e->k_launches[k]++;
        }
    }
    for (cudaEvent_t ev : e->tev) cudaEventDestroy(ev);
    e->tev.clear(); e->tev_batch.clear();
    for (int k = 0; k < 4; k++) { out->ms[k] = e->k_ms[k]; out->bytes[k] = e->k_bytes[k]; out->launches[k] = e->k_launches[k]; }
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
    for (Inst *in : e->insts) {
        if (in->out_format != H264B200_OUT_I420) continue;      /* the I420 mirror of an RGBA instance is not filled */
        std::vector<uint8_t> tmp(in->frame_bytes);
        for (uint32_t s = 0; s < in->n_slots; s++) {
            if (cudaMemcpy(tmp.data(), in->d_frames + (size_t)s * in->frame_bytes, in->frame_bytes, cudaMemcpyDeviceToHost) != cudaSuccess) return 0xffffffffu;
            if (memcmp(tmp.data(), in->h_frames + (size_t)s * in->frame_bytes, in->frame_bytes)) bad++;
        }
    }
    return bad;
}

/* ------------------------------------------------------- default backend */
static h264b200_engine *g_default_engine;
static std::mutex g_default_mu;

extern "C" h264_backend_t *h264_default_backend(void)
{
    std::lock_guard<std::mutex> lk(g_default_mu);
    if (!g_default_engine) {
        int dev = -1;
        const char *s = getenv("H264B200_DEVICE");
        if (s && *s) dev = atoi(s);
        g_default_engine = h264b200EngineCreateEx(dev, 0);
        if (!g_default_engine) {
            fprintf(stderr, "h264b200: no CUDA engine: this library has no CPU reconstruction path\n");
            return NULL;
        }
    }
    return &g_default_engine->be;
}
