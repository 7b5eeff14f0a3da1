/* h264_runner.c — many independent streams (or IDR-bounded GOP segments) through
 * one engine: h264b200DecodeStreams and h264b200SplitGops
 * (include/h264b200_batch.h).
 *
 * The reference's closest relative is TestBenchMultipleInstance.c:60-350, which
 * steps N decoder instances round-robin in one thread.  Here the same
 * round-robin is the unit of GPU batching, without ever idling a parser thread:
 * work items (round r, stream s) are claimed from one atomic counter, a
 * picture's serial CAVLC parse runs on whichever thread claimed it, and the
 * streams are split into two groups that are launched separately — the last
 * thread to finish a group's pictures of a round launches them as one batch
 * while everybody else is already parsing the other group.  The only thing a
 * claim ever waits for is that the same stream's previous picture has been
 * launched (consecutive pictures of a stream depend on each other), which in
 * steady state happened half a round earlier.  Output of round r is collected
 * right after the stream's round r+1 picture has been parsed
 * (h264b200NextOutputPictureAsync / h264b200PictureWait): the GPU had a whole
 * round to finish it.
 * Streams are independent (no mutable globals in the decoder core), and an IDR
 * picture empties the DPB (h264bsd_dpb.c:675-708), which is what makes
 * h264b200SplitGops' segments decodable on their own — on another instance,
 * thread or GPU.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include "h264b200.h"
#include "h264b200_batch.h"

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }

#define MAX_PENDING 20
#define STALL_SECONDS 300.0   /* device-parse pipeline: no launch, no scan, no collected picture for this long = stuck (a single Kp launch over 4K
                                 pictures takes ~1 s, under a profiler's kernel replay a hundred times that) */
#ifndef H264B200_HOST_SHARE_DEFAULT
#define H264B200_HOST_SHARE_DEFAULT "0"      /* "auto" once it has been measured to pay on the box at hand */
#endif

typedef struct { u8 *ptr; u32 ticket, pic_id, err; } pending_t;

typedef struct {
    storage_t st;
    uint8_t *buf; const uint8_t *src; size_t len, pos;
    u32 pic_id, out_index;
    int inited, finished, failed, flushed;
    pending_t cur[MAX_PENDING], prev[MAX_PENDING];
    int n_cur, n_prev;
    u32 width, height;
    /* device-parse pipeline: outputs popped while scanning ahead, oldest first */
    pending_t *outq; u32 outq_cap, outq_head, outq_n;
    u32 depth;                  /* look-ahead this stream may use */
    u32 chunk;                  /* pictures scanned per round at most (= pictures per stream in a Kp launch) */
    int done;                   /* device-parse: finished, everything launched and handed out */
    int host_parse;             /* device-parse engine, but this stream's slice data is parsed by the worker threads (the host share) */
} rstream_t;

typedef struct runner runner_t;
typedef struct {
    runner_t *r; uint32_t tid;
    pthread_t th;
    uint64_t pictures, bytes_out; uint32_t err_mbs;
    double parse_s, wait_s;
    uint32_t *own; uint32_t n_own;            /* device-parse: the streams this worker owns (a decoder instance is single-threaded) */
} worker_t;

#define MAX_GROUPS 2
struct runner {
    h264b200_engine_t *e;
    rstream_t *s; uint32_t n_streams, n_threads;
    h264b200_picture_cb cb; void *user;
    pthread_mutex_t mu;
    uint64_t next_item;                       /* next (round, stream) work item: round = item / n_streams */
    uint32_t n_groups, gstart[MAX_GROUPS + 1];/* group g = streams [gstart[g], gstart[g+1]) */
    uint32_t arrived[MAX_GROUPS], produced[MAX_GROUPS], dead[MAX_GROUPS];   /* under mu: the group's current round */
    uint32_t launched[MAX_GROUPS];            /* rounds of the group that have been handed to the GPU (release / acquire) */
    int stop;                                 /* every group went through a round without producing a picture */
    uint32_t n_workers;                       /* device-parse: worker threads (the streams are dealt out to them) ... */
    uint32_t workers_alive;                   /* ... still sweeping */
    int inline_drive;                         /* device-parse with one thread: the worker takes the scheduling steps itself */
    uint32_t activity;                        /* device-parse: pictures scanned + outputs collected so far, all workers (progress indicator for the scheduling thread) */
    uint32_t rounds;                          /* batches launched */
};

static void pop_outputs(rstream_t *s)
{
    u32 id, idr, err, ticket; u8 *p;
    while (s->n_cur < MAX_PENDING && (p = h264b200NextOutputPictureAsync(&s->st, &id, &idr, &err, &ticket)) != NULL) {
        pending_t *q = &s->cur[s->n_cur++];
        q->ptr = p; q->ticket = ticket; q->pic_id = id; q->err = err;
    }
}

/* advance one stream by one picture; 1 if a picture was produced */
static int step_stream(rstream_t *s)
{
    while (s->pos < s->len) {
        u32 nread = 0, rest = (u32)(s->len - s->pos > 0x7fffffffu ? 0x7fffffffu : s->len - s->pos);
        u32 rc = h264bsdDecode(&s->st, s->buf + s->pos, rest, s->pic_id, &nread);
        s->pos += nread;
        if (rc == H264BSD_PIC_RDY) { s->pic_id++; pop_outputs(s); return 1; }
        if (rc == H264BSD_HDRS_RDY) { pop_outputs(s); continue; }      /* a flushed DPB may hold pictures (H264SwDecApi.c:417-424) */
        if (rc == H264BSD_MEMALLOC_ERROR) { s->failed = 1; break; }
        if (nread == 0 && rc != H264BSD_RDY) { s->failed = 1; break; } /* no progress on an error: give up on this stream */
    }
    if (!s->flushed) { s->flushed = 1; h264bsdFlushBuffer(&s->st); pop_outputs(s); }
    s->finished = 1;
    return 0;
}

static void consume_prev(worker_t *w, rstream_t *s, uint32_t stream_index)
{
    runner_t *r = w->r;
    int i;
    for (i = 0; i < s->n_prev; i++) {
        pending_t *q = &s->prev[i];
        double t0 = now_s();
        u32 rc = h264b200PictureWait(&s->st, q->ticket);
        w->wait_s += now_s() - t0;
        if (rc == 0xffffffffu || rc == 0xfffffffeu) { s->failed = 1; continue; }
        if (!s->width) { s->width = 16 * h264bsdPicWidth(&s->st); s->height = 16 * h264bsdPicHeight(&s->st); }
        if (r->cb) r->cb(r->user, stream_index, s->out_index, q->ptr, s->width, s->height, q->pic_id, q->err);
        h264b200PictureRelease(&s->st, q->ticket);      /* the frame slot's host mirror may be overwritten from now on */
        s->out_index++;
        w->pictures++; w->bytes_out += (uint64_t)s->width * s->height * 3 / 2; w->err_mbs += q->err;
    }
    s->n_prev = 0;
}

/* one more picture of group g's round is parsed; the last one launches the group (and whatever of the other
 * group is already queued) */
static void group_arrive(runner_t *r, uint32_t g, uint32_t round, uint32_t produced)
{
    uint32_t k, all_dead = 1;
    pthread_mutex_lock(&r->mu);
    r->produced[g] += produced;
    if (++r->arrived[g] == r->gstart[g + 1] - r->gstart[g]) {
        h264b200EngineSubmit(r->e);
        r->rounds++;
        if (!r->produced[g]) r->dead[g] = 1;
        r->produced[g] = 0; r->arrived[g] = 0;
        for (k = 0; k < r->n_groups; k++) all_dead &= r->dead[k];
        if (all_dead) __atomic_store_n(&r->stop, 1, __ATOMIC_RELEASE);
        __atomic_store_n(&r->launched[g], round + 1, __ATOMIC_RELEASE);
    }
    pthread_mutex_unlock(&r->mu);
}

static void *worker_main(void *arg)
{
    worker_t *w = (worker_t *)arg; runner_t *r = w->r;
    while (!__atomic_load_n(&r->stop, __ATOMIC_ACQUIRE)) {
        const uint64_t item = __atomic_fetch_add(&r->next_item, 1, __ATOMIC_RELAXED);
        const uint32_t round = (uint32_t)(item / r->n_streams), idx = (uint32_t)(item % r->n_streams);
        const uint32_t g = (r->n_groups > 1 && idx >= r->gstart[1]) ? 1 : 0;
        rstream_t *s = &r->s[idx];
        uint32_t produced = 0;
        /* the stream's previous picture must have been launched (this also hands the decoder state, which is
         * single-threaded, from the thread that parsed that picture to this one) */
        while (__atomic_load_n(&r->launched[g], __ATOMIC_ACQUIRE) < round) {
            if (__atomic_load_n(&r->stop, __ATOMIC_ACQUIRE)) return NULL;
            sched_yield();
        }
        /* What the previous round popped was launched a whole round ago.  It is consumed BEFORE the next picture is
         * parsed: the DPB may hand that picture the very frame slot whose host mirror the callback is about to read,
         * and the engine holds a picture back while an output of its slot is unreleased. */
        consume_prev(w, s, idx);
        if (!s->finished && !s->failed) {
            double t0 = now_s();
            produced = (uint32_t)step_stream(s);
            w->parse_s += now_s() - t0;
        }
        memcpy(s->prev, s->cur, (size_t)s->n_cur * sizeof(pending_t)); s->n_prev = s->n_cur; s->n_cur = 0;
        group_arrive(r, g, round, produced);
    }
    return NULL;
}

/* ------------------------------------------------ device-parse pipeline */
/* With kernel Kp the host does no slice-data parsing: a picture costs it a NAL scan, a slice header and the DPB
 * bookkeeping.  So every stream SCANS AHEAD of the GPU by up to `depth` pictures (they wait in the engine), which is what
 * gives Kp thousands of independent pictures in flight.  The pipeline is FREE-RUNNING: worker threads sweep over their
 * own streams — hand completed pictures to the callback, release them, top the look-ahead up — and never block on the
 * GPU; one scheduling thread polls h264b200EngineDrive, which launches Kp whenever SMs of its share are free and a
 * reconstruction round whenever the streams' oldest pictures are parsed, a few rounds deep.  (The first version moved
 * in lock step — all threads swept all streams, then ONE scheduling step — and its ~1000 copy / launch calls per step
 * were a serial section the whole pipeline waited for: 10.5k frames/s where the device alone replays 13.9k.) */
static void outq_push(rstream_t *s, const pending_t *q)
{
    if (s->outq_n == s->outq_cap) {
        u32 ncap = s->outq_cap ? s->outq_cap * 2 : 64, i;
        pending_t *n = (pending_t *)malloc(ncap * sizeof *n);
        if (!n) { s->failed = 1; return; }
        for (i = 0; i < s->outq_n; i++) n[i] = s->outq[(s->outq_head + i) % s->outq_cap];
        free(s->outq); s->outq = n; s->outq_cap = ncap; s->outq_head = 0;
    }
    s->outq[(s->outq_head + s->outq_n++) % s->outq_cap] = *q;
}
static void dev_pop_outputs(rstream_t *s)
{
    pending_t q; u32 idr;
    while ((q.ptr = h264b200NextOutputPictureAsync(&s->st, &q.pic_id, &idr, &q.err, &q.ticket)) != NULL) outq_push(s, &q);
}
/* scan one picture ahead (NAL units up to the next H264BSD_PIC_RDY); 1 if a picture was queued */
static int dev_scan_one(rstream_t *s)
{
    while (s->pos < s->len) {
        u32 nread = 0, rest = (u32)(s->len - s->pos > 0x7fffffffu ? 0x7fffffffu : s->len - s->pos);
        u32 rc = h264bsdDecode(&s->st, s->buf + s->pos, rest, s->pic_id, &nread);
        s->pos += nread;
        if (rc == H264BSD_PIC_RDY) { s->pic_id++; dev_pop_outputs(s); return 1; }
        if (rc == H264BSD_HDRS_RDY) { dev_pop_outputs(s); continue; }
        if (rc == H264BSD_MEMALLOC_ERROR) { s->failed = 1; break; }
        if (nread == 0 && rc != H264BSD_RDY) { s->failed = 1; break; }
    }
    if (!s->flushed) {          /* the last picture has no following access unit to end it */
        const u32 before = h264b200PicturesPending(&s->st);
        s->flushed = 1; h264bsdFlushBuffer(&s->st); dev_pop_outputs(s);
        s->finished = 1;
        return h264b200PicturesPending(&s->st) > before;
    }
    s->finished = 1;
    return 0;
}
/* Hand over the outputs whose pictures are COMPLETE in host memory, oldest first; never blocks.  Returns how many. */
static u32 dev_consume(worker_t *w, rstream_t *s, uint32_t stream_index)
{
    runner_t *r = w->r;
    u32 n = 0;
    while (s->outq_n) {
        pending_t *q = &s->outq[s->outq_head];
        h264b200_picstat_t ps;
        u32 rc;
        const u32 state = h264b200PictureState(&s->st, q->ticket);
        if (state == 1 || state == 2) break;                      /* in flight, or still queued in the engine */
        rc = state == 0 ? h264b200PictureWait(&s->st, q->ticket) : 0xffffffffu;   /* complete: returns at once */
        if (rc == H264B200_WAIT_NOT_LAUNCHED) break;
        if (rc == 0xffffffffu || rc == 0xfffffffeu) s->failed = 1;
        else if (!h264b200PictureStatus(&s->st, q->ticket, &ps) && (ps.flags & H264B200_PS_DROPPED)) ;   /* incomplete last picture: not output */
        else {
            if (!h264b200PictureStatus(&s->st, q->ticket, &ps)) q->err = ps.err_mbs;
            if (!s->width) { s->width = 16 * h264bsdPicWidth(&s->st); s->height = 16 * h264bsdPicHeight(&s->st); }
            if (r->cb) r->cb(r->user, stream_index, s->out_index, q->ptr, s->width, s->height, q->pic_id, q->err);
            s->out_index++;
            w->pictures++; w->bytes_out += (uint64_t)s->width * s->height * 3 / 2; w->err_mbs += q->err;
        }
        h264b200PictureRelease(&s->st, q->ticket);
        s->outq_head = (s->outq_head + 1) % s->outq_cap; s->outq_n--;
        n++;
    }
    return n;
}

static void idle_wait(unsigned us) { struct timespec t; t.tv_sec = 0; t.tv_nsec = (long)us * 1000L; nanosleep(&t, NULL); }

/* A worker owns a fixed set of streams (a decoder instance is single-threaded) and sweeps over them for
 * as long as one of them is alive: collect what has arrived, release it, top the look-ahead up.  Nothing in a sweep
 * blocks on the GPU; a sweep that found nothing to do sleeps for a moment. */
static void *dev_worker_main(void *arg)
{
    worker_t *w = (worker_t *)arg; runner_t *r = w->r;
    uint32_t sweep = 0, quiet = 0;
    double last_collect = 0, last_progress = now_s();
    for (;; sweep++) {
        uint32_t live = 0, activity = 0, idx;
        uint32_t oi;
        for (oi = 0; oi < w->n_own; oi++) {
            rstream_t *s = &r->s[(idx = w->own[oi])];
            uint32_t burst;
            if (!s->inited || s->done) continue;
            live++;
            if (sweep < 4 && !s->host_parse) {                    /* the instances' buffers may hold less than was asked for */
                const uint32_t wnd = h264b200EngineWindow(r->e);
                if (wnd < s->depth) { s->depth = wnd; if (s->chunk > (wnd >= 4 ? wnd / 2 : 1)) s->chunk = wnd >= 4 ? wnd / 2 : 1; }
            }
            activity += dev_consume(w, s, idx);
            burst = s->chunk;
            while (burst-- && !s->finished && !s->failed && h264b200PicturesPending(&s->st) < s->depth) {
                double t0 = now_s(), t1;
                activity += 1 + (uint32_t)dev_scan_one(s);
                t1 = now_s();
                w->parse_s += t1 - t0;
                /* A picture cannot be launched while an older output of its frame slot is unreleased, so collecting must
                 * not wait for the scanning of a whole sweep (tens of milliseconds: the rounds went out with ~15 % of
                 * the streams missing): every half millisecond of scanning, a pass over all own streams that only
                 * collects — a few atomic reads per stream. */
                if (t1 - last_collect > 0.0005) {
                    uint32_t j;
                    for (j = 0; j < w->n_own; j++) { rstream_t *o = &r->s[w->own[j]]; if (o->inited && !o->done && o->outq_n) activity += dev_consume(w, o, w->own[j]); }
                    last_collect = now_s();
                }
            }
            if ((s->finished || s->failed) && !s->outq_n && !h264b200PicturesPending(&s->st)) s->done = 1;
        }
        if (!live) break;
        if (__atomic_load_n(&r->stop, __ATOMIC_ACQUIRE)) {        /* the scheduling thread saw no progress for seconds: give up on what is left */
            for (oi = 0; oi < w->n_own; oi++) { rstream_t *o = &r->s[w->own[oi]]; if (o->inited && !o->done) { o->failed = 1; o->done = 1; } }
            break;
        }
        if (r->inline_drive) {                                    /* one thread for everything: a scheduling step per sweep */
            u32 kp = 0;
            const u32 n = h264b200EngineDrive(r->e, !activity, &kp);
            if (n) r->rounds++;
            if (n || kp || activity) { quiet = 0; last_progress = now_s(); }
            else if (++quiet > 1000 && now_s() - last_progress > STALL_SECONDS) __atomic_store_n(&r->stop, 1, __ATOMIC_RELEASE);
            if (!n && !kp && !activity) idle_wait(50);
            continue;
        }
        if (activity) __atomic_fetch_add(&r->activity, activity, __ATOMIC_RELEASE);
        else { double t0 = now_s(); idle_wait(300); w->wait_s += now_s() - t0; }
    }
    __atomic_fetch_sub(&r->workers_alive, 1, __ATOMIC_RELEASE);
    return NULL;
}

/* The scheduling thread: drives the engine while the workers feed it.  When neither it nor any worker has made progress
 * for a couple of milliseconds — the look-ahead windows are full, or the streams are ending and fewer pictures than a
 * launch is normally worth are left — it tells the engine so (`idle`), which then launches whatever is ready once the
 * device has nothing else to do. */
static void *dev_driver_main(void *arg)
{
    runner_t *r = (runner_t *)arg;
    uint32_t last_activity = 0, quiet = 0;
    double last_progress = now_s();
    while (__atomic_load_n(&r->workers_alive, __ATOMIC_ACQUIRE) > 0) {
        u32 kp = 0;
        u32 n = h264b200EngineDrive(r->e, quiet >= 8, &kp);
        if (n) r->rounds++;
        if (n || kp) { quiet = 0; last_progress = now_s(); continue; }
        {
            const uint32_t a = __atomic_load_n(&r->activity, __ATOMIC_ACQUIRE);
            if (a != last_activity) { last_activity = a; quiet = 0; last_progress = now_s(); } else quiet++;
        }
        if (quiet > 1000 && now_s() - last_progress > STALL_SECONDS) __atomic_store_n(&r->stop, 1, __ATOMIC_RELEASE);   /* minutes without any progress anywhere: give up instead of hanging */
        idle_wait(250);
    }
    return NULL;
}

/* How many of the streams of a device-parse run are parsed by worker threads instead of kernel Kp
 * (h264b200SetHostParse).  Both parsers write the same records, so a stream can take either; a host-parsed picture
 * costs a thread h = 4.1 ms (and 12x the upload of its slices), a Kp-parsed one costs the GPU ~45 us of all its SMs next
 * to 22 us of reconstruction.  The share has workers of its own (all but three, see h264b200DecodeStreams); each of them
 * keeps up with  round time / h  streams, and the round time on the development box is ~20 ms.  H264B200_HOST_STREAMS=n
 * sets the share, "auto" sizes it from the worker count (80 % of what the parsing workers can sustain, a third of the
 * streams at most); default: see host_share_default. */
static uint32_t host_share(uint32_t n_streams, uint32_t n_threads)
{
    const char *env = getenv("H264B200_HOST_STREAMS");
    const uint32_t n_workers = n_threads > 1 ? n_threads - 1 : 1;
    uint32_t h;
    if (!env) env = H264B200_HOST_SHARE_DEFAULT;
    if (strcmp(env, "auto")) { const long v = atol(env); return v <= 0 ? 0 : (uint32_t)v > n_streams ? n_streams : (uint32_t)v; }
    if (n_workers < 6 || n_streams < 4 * n_threads) return 0;      /* too few threads to spare any, or too few streams to need it */
    h = (uint32_t)(0.8 * (double)(n_workers - 3) * 20.0 / 4.1);
    return h > n_streams / 3 ? n_streams / 3 : h;
}

int h264b200DecodeStreams(h264b200_engine_t *e, const h264b200_stream_t *streams, uint32_t n_streams,
                          uint32_t n_threads, h264b200_picture_cb cb, void *user, h264b200_run_stats_t *out)
{
    runner_t r; worker_t *w; uint32_t i, depth = 1, chunk = 1, n_host = 0; double t0 = now_s(); int rc = 0, dev;
    if (!e || !streams || !n_streams) return -1;
    memset(&r, 0, sizeof r);
    if (!n_threads) { long n = sysconf(_SC_NPROCESSORS_ONLN); n_threads = n > 0 ? (uint32_t)n : 1; }
    if (n_threads > n_streams) n_threads = n_streams;
    r.e = e; r.n_streams = n_streams; r.n_threads = n_threads; r.cb = cb; r.user = user;
    r.s = (rstream_t *)calloc(n_streams, sizeof(rstream_t));
    w = (worker_t *)calloc(n_threads, sizeof(worker_t));
    if (!r.s || !w) { free(r.s); free(w); return -1; }
    pthread_mutex_init(&r.mu, NULL);
    dev = (h264b200EngineFlags(e) & H264B200_ENGINE_DEVICE_PARSE) != 0;
    if (dev) {
        /* Look-ahead per stream and pictures per stream in one Kp launch; H264B200_WINDOW / H264B200_KP_CHUNK override. */
        const char *wenv = getenv("H264B200_WINDOW"), *cenv = getenv("H264B200_KP_CHUNK");
        const uint32_t slots = h264b200EngineParseSlots(e);
        uint32_t n_dev;
        n_host = host_share(n_streams, n_threads);
        n_dev = n_streams > n_host ? n_streams - n_host : 1;
        if (slots >= n_dev) {
            /* exclusive Kp launches: `slots / n_dev` pictures of every device-parsed stream are being parsed at any time (a
             * picture takes ~0.22 s whatever the load); small launches (2 pictures per stream) keep the flow even and the window
             * small: pictures in Kp + the next launch being scanned + those parsed and waiting for their round */
            chunk = 2;
            depth = slots / n_dev + chunk + 4;
        } else { chunk = 4; depth = 16; }
        if (wenv && atoi(wenv) > 0) depth = (uint32_t)atoi(wenv);
        if (cenv && atoi(cenv) > 0) chunk = (uint32_t)atoi(cenv);
        if (chunk > depth) chunk = depth;
        h264b200EngineSetStreams(e, n_dev);
        h264b200EngineSetWindow(e, depth, n_dev * chunk);
        depth = h264b200EngineWindow(e);
    }
    /* two groups once every thread has a few streams per group; otherwise one (a round is then one batch).  Batches
     * retained for h264b200EngineReplay are always whole rounds. */
    r.n_groups = (!dev && n_streams >= 4 * n_threads && !(h264b200EngineFlags(e) & H264B200_ENGINE_RETAIN)) ? 2 : 1;
    r.gstart[0] = 0; r.gstart[1] = r.n_groups > 1 ? n_streams / 2 : n_streams; r.gstart[2] = n_streams;
    for (i = 0; i < n_streams; i++) {
        rstream_t *s = &r.s[i];
        s->len = streams[i].len;
        /* decoded straight from the caller's bytes: the instance is told never to write to them (a NAL unit with
         * emulation prevention bytes is unescaped into decoder-owned memory) */
        s->buf = (uint8_t *)(uintptr_t)streams[i].data;
        s->src = streams[i].data;
        if (!s->buf || h264b200InitOnEngine(&s->st, 0, e) != HANTRO_OK) { s->failed = 1; rc = -1; continue; }
        h264b200SetReadOnlyInput(&s->st, 1);
        s->inited = 1; s->depth = depth; s->chunk = chunk;
        /* the host share (off by default), spread evenly over the stream indices */
        if (dev && (uint32_t)(((uint64_t)(i + 1) * n_host) / n_streams) != (uint32_t)(((uint64_t)i * n_host) / n_streams)) {
            h264b200SetHostParse(&s->st, 1);
            s->host_parse = 1; s->depth = 2; s->chunk = 1;     /* one picture per round, one queued ahead (the input ring holds three) */
        }
    }
    for (i = 0; i < n_threads; i++) { w[i].r = &r; w[i].tid = i; }
    if (dev) {
        /* free-running pipeline: one thread schedules the engine (dev_driver_main), the calling thread and the others are
         * workers over their own streams; with a single thread the worker takes the scheduling steps itself */
        pthread_t drv;
        uint32_t scan_workers, k_dev = 0, k_host = 0;
        r.n_workers = n_threads > 1 ? n_threads - 1 : 1;
        r.inline_drive = n_threads == 1;
        r.workers_alive = r.n_workers;
        /* Who owns which stream.  Scanning a picture is ~0.2 ms, parsing one on the host ~4 ms: a worker that parses would
         * serve device-parsed streams of its own late (their outputs uncollected, their rounds short), so the host share has
         * workers of its own and three workers keep all the device-parsed streams (measured: 3 reach what 15 reach). */
        scan_workers = (n_host && r.n_workers >= 6) ? 3 : r.n_workers;
        for (i = 0; i < r.n_workers; i++) { w[i].own = (uint32_t *)malloc((n_streams + 1) * sizeof(uint32_t)); w[i].n_own = 0; if (!w[i].own) { rc = -1; r.n_workers = i; r.workers_alive = i; break; } }
        for (i = 0; i < n_streams && r.n_workers; i++) {
            worker_t *o;
            if (r.s[i].host_parse && scan_workers < r.n_workers) o = &w[scan_workers + k_host++ % (r.n_workers - scan_workers)];
            else o = &w[k_dev++ % scan_workers];
            o->own[o->n_own++] = i;
        }
        for (i = 1; i < r.n_workers; i++) pthread_create(&w[i].th, NULL, dev_worker_main, &w[i]);
        if (r.inline_drive) dev_worker_main(&w[0]);
        else {
            pthread_create(&drv, NULL, dev_driver_main, &r);
            dev_worker_main(&w[0]);
            pthread_join(drv, NULL);
        }
        for (i = 1; i < r.n_workers; i++) pthread_join(w[i].th, NULL);
        for (i = 0; i < r.n_workers; i++) free(w[i].own);
        while (h264b200EngineSubmit(e)) ;                                          /* nothing should be left; be safe */
        for (i = 0; i < n_streams; i++) if (r.s[i].inited && (r.s[i].outq_n || !r.s[i].done)) r.s[i].failed = 1;
    } else {
        for (i = 1; i < n_threads; i++) pthread_create(&w[i].th, NULL, worker_main, &w[i]);
        worker_main(&w[0]);
        for (i = 1; i < n_threads; i++) pthread_join(w[i].th, NULL);
        for (i = 0; i < n_streams; i++) consume_prev(&w[0], &r.s[i], i);           /* what the last round popped */
    }
    h264b200EngineSync(e);
    if (out) {
        memset(out, 0, sizeof *out);
        for (i = 0; i < n_threads; i++) {
            out->pictures += w[i].pictures; out->bytes_out += w[i].bytes_out; out->err_mbs += w[i].err_mbs;
            out->parse_seconds += w[i].parse_s; out->wait_seconds += w[i].wait_s;
        }
        for (i = 0; i < n_streams; i++) { out->bytes_in += r.s[i].len; if (r.s[i].failed) out->failed_streams++; }
        out->rounds = r.rounds; out->threads = n_threads;
        for (i = 0; i < n_streams; i++) out->host_streams += (uint32_t)r.s[i].host_parse;
    }
    for (i = 0; i < n_streams; i++) {
        if (r.s[i].failed) rc = -1;
        if (r.s[i].inited) h264bsdShutdown(&r.s[i].st);
        free(r.s[i].outq);
    }
    pthread_mutex_destroy(&r.mu);
    free(r.s); free(w);
    if (out) out->seconds = now_s() - t0;
    return rc;
}

/* ------------------------------------------------------------ GOP splitter */
/* next start code prefix (00 00 01) at or after i; returns len if none.  *sc_len = 3 or 4 (leading zero byte) */
static size_t next_start_code(const uint8_t *d, size_t len, size_t i, size_t *prefix_start)
{
    while (i + 3 <= len) {
        const uint8_t *z = (const uint8_t *)memchr(d + i, 0, len - i);
        if (!z) break;
        i = (size_t)(z - d);
        if (i + 3 > len) break;
        if (d[i + 1] == 0) {
            size_t j = i + 2;
            while (j < len && d[j] == 0) j++;
            if (j < len && d[j] == 1) { *prefix_start = i; return j + 1; }
            i = j;
        } else i += 2;
    }
    *prefix_start = len;
    return len;
}

int h264b200SplitGops(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap,
                      size_t *seg_off, size_t *seg_len, uint32_t max_segs)
{
    size_t pre, nal, hdr_cap = 4096, hdr_len = 0, hdr_len_au = 0, au_start = (size_t)-1, cut_prev = (size_t)-1, o = 0;
    uint8_t *hdr = (uint8_t *)malloc(hdr_cap), *hdr_at_cut = NULL;
    size_t hdr_at_cut_len = 0;
    int n = 0;
    if (!hdr) return -1;
    nal = next_start_code(data, len, 0, &pre);
    while (nal < len) {
        size_t next_pre, next_nal = next_start_code(data, len, nal, &next_pre);
        int type = data[nal] & 31;
        if (au_start == (size_t)-1) { au_start = pre; hdr_len_au = hdr_len; }              /* first NAL after a VCL NAL opens a candidate access unit */
        if (type == 5 && nal + 1 < len && (data[nal + 1] & 0x80)) { /* IDR slice with first_mb_in_slice == 0: cut at its access unit */
            if (cut_prev != (size_t)-1) {
                size_t body = au_start - cut_prev;
                if ((uint32_t)n >= max_segs || o + hdr_at_cut_len + body > out_cap) { free(hdr); free(hdr_at_cut); return -1; }
                seg_off[n] = o;
                memcpy(out + o, hdr_at_cut, hdr_at_cut_len); o += hdr_at_cut_len;
                memcpy(out + o, data + cut_prev, body); o += body;
                seg_len[n] = o - seg_off[n]; n++;
            }
            cut_prev = au_start;
            free(hdr_at_cut);
            hdr_at_cut = (uint8_t *)malloc(hdr_len_au + 1); hdr_at_cut_len = hdr_len_au;   /* sets sent BEFORE this access unit */
            if (!hdr_at_cut) { free(hdr); return -1; }
            memcpy(hdr_at_cut, hdr, hdr_len_au);
        }
        if (type == 7 || type == 8) {                              /* remember parameter sets (sent before the access unit that needs them) */
            size_t l = next_pre - pre;
            if (hdr_len + l > hdr_cap) { hdr_cap = (hdr_len + l) * 2; hdr = (uint8_t *)realloc(hdr, hdr_cap); if (!hdr) { free(hdr_at_cut); return -1; } }
            memcpy(hdr + hdr_len, data + pre, l); hdr_len += l;
        }
        if (type == 1 || type == 5) au_start = (size_t)-1;         /* a VCL NAL closes the candidate */
        pre = next_pre; nal = next_nal;
    }
    if (cut_prev != (size_t)-1) {
        size_t body = len - cut_prev;
        if ((uint32_t)n >= max_segs || o + hdr_at_cut_len + body > out_cap) { free(hdr); free(hdr_at_cut); return -1; }
        seg_off[n] = o;
        memcpy(out + o, hdr_at_cut, hdr_at_cut_len); o += hdr_at_cut_len;
        memcpy(out + o, data + cut_prev, body); o += body;
        seg_len[n] = o - seg_off[n]; n++;
    }
    free(hdr); free(hdr_at_cut);
    return n;
}
