/* oracle/engine_shim.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * The handful of engine entry points h264_runner.c calls, over the CPU restatement backend
 * (recon_cpu.c), so that the multi-threaded scheduling of h264b200DecodeStreams — work items,
 * launch groups, hand-over of a stream between parser threads, output collection one round late —
 * runs in the CPU test-suite (tests/test_runner_cpu.py) with the product's own runner source.
 * Here every picture is reconstructed synchronously when it is submitted, so "launching" a
 * group is a no-op; what the tests exercise is the host-side logic. */
#include <stdlib.h>
#include "h264b200.h"
#include "h264b200_batch.h"
#include "../broadway_b200/csrc/h264_internal.h"

struct h264b200_engine { uint32_t flags; uint32_t submits; uint32_t window; h264_backend_t be; };
h264_backend_t recon_cpu_backend(int device_parse);      /* recon_cpu.c */
uint32_t recon_cpu_advance(void *ctx);

u32 h264_decoder_create(storage_t *pStorage, u32 noOutputReordering, h264_backend_t *be);

h264b200_engine_t *h264b200EngineCreateEx(int device, uint32_t flags)
{
    h264b200_engine_t *e = (h264b200_engine_t *)calloc(1, sizeof *e);
    (void)device;
    if (e) { e->flags = flags; e->window = 1; e->be = recon_cpu_backend((flags & H264B200_ENGINE_DEVICE_PARSE) != 0);
             if (flags & H264B200_ENGINE_BATCHED) e->be.ctx = e; }   /* batched device-parse instances defer their launches to h264b200EngineAdvance */
    return e;
}
h264b200_engine_t *h264b200EngineCreate(int device) { return h264b200EngineCreateEx(device, H264B200_ENGINE_BATCHED); }
void h264b200EngineDestroy(h264b200_engine_t *e) { free(e); }
void h264b200EngineSetFlags(h264b200_engine_t *e, uint32_t flags) { if (e) { e->flags = flags; e->be.parse_mode = (flags & H264B200_ENGINE_DEVICE_PARSE) != 0; } }
uint32_t h264b200EngineFlags(h264b200_engine_t *e) { return e ? e->flags : 0; }
u32 h264b200InitOnEngine(storage_t *pStorage, u32 noOutputReordering, h264b200_engine_t *e)
{
    if (!e) return HANTRO_NOK;
    return h264_decoder_create(pStorage, noOutputReordering, &e->be);    /* recon_cpu.c, host- or device-parse (kp_cpu.cpp) as the engine flags say */
}
u32 h264b200EngineSubmit(h264b200_engine_t *e)
{
    u32 total = 0, n;
    if (!e) return 0;
    e->submits++;
    while ((n = recon_cpu_advance(e)) != 0) total += n;
    return total;
}
void h264b200EngineSync(h264b200_engine_t *e) { (void)e; }
u32 h264b200EngineAdvance(h264b200_engine_t *e) { if (!e) return 0; e->submits++; return recon_cpu_advance(e); }
u32 h264b200EngineDrive(h264b200_engine_t *e, int relaxed, u32 *kp_pictures) { (void)relaxed; if (kp_pictures) *kp_pictures = 0; if (!e) return 0; e->submits++; return recon_cpu_advance(e); }
void h264b200EngineSetWindow(h264b200_engine_t *e, uint32_t depth, uint32_t parse_threshold) { (void)parse_threshold; if (e) e->window = depth ? depth : 1; }
uint32_t h264b200EngineWindow(h264b200_engine_t *e) { return e ? e->window : 0; }
uint32_t h264b200EngineParseSlots(h264b200_engine_t *e) { (void)e; return 0; }
void h264b200EngineSetStreams(h264b200_engine_t *e, uint32_t n) { (void)e; (void)n; }
