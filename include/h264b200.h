/* h264b200.h — C-ABI of libh264b200.so, the B200-native H.264 Baseline decoder.
 *
 * DROP-IN BOUNDARY.  The first block declares, with the reference's names,
 * argument meaning, return codes and ownership rules, the entry points of the
 * reference decoder core (Decoder/src/h264bsd_decoder.h:60-80 in
 * maexx393/Broadway); the callers above them (H264SwDecApi.c, Decoder.c,
 * DecTestBench.c) link unchanged.  Differences a maintainer must know:
 *
 *  - `storage_t` is OPAQUE here.  The reference exposes its 4.6 KB struct
 *    (h264bsd_storage.h:74-149) and H264SwDecApi.c pokes a few fields
 *    (storage.dpb->flushed/numOut/outIndex :417-424, activeSps :219); this
 *    library therefore also exports the whole H264SwDec* API
 *    (include/h264b200_swdec.h); Decoder.c (the broadway* shim) links against it unchanged
 *    so that nothing above the boundary needs those fields.
 *  - The macroblock-layer parse stays on the host; reconstruction (dequant +
 *    inverse transforms, inter/intra prediction, deblocking) runs as CUDA
 *    kernels on one B200.  There is NO CPU reconstruction path: without a
 *    usable CUDA device h264bsdDecode returns H264BSD_MEMALLOC_ERROR at the
 *    first slice and prints the CUDA error to stderr.
 *  - Pictures are returned as pointers into pinned HOST memory (planar I420,
 *    uncropped, 16*PicWidth x 16*PicHeight, Y then Cb then Cr — the layout
 *    Decoder.c:113-147 and DecoderPost.js:68-72 expect), valid until the next
 *    h264bsdDecode call that starts a new picture, exactly as in the reference
 *    (h264bsd_dpb.c:680, :1259).
 *
 * The second block is the batch engine (no reference equivalent; the nearest is
 * TestBenchMultipleInstance.c:60-350, a round-robin over N instances): many
 * decoder instances attached to one engine submit their finished pictures into
 * one batched kernel launch per stage, which is what lets the HBM-bound
 * kernels run near the roofline (one 1080p picture is ~1 us of HBM traffic).
 */
#ifndef H264B200_H
#define H264B200_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned char u8;    /* Decoder/inc/basetype.h:28-33 */
typedef unsigned int  u32;
typedef int           i32;

/* h264bsd_decoder.h:43-50 */
enum { H264BSD_RDY = 0, H264BSD_PIC_RDY, H264BSD_HDRS_RDY, H264BSD_ERROR, H264BSD_PARAM_SET_ERROR, H264BSD_MEMALLOC_ERROR };
#define HANTRO_OK  0
#define HANTRO_NOK 1

/* Caller-allocated handle (the reference's storage_t is caller-allocated too:
 * it is a member of decContainer_t, h264bsd_container.h).  Contents are private. */
typedef struct storage { void *impl; uint64_t reserved[7]; } storage_t;

/* ---- reference entry points: h264bsd_decoder.h:60-66 ---- */
u32  h264bsdInit(storage_t *pStorage, u32 noOutputReordering);
u32  h264bsdDecode(storage_t *pStorage, u8 *byteStrm, u32 len, u32 picId, u32 *readBytes);
u8  *h264bsdNextOutputPicture(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs);
void h264bsdShutdown(storage_t *pStorage);
/* ---- accessors: h264bsd_decoder.h:68-80 ---- */
u32  h264bsdPicWidth(storage_t *pStorage);      /* in macroblocks */
u32  h264bsdPicHeight(storage_t *pStorage);     /* in macroblocks */
u32  h264bsdVideoRange(storage_t *pStorage);
u32  h264bsdMatrixCoefficients(storage_t *pStorage);
void h264bsdCroppingParams(storage_t *pStorage, u32 *croppingFlag, u32 *left, u32 *width, u32 *top, u32 *height);
void h264bsdSampleAspectRatio(storage_t *pStorage, u32 *sarWidth, u32 *sarHeight);
u32  h264bsdCheckValidParamSets(storage_t *pStorage);
void h264bsdFlushBuffer(storage_t *pStorage);
u32  h264bsdProfile(storage_t *pStorage);

/* ---- batch engine (this library only) ---- */
typedef struct h264b200_engine h264b200_engine_t;

/* Create an engine on CUDA device `device` (-1: current device). NULL on failure
 * (reason on stderr). */
h264b200_engine_t *h264b200EngineCreate(int device);
void h264b200EngineDestroy(h264b200_engine_t *e);
/* Like h264bsdInit, but finished pictures of this instance are queued in `e`
 * instead of being launched one by one; they run when h264b200EngineSubmit is
 * called (or implicitly when the instance finishes a second picture, or when a
 * frame that is still queued is requested through h264bsdNextOutputPicture). */
u32  h264b200InitOnEngine(storage_t *pStorage, u32 noOutputReordering, h264b200_engine_t *e);
/* Launch every queued picture of every attached instance as one batch
 * (asynchronous). Returns the number of pictures launched. */
u32  h264b200EngineSubmit(h264b200_engine_t *e);
/* Block until everything submitted so far has completed. */
void h264b200EngineSync(h264b200_engine_t *e);
/* Counters since creation: kernels launched, pictures reconstructed, H2D and D2H bytes, reconstruction rounds, launches of
 * the device-side slice parser (kernel Kp) and pictures it parsed. */
typedef struct { uint64_t kernel_launches, pictures, h2d_bytes, d2h_bytes, batches, kp_launches, kp_pictures; } h264b200_stats_t;
void h264b200EngineStats(h264b200_engine_t *e, h264b200_stats_t *out);
/* Device-side error word of the last completed batch: bit 0 = a residual left
 * [-512,511] (the reference's mid-parse check, h264bsd_transform.c:181-185). */
u32  h264b200EngineErrorFlags(h264b200_engine_t *e);

/* Library/CUDA availability probe: 0 if a CUDA device is usable. */
int  h264b200Probe(char *msg, size_t msg_cap);

#ifdef __cplusplus
}
#endif
#endif
