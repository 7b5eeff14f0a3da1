/* h264b200_batch.h — batch-level entry points of libh264b200.so that have no
 * direct counterpart in the reference: they exist because ONE 1080p picture is
 * about a microsecond of HBM traffic, so a B200 is only used well when many
 * independent pictures (one per stream / IDR-bounded GOP segment) go through
 * each kernel launch.  The nearest reference code is the multi-instance test
 * bench (Decoder/src/TestBenchMultipleInstance.c:60-350: N decoder instances
 * stepped round-robin in one thread); h264b200DecodeStreams is that loop with
 * worker threads for the serial CAVLC parse and one batched GPU launch per
 * round.
 */
#ifndef H264B200_BATCH_H
#define H264B200_BATCH_H
#include "h264b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* engine flags */
#define H264B200_ENGINE_BATCHED 1u  /* pictures wait in the engine until h264b200EngineSubmit */
#define H264B200_ENGINE_RETAIN  2u  /* keep records + coefficient levels of every batch in HBM for h264b200EngineReplay */
#define H264B200_ENGINE_NO_D2H  4u  /* do not copy finished frames to the host mirrors (kernel-only measurements) */
#define H264B200_ENGINE_DEVICE_PARSE 8u /* instances parse slice data on the device (kernel Kp, include/h264b200_slices.h): the host
                                       keeps NAL extraction, parameter sets, slice headers and the DPB; a picture ends when the next
                                       access unit begins or at h264bsdFlushBuffer, so H264BSD_PIC_RDY comes one NAL unit later
                                       (with *readBytes == 0, the reference's own re-feed protocol, h264bsd_decoder.c:267-268) */

#define H264B200_ENGINE_NO_RECON 16u /* parity aid: reconstruction rounds launch no K1..K4, so that what kernel Kp wrote (levels, not yet
                                       transformed in place by K1) can be read back with h264b200DebugFetchParse; frames are garbage */
#define H264B200_ENGINE_TAP_PREDEBLOCK 32u /* parity aid: every picture is copied aside between K3 and K4 (the reference's picture at its
                                       h264bsdFilterPicture call, h264bsd_decoder.c:489-491); read it with h264b200DebugFetchPredeblock */

h264b200_engine_t *h264b200EngineCreateEx(int device, uint32_t flags);
void h264b200EngineSetFlags(h264b200_engine_t *e, uint32_t flags);
uint32_t h264b200EngineFlags(h264b200_engine_t *e);

/* Non-blocking variant of h264bsdNextOutputPicture (h264bsd_decoder.c:642): pops
 * the display queue and returns the host address the picture WILL occupy; the
 * samples are valid after h264b200PictureWait(ticket) returned 0.  Lets a
 * caller that drives many instances defer the wait until the batch holding the
 * picture has been launched. */
u8  *h264b200NextOutputPictureAsync(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs, u32 *ticket);
/* 0: picture complete; otherwise the engine error flags, 0xffffffff on a CUDA failure, or 0xfffffffe when the
 * caller waited so long that a later picture was already reconstructed into the same frame buffer. */
u32  h264b200PictureWait(storage_t *pStorage, u32 ticket);
/* never blocks: 0 complete, 1 launched and in flight, 2 still queued in the engine, 0xffffffff error */
u32  h264b200PictureState(storage_t *pStorage, u32 ticket);
#define H264B200_WAIT_NOT_LAUNCHED 0xfffffffdu   /* the picture is still queued in the engine (device-parse look-ahead): call h264b200EngineAdvance */
/* The caller is done with the picture behind `ticket`: its host buffer may be overwritten by a later picture of the
 * same frame slot.  Until then the engine holds that later picture back (it is launched by a later
 * h264b200EngineAdvance).  Only needed after h264b200NextOutputPictureAsync. */
void h264b200PictureRelease(storage_t *pStorage, u32 ticket);
/* status words of a device-parsed picture (concealed macroblocks ...), valid after h264b200PictureWait returned 0; returns 0 on success */
#include "h264b200_slices.h"
u32  h264b200PictureStatus(storage_t *pStorage, u32 ticket, h264b200_picstat_t *out);
/* pictures of this instance handed to the engine but not yet launched */
u32  h264b200PicturesPending(storage_t *pStorage);
/* 1 if the instance parses slice data on the device */
u32  h264b200DeviceParse(storage_t *pStorage);
/* on != 0: the instance parses slice data on the host (h264_slice.c) even when its engine is a device-parse engine; call
 * before the first picture.  Host- and device-parsed instances share the reconstruction rounds of the engine. */
void h264b200SetHostParse(storage_t *pStorage, u32 on);
/* on != 0: h264bsdDecode never writes to byteStrm (NAL units with emulation prevention bytes are unescaped into
 * decoder-owned memory instead of in place); h264b200DecodeStreams uses it to decode straight from the caller's streams */
void h264b200SetReadOnlyInput(storage_t *pStorage, u32 on);
/* One scheduling step of the engine: (1) if enough queued device-parse pictures have accumulated (or a stream would
 * otherwise stall) launch kernel Kp over all of them; (2) launch ONE round of reconstruction: the oldest queued picture
 * of every instance whose output buffer is free.  Returns the number of pictures reconstructed by this call.
 * h264b200EngineSubmit repeats this until nothing is left that can be launched. */
u32  h264b200EngineAdvance(h264b200_engine_t *e);
/* One step of the free-running schedule, for a caller that dedicates a thread to the engine (h264b200DecodeStreams does):
 * call it a few thousand times per second.  Launches kernel Kp as soon as `parse_threshold` unparsed pictures are queued
 * and SMs of Kp's share are free (sized to those SMs, oldest picture of every instance first, whole levels), and ONE
 * reconstruction round over the pictures whose Kp launch has finished once at least 7 of 8 instances with queued pictures
 * are ready, while fewer than three rounds are on the device.  idle != 0 = the caller's other threads have had nothing to
 * do for a while (look-ahead windows full, or the streams are ending): when nothing is running on the device either, the
 * "enough to be worth it" conditions are dropped.  *kp_pictures (may be NULL) = pictures handed to Kp by this call;
 * returns the pictures of the round it launched, 0 if none. */
u32  h264b200EngineDrive(h264b200_engine_t *e, int idle, u32 *kp_pictures);
/* look-ahead of the device-parse path: how many pictures per instance may be queued (parse buffers are allocated for
 * depth + 2), and how many queued pictures make kernel Kp worth launching.  Call before the instances are created. */
void h264b200EngineSetWindow(h264b200_engine_t *e, uint32_t depth, uint32_t parse_threshold);
uint32_t h264b200EngineWindow(h264b200_engine_t *e);   /* the look-ahead the instances can actually hold: <= the depth asked for */
/* number of instances about to share the engine: bounds the device memory each spends on look-ahead buffers (call before they are created) */
void h264b200EngineSetStreams(h264b200_engine_t *e, uint32_t n_streams);
/* pictures one launch of kernel Kp parses at full rate (one per warp of the SMs the launch owns); 0: no such limit */
uint32_t h264b200EngineParseSlots(h264b200_engine_t *e);

/* ---- output formatting on the device (K5) ----
 * H264B200_OUT_I420 (default): the reference's output, uncropped MB-aligned planar I420.
 * H264B200_OUT_RGBA: cropped to the SPS cropping rectangle (the reference only reports it,
 * h264bsd_decoder.c:886-917) and converted with the wrapper's BT.601 fixed-point formula
 * (templates/DecoderPost.js:514-560): cropWidth*cropHeight*4 bytes, R,G,B,255 per sample.  The picture
 * pointers returned afterwards address that buffer.  May be called before or after the headers. */
#define H264B200_OUT_I420 0u
#define H264B200_OUT_RGBA 1u
u32  h264b200SetOutputFormat(storage_t *pStorage, u32 format);

/* ---- resident replay (measurement): re-run K1..K4 over every retained batch,
 * inputs already in HBM, no host<->device copies.  Asynchronous; returns the
 * number of pictures enqueued.  Requires H264B200_ENGINE_RETAIN while decoding. */
u32  h264b200EngineReplay(h264b200_engine_t *e, u32 reps, int time_kernels);
/* device time (ms, CUDA events on the engine's compute stream) of the last h264b200EngineReplay; blocks until it is done */
double h264b200EngineReplayMs(h264b200_engine_t *e);
void h264b200EngineDropRetained(h264b200_engine_t *e);
/* number of frame slots whose device content differs from the host mirror the
 * normal decode filled (0 = the replay reproduced the same pictures) */
u32  h264b200EngineCheckResident(h264b200_engine_t *e);
/* accumulated CUDA-event time, algorithmic bytes and launches per
 * kernel family [K1 transform, K2 inter, K3 intra, K4 deblock, Kp slice-data parse] of timed replays.  Kp runs on its
 * own stream concurrently with the reconstruction rounds of earlier pictures: its time is that of its launches, not a
 * share of the step. */
typedef struct { double ms[5]; uint64_t bytes[5]; uint64_t launches[5]; } h264b200_kernel_times_t;
void h264b200EngineKernelTimes(h264b200_engine_t *e, h264b200_kernel_times_t *out, int reset);

/* Parity / debugging aid: records, coefficient slots and findings of kernel Kp for the picture this device-parse instance
 * submitted most recently (back = 0) or `back` pictures earlier (while its parse buffer has not been re-used).  mbs:
 * PicWidth*PicHeight records; coef: up to coef_cap slots of 16 int16; res: 12 uint32 {coef_used, n_intra, n_inter,
 * any_deblock, n_conceal, conceal_offset, -, -, err_mbs, flags, decoded_mbs, coef_slots}.  Blocks until the engine is idle.
 * Returns 0, or a negative value when the picture is not available. */
int  h264b200DebugFetchParse(storage_t *pStorage, int back, void *mbs, int16_t *coef, uint32_t coef_cap, uint32_t *res);
/* the picture most recently launched for this instance as it was BEFORE deblocking (engine flag H264B200_ENGINE_TAP_PREDEBLOCK);
 * returns the number of bytes written to out (width*height*3/2) or a negative value */
long h264b200DebugFetchPredeblock(storage_t *pStorage, uint8_t *out, size_t cap);

/* ---- many streams through one engine ---- */
typedef struct { const uint8_t *data; size_t len; } h264b200_stream_t;   /* Annex-B; not modified (copied internally) */
/* Called from worker threads (concurrently for different streams; in decode order
 * within a stream).  i420 is valid only during the call. */
typedef void (*h264b200_picture_cb)(void *user, uint32_t stream, uint32_t index, const uint8_t *i420,
                                    uint32_t width, uint32_t height, uint32_t pic_id, uint32_t num_err_mbs);
typedef struct {
    uint64_t pictures, bytes_in, bytes_out;
    uint32_t err_mbs, failed_streams, rounds, threads;
    double   seconds;            /* wall clock of the whole call */
    double   parse_seconds;      /* summed over threads: time inside h264bsdDecode */
    double   wait_seconds;       /* summed over threads: time blocked on the GPU */
    uint32_t host_streams;       /* device-parse engines: streams whose slice data the worker threads parsed (the host share) */
    uint32_t reserved;
} h264b200_run_stats_t;
/* Decode n_streams independent streams (or GOP segments) with n_threads parser
 * threads (0: one per online CPU).  Every round parses one picture of every live
 * stream; the pictures of a round are launched as one batch (two, one per half of
 * the streams, when there are at least four streams per thread, so that no
 * thread ever waits for a round to end; always one while H264B200_ENGINE_RETAIN is
 * set, so that a retained batch is a whole round) while the next pictures are being parsed.
 * On a device-parse engine (kernel Kp) the pipeline is free-running instead: one of the n_threads schedules the engine
 * (h264b200EngineDrive), the others sweep over their own streams — scan ahead, collect, release — and never block on
 * the GPU (four threads per GPU reach the throughput of sixteen).  H264B200_HOST_STREAMS=n|auto hands a share of the streams to the threads' own parser
 * (h264b200SetHostParse: same records, same rounds, less work for Kp; off by default).
 * `rounds` in the statistics counts the batches.  Returns 0 on success. */
int h264b200DecodeStreams(h264b200_engine_t *e, const h264b200_stream_t *streams, uint32_t n_streams,
                          uint32_t n_threads, h264b200_picture_cb cb, void *user, h264b200_run_stats_t *out);

/* Cut an Annex-B stream into self-contained segments at IDR access units
 * (an IDR empties the DPB, h264bsd_dpb.c:675-708, so nothing crosses a cut).
 * Segment i is written to out + seg_off[i], seg_len[i] bytes: every parameter
 * set NAL seen before the cut, then the stream bytes up to the next cut.
 * Returns the number of segments (<= max_segs), or -1 if out_cap is too small
 * (len * 2 + 4096 is always enough for streams whose parameter sets are sent once). */
int h264b200SplitGops(const uint8_t *data, size_t len, uint8_t *out, size_t out_cap,
                      size_t *seg_off, size_t *seg_len, uint32_t max_segs);

/* MP4 (ISO-BMFF) -> Annex-B, the C counterpart of Player/mp4.js (avcC :414-431, length-prefixed NAL
 * extraction :711-723): parameter sets of the first avc1 track, then every sample's NAL units, each
 * behind a 00 00 00 01 start code.  out_cap of len + 64 KiB is always enough.  Returns the number of
 * samples written, -1 on a malformed / unsupported file, -2 if out_cap is too small. */
long h264b200Mp4ToAnnexB(const uint8_t *mp4, size_t len, uint8_t *out, size_t out_cap, size_t *out_len);

#ifdef __cplusplus
}
#endif
#endif
