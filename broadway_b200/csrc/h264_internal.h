/* h264_internal.h — private structures of the host side of the decoder:
 * parameter sets, slice header, decoded picture buffer bookkeeping, the
 * per-instance state behind the opaque `storage_t`, and the backend interface
 * through which finished pictures (macroblock records + coefficient slots) are
 * handed to the reconstruction engine (CUDA in the product library).
 */
#ifndef B200_H264_INTERNAL_H
#define B200_H264_INTERNAL_H
#include <stdint.h>
#include <stddef.h>
#include "h264b200_records.h"
#include "h264b200_slices.h"
#include "h264_bits.h"
#include "h264_fmo.h"

/* host-side allocations go through the embedder hooks of the API layer (H264SwDecApi.h:163-173; weak defaults in h264_swdec.c) */
void *H264SwDecMalloc(unsigned int size);
void H264SwDecFree(void *ptr);
void H264SwDecMemset(void *ptr, int value, unsigned int count);
static inline void *h264_malloc(size_t n) { return n > 0xffffffffu ? NULL : H264SwDecMalloc((unsigned int)n); }
static inline void *h264_calloc(size_t k, size_t n) { void *p = (k && n > 0xffffffffu / k) ? NULL : H264SwDecMalloc((unsigned int)(k * n)); if (p) H264SwDecMemset(p, 0, (unsigned int)(k * n)); return p; }
static inline void h264_free(void *p) { if (p) H264SwDecFree(p); }

#ifdef __cplusplus
extern "C" {
#endif

#define H264_MAX_SPS 32
#define H264_MAX_PPS 256
#define H264_MAX_REFS 16
#define H264_MAX_SLOTS 18          /* dpbSize(<=16) + current + 1 spare */

enum { NAL_SLICE = 1, NAL_IDR = 5, NAL_SEI = 6, NAL_SPS = 7, NAL_PPS = 8, NAL_AUD = 9, NAL_EOSEQ = 10, NAL_EOSTREAM = 11, NAL_FILLER = 12 };

typedef struct {
    uint8_t  valid;
    uint8_t  profile_idc, level_idc, sps_id;
    uint32_t max_frame_num;            /* 2^(log2_max_frame_num) */
    uint8_t  log2_max_frame_num;
    uint8_t  poc_type;
    uint8_t  log2_max_poc_lsb;
    uint32_t max_poc_lsb;
    uint8_t  delta_pic_order_always_zero;
    int32_t  offset_for_non_ref_pic, offset_for_top_to_bottom;
    uint32_t num_ref_frames_in_poc_cycle;
    int32_t  offset_for_ref_frame[256];
    uint32_t num_ref_frames;
    uint8_t  gaps_allowed;
    uint32_t width_mbs, height_mbs;
    uint8_t  crop_flag;
    uint32_t crop_left, crop_right, crop_top, crop_bottom;
    uint32_t max_dpb_size;
    /* VUI subset the API reports / the DPB uses (h264bsd_vui.c) */
    uint8_t  vui_present, aspect_ratio_present, aspect_ratio_idc;
    uint32_t sar_width, sar_height;
    uint8_t  video_signal_present, video_full_range, colour_desc_present, matrix_coefficients;
    uint8_t  bitstream_restriction;
    uint32_t num_reorder_frames, max_dec_frame_buffering;
} h264_sps_t;

typedef struct {
    uint8_t  valid;
    uint8_t  pps_id, sps_id;
    uint8_t  pic_order_present;
    uint32_t num_slice_groups;
    h264_fmo_t fmo;                    /* valid when num_slice_groups > 1; fmo.group_id is owned by the stored copy */
    uint32_t fmo_map_units;            /* type 6: pic_size_in_map_units */
    uint32_t num_ref_idx_l0_default;
    int32_t  pic_init_qp;
    int32_t  chroma_qp_index_offset;
    uint8_t  deblocking_control_present, constrained_intra_pred, redundant_pic_cnt_present;
} h264_pps_t;

typedef struct { uint8_t idc; uint32_t val; } h264_reorder_cmd_t;
typedef struct { uint8_t op; uint32_t diff_pic_nums, long_term_pic_num, long_term_frame_idx, max_long_term_frame_idx; } h264_mmco_t;

typedef struct {
    uint32_t first_mb;
    uint8_t  slice_type;               /* 0 P, 2 I (mod 5) */
    uint32_t pps_id, frame_num, idr_pic_id, poc_lsb;
    int32_t  delta_poc_bottom, delta_poc[2];
    uint32_t redundant_pic_cnt;
    uint32_t num_ref_idx_active;
    uint8_t  reorder_flag; uint32_t n_reorder; h264_reorder_cmd_t reorder[H264_MAX_REFS + 2];
    uint8_t  no_output_of_prior_pics, long_term_reference_flag, adaptive_marking;
    uint32_t n_mmco; h264_mmco_t mmco[36];
    int32_t  slice_qp;
    uint32_t slice_group_change_cycle;
    uint8_t  disable_deblocking_idc; int8_t alpha_off, beta_off;   /* offsets already *2 */
} h264_slice_hdr_t;

/* -------------------------------------------------------------------- DPB */
enum { PIC_UNUSED = 0, PIC_NON_EXISTING, PIC_SHORT, PIC_LONG };
typedef struct {
    int      slot;                     /* frame-pool slot (stable id of the frame storage) */
    int      status;
    int32_t  pic_num;                  /* PicNum / LongTermPicNum */
    uint32_t frame_num;
    int32_t  poc;
    uint8_t  to_be_displayed;
    uint32_t pic_id, num_err_mbs, is_idr;
} h264_dpb_pic_t;

typedef struct { int slot; uint32_t pic_id, num_err_mbs, is_idr; } h264_out_t;

typedef struct {
    h264_dpb_pic_t buf[H264_MAX_REFS + 1];   /* sorted; buf[dpb_size] is the picture being decoded */
    h264_dpb_pic_t *list[H264_MAX_REFS + 1]; /* RefPicList0 */
    h264_out_t out[H264_MAX_REFS + 2];
    uint32_t num_out, out_index;
    uint32_t max_ref_frames, dpb_size, max_frame_num, max_long_term_idx;
    uint32_t num_ref_frames, fullness, prev_ref_frame_num;
    uint8_t  no_reordering, flushed, last_has_mmco5, allocated;
    int      spare_slot;               /* a frame slot beyond dpb_size + 1, rotated in at every picture (h264_dpb_rotate_spare) */
} h264_dpb_t;

#define H264_NO_LONG_TERM 0xFFFF

void h264_dpb_init(h264_dpb_t *d, uint32_t dpb_size, uint32_t max_ref_frames, uint32_t max_frame_num, int no_reordering);
int  h264_dpb_current_slot(h264_dpb_t *d);
void h264_dpb_rotate_spare(h264_dpb_t *d);
int  h264_dpb_check_gaps(h264_dpb_t *d, uint32_t frame_num, int is_ref, int gaps_allowed);
void h264_dpb_init_ref_list(h264_dpb_t *d);
int  h264_dpb_reorder(h264_dpb_t *d, const h264_slice_hdr_t *sh);
int  h264_dpb_ref_slot(const h264_dpb_t *d, uint32_t ref_idx);  /* -1: missing / non-existing */
int  h264_dpb_mark(h264_dpb_t *d, const h264_slice_hdr_t *sh, int is_ref, int is_idr, int32_t poc, uint32_t pic_id, uint32_t num_err);
void h264_dpb_flush(h264_dpb_t *d);
const h264_out_t *h264_dpb_next_output(h264_dpb_t *d);

/* ---------------------------------------------------------------- backend */
/* One picture's worth of host-written input for the reconstruction engine. */
typedef struct {
    h264b200_mb_t *mbs;        /* width_mbs*height_mbs records (pinned in the CUDA backend) */
    int16_t  *coef;            /* coefficient slots */
    uint32_t  coef_cap;        /* capacity in slots */
    uint32_t  coef_used;       /* slots written */
    uint32_t  n_intra, n_inter;/* macroblock class counts (lets the engine skip kernels) */
    uint32_t  any_deblock;     /* some macroblock has filtering enabled */
    uint32_t  n_conceal;       /* H264B200_MB_CONCEAL macroblocks; their addresses, in concealment order, are uint32 values ... */
    uint32_t  conceal_offset;  /* ... starting at this coefficient slot */
    int       cur_slot;
    uint8_t   ref_slots_used[H264_MAX_SLOTS];
    void     *priv;
    /* device-parse path (kernel Kp): instead of records and slots the host assembles ONE block per picture,
     * h264b200_pichdr_t + per slice {h264b200_slice_t, RBSP, slice group map} (include/h264b200_slices.h) */
    uint8_t  *block;           /* NULL: this picture is parsed on the host */
    uint32_t  block_cap, block_used;
    uint32_t  has_p_slice;     /* a P slice was queued (inter prediction may be needed) */
} h264_pic_input_t;

typedef struct h264_backend h264_backend_t;
struct h264_backend {
    /* (re)allocate the frame pool of one decoder instance: n_slots frames of width x height MBs */
    void *(*inst_create)(h264_backend_t *be, uint32_t width_mbs, uint32_t height_mbs, uint32_t n_slots);
    void  (*inst_destroy)(h264_backend_t *be, void *inst);
    /* get a free input buffer for the next picture (may wait for an in-flight one) */
    h264_pic_input_t *(*pic_begin)(h264_backend_t *be, void *inst);
    /* grow pic->coef to at least min_slots (contents preserved); 0 on success */
    int   (*coef_grow)(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_slots);
    /* picture complete: reconstruct + deblock into frame slot pic->cur_slot (asynchronous) */
    int   (*pic_submit)(h264_backend_t *be, void *inst, h264_pic_input_t *pic);
    /* host-visible I420 frame of `slot`, valid once the picture last submitted into it is done */
    uint8_t *(*frame_host)(h264_backend_t *be, void *inst, int slot, uint32_t *error_flags);
    void  (*destroy)(h264_backend_t *be);
    void  *ctx;
    /* optional: address frame_host will return for `slot`, without launching or waiting */
    uint8_t *(*frame_host_async)(h264_backend_t *be, void *inst, int slot, uint32_t *gen);
    /* optional: wait for generation `gen` (from frame_host_async) of `slot`; 0 ok, 1 overwritten by a later picture, -1 error */
    int (*frame_wait)(h264_backend_t *be, void *inst, int slot, uint32_t gen, uint32_t *error_flags);
    /* optional: output format of an instance (H264B200_OUT_*) and the cropping rectangle in luma samples */
    int (*set_output)(h264_backend_t *be, void *inst, int format, int crop_left, int crop_top, int crop_width, int crop_height);
    /* optional, device-parse backends: make room for at least min_bytes in pic->block (contents preserved); 0 on success.
     * A backend that provides it hands out pictures with pic->block != NULL when the instance was created in
     * device-parse mode (see parse_mode below). */
    int (*block_grow)(h264_backend_t *be, void *inst, h264_pic_input_t *pic, uint32_t min_bytes);
    /* optional: status words of the picture last reconstructed into `slot` (valid after frame_host / frame_wait) */
    int (*frame_status)(h264_backend_t *be, void *inst, int slot, h264b200_picstat_t *out);
    /* optional: the caller is done with generation `gen` of `slot` (frame_host_async): its host mirror may be overwritten */
    void (*frame_release)(h264_backend_t *be, void *inst, int slot, uint32_t gen);
    /* optional: pictures of the instance submitted but not yet launched */
    uint32_t (*inst_pending)(h264_backend_t *be, void *inst);
    /* 1: instances are to be created in device-parse mode (h264b200_slices.h); read by the decoder at activation */
    int parse_mode;
    /* optional, never blocks: 0 generation `gen` of `slot` is in its host mirror, 1 launched but still in flight, 2 not launched yet, -1 error */
    int (*frame_state)(h264_backend_t *be, void *inst, int slot, uint32_t gen);
    /* optional: inst_create with the parse mode chosen per instance (host_parse != 0: records come from the host parser
     * even though parse_mode is set) — h264b200SetHostParse, the host / device split of h264b200DecodeStreams */
    void *(*inst_create_ex)(h264_backend_t *be, uint32_t width_mbs, uint32_t height_mbs, uint32_t n_slots, int host_parse);
};

/* implemented by whichever backend is linked: the CUDA engine in libh264b200.so */
h264_backend_t *h264_default_backend(void);

/* ------------------------------------------------------- per-MB host context */
typedef struct {
    uint8_t  tc[24];           /* TotalCoeff by luma4x4BlkIdx, Cb 16..19, Cr 20..23 */
    int8_t   ref_idx[4];       /* refIdxL0 per 8x8 (-1 intra) */
    uint8_t  kind;             /* H264B200_MB_* */
    uint8_t  decoded;
    uint16_t slice_id;         /* 0 = not decoded in this picture */
} h264_mbctx_t;

/* ------------------------------------------------------- decoder instance */
typedef struct h264_decoder {
    h264_backend_t *be;
    void *be_inst;
    int   owns_backend;
    h264_sps_t *sps[H264_MAX_SPS];
    h264_pps_t *pps[H264_MAX_PPS];
    int active_sps_id, active_pps_id, old_sps_id;
    h264_sps_t *active_sps; h264_pps_t *active_pps;
    int pending_activation;
    uint32_t width_mbs, height_mbs, pic_size_mbs, n_slots;
    int no_reordering_app;

    /* NAL / access unit state (h264bsd_storage.c:632-800) */
    struct { int first_call; uint8_t prev_ref_idc, prev_type; uint32_t prev_frame_num, prev_idr_pic_id, prev_poc_lsb;
             int32_t prev_delta_poc_bottom, prev_delta_poc[2]; } aub;
    int pic_started, valid_slice_in_au, skip_redundant;
    uint8_t *prev_buf_ptr; uint32_t prev_bytes_consumed; int prev_buf_not_finished;
    const uint8_t *nal_data; size_t nal_len;      /* current RBSP (inside the caller's buffer, or in nal_scratch) */
    int force_host_parse;                         /* h264b200SetHostParse: this instance parses slice data on the host whatever the engine's mode */
    int ro_input;                                 /* h264b200SetReadOnlyInput: never write to the caller's buffer */
    struct { uint8_t *p; uint32_t cap; } nal_scratch;   /* read-only input: a NAL with emulation prevention bytes is unescaped here */
    uint8_t cur_nal_type, cur_nal_ref_idc;
    uint8_t pic_nal_type, pic_nal_ref_idc;        /* of the last valid slice */

    h264_slice_hdr_t sh;                          /* last valid slice header */
    uint32_t slice_id;                            /* restarts at 1 each picture */
    uint32_t slice_last_mb;                       /* I slices: address of the last macroblock parsed without error (0: none) */
    uint32_t num_decoded_mbs, num_err_mbs;
    uint32_t current_pic_id;

    /* POC state (h264bsd_pic_order_cnt.c) */
    struct { uint32_t prev_poc_lsb; int32_t prev_poc_msb; uint32_t prev_frame_num, prev_frame_num_offset; int contains_mmco5; } poc;

    h264_dpb_t dpb;
    h264_mbctx_t *mbctx;                          /* pic_size_mbs */
    uint8_t *slice_group_map;                     /* pic_size_mbs, NULL without FMO (one slice group) */
    h264_pic_input_t *pic;                        /* input buffer of the picture being parsed */
    int last_output_slot;
    int out_format;                               /* H264B200_OUT_* requested through h264b200SetOutputFormat */
    int device_parse;                             /* slice data is parsed by kernel Kp: slices are queued, pictures end at the
                                                     next access unit boundary (or h264bsdFlushBuffer) */
} h264_decoder_t;

/* parameter sets / headers (h264_params.c) */
int h264_parse_sps(br_t *b, h264_sps_t *sps);
int h264_parse_pps(br_t *b, h264_pps_t *pps);
int h264_peek_pps_id(br_t b, uint32_t *pps_id);   /* by value: does not consume */
int h264_parse_slice_header(br_t *b, h264_slice_hdr_t *sh, const h264_sps_t *sps, const h264_pps_t *pps, int nal_type, int nal_ref_idc);
int32_t h264_decode_poc(h264_decoder_t *d, const h264_slice_hdr_t *sh, int nal_type, int nal_ref_idc);

/* CAVLC (h264_cavlc.c) */
void h264_cavlc_init(void);
#include "kp_types.h"
void h264_kp_fill_tables(KpTables *t);
/* the block decoder itself is inline: h264_cavlc_inl.h */

/* slice data (h264_slice.c): parse all macroblocks of one slice into d->pic */
int h264_decode_slice_data(h264_decoder_t *d, br_t *b, const h264_slice_hdr_t *sh);

#ifdef __cplusplus
}
#endif
#endif
