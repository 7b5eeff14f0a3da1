/* h264b200_slices.h — host -> device wire format of the DEVICE-PARSE path (kernel Kp).
 *
 * In the host-parse path (include/h264b200_records.h) the host runs the serial
 * slice_data()/macroblock_layer()/CAVLC parse of the reference
 * (h264bsd_slice_data.c:85-235, h264bsd_macroblock_layer.c:133-242, :353-496,
 * :699-869, h264bsd_cavlc.c:395-915, h264bsd_inter_prediction.c:499-1031) and
 * ships 128-byte macroblock records + coefficient slots: 2.8 MB per 1080p
 * picture over PCIe and ~4 ms of one host core.  In the device-parse path the
 * host keeps only what is cheap and stateful — NAL extraction, parameter sets,
 * the slice HEADER, picture order count, the DPB with its frame-pool slot
 * assignment, access-unit boundaries — and ships, per picture, one block:
 *
 *     [h264b200_pichdr_t 64 B]
 *     [h264b200_slice_t 64 B][RBSP of the slice NAL, zero padded to 16 B][slice group map, padded]   x n_slices
 *
 * Kernel Kp (csrc/kp_core.h, kp_parse.cuh) parses the slice data of MANY
 * pictures at once — one warp per picture, pictures of one stream are mutually
 * independent at this level — and writes the very same h264b200_mb_t records and
 * coefficient slots into HBM that the host parser would have produced, so the
 * reconstruction kernels K1..K4 are untouched.  Lost or damaged slices are
 * handled on the device the way the host path does it (the macroblocks of a
 * failed slice are given back, h264bsd_slice_data.c:302-358, and concealed when
 * the picture ends, h264bsd_conceal.c:125-255); the number of concealed
 * macroblocks travels back with the frame (h264b200_picstat_t).
 */
#ifndef H264B200_SLICES_H
#define H264B200_SLICES_H
#include <stdint.h>

#define H264B200_PICHDR_MAGIC 0x48423250u   /* "P2BH" */

typedef struct {
    uint32_t magic;
    uint32_t n_slices;
    uint32_t total_bytes;        /* of the whole block, multiple of 16 */
    uint16_t width_mbs, height_mbs;
    int8_t   conceal_ref_slot;   /* frame-pool slot of the first usable reference picture when the picture ends, -1: none
                                    (h264bsd_conceal.c:150-170) */
    uint8_t  conceal_as_p;       /* lost macroblocks are concealed as in a P picture (last valid slice was P, or no valid slice) */
    uint8_t  tentative;          /* the picture was ended by h264bsdFlushBuffer, not by a following access unit: if its slices do not
                                    cover it, it is NOT a picture (the host parser would never have reported it ready) and must be dropped */
    uint8_t  pad0;
    uint32_t reserved[11];
} h264b200_pichdr_t;

typedef struct {
    uint32_t size;               /* bytes from this descriptor to the next one, multiple of 16 */
    uint32_t rbsp_len;           /* RBSP bytes (emulation prevention removed, NAL header byte excluded) behind the descriptor */
    uint32_t bit_off;            /* slice_data() starts at this bit of the RBSP: the header was parsed on the host */
    uint32_t payload_bits;       /* RBSP bits before rbsp_stop_one_bit */
    uint32_t first_mb;
    uint32_t map_off;            /* offset from the descriptor of the mbs-long slice group map valid for this slice; 0: one slice group */
    uint16_t slice_id;           /* 1.. within the picture */
    uint8_t  is_p;
    uint8_t  num_ref_idx_active;
    int8_t   slice_qp, chroma_qp_off, alpha_off, beta_off;   /* offsets already doubled */
    uint8_t  disable_deblocking_idc, constrained_intra;
    uint8_t  pad0[2];
    int8_t   ref_slot[17];       /* frame-pool slot per refIdxL0 after list reordering; -1: missing / non-existing picture */
    uint8_t  pad1[3];
    uint32_t reserved[2];
} h264b200_slice_t;

/* travels back with every frame of a device-parsed picture (behind the I420 samples) */
typedef struct {
    uint32_t err_mbs;            /* macroblocks concealed or missing (numErrMbs of h264bsdNextOutputPicture) */
    uint32_t flags;              /* H264B200_PS_* */
    uint32_t decoded_mbs;        /* macroblocks the slices delivered */
    uint32_t coef_slots;         /* coefficient slots written */
} h264b200_picstat_t;
#define H264B200_PS_SLICE_ERROR 1u   /* at least one slice ended with a syntax error */
#define H264B200_PS_INCOMPLETE  2u   /* the slices did not cover the picture */
#define H264B200_PS_DROPPED     4u   /* incomplete last picture of a stream: not to be output (see h264b200_pichdr_t.tentative) */

#if defined(__cplusplus)
static_assert(sizeof(h264b200_pichdr_t) == 64 && sizeof(h264b200_slice_t) == 64 && sizeof(h264b200_picstat_t) == 16, "wire format");
#else
_Static_assert(sizeof(h264b200_pichdr_t) == 64 && sizeof(h264b200_slice_t) == 64 && sizeof(h264b200_picstat_t) == 16, "wire format");
#endif

#endif
