/* h264b200_records.h — the host -> device wire format of one decoded picture.
 *
 * The reference interleaves parsing and pixel reconstruction per macroblock
 * (h264bsd_slice_data.c:176-188): h264bsdDecodeMacroblockLayer fills a 2088-byte
 * macroblockLayer_t (h264bsd_macroblock_layer.h:141-161) and
 * h264bsdDecodeMacroblock (h264bsd_macroblock_layer.c:964-1134) consumes it at
 * once, leaving a 256-byte mbStorage_t (:166-188) behind for deblocking.  Here
 * the split point is after the QP update (:1043-1049): the host C parser emits,
 * per picture,
 *
 *   - one 128-byte `h264b200_mb_t` per macroblock, raster (MB address) order:
 *     everything the reconstruction and deblocking kernels need that is known
 *     from syntax alone (types, final intra modes and neighbour availability,
 *     final motion vectors and reference slots, QPs, deblocking controls), and
 *   - a packed array of 32-byte coefficient slots (16 x int16), only for 4x4
 *     blocks that carry residual, holding dequant-READY levels already moved from
 *     zig-zag scan order to raster order (the reference un-zig-zags inside
 *     h264bsdProcessBlock, h264bsd_transform.c:118-153).
 *
 * Kernel K1 transforms the slots in place (levels -> residual); K2 (inter) and
 * K3 (intra) read them; K4 (deblocking) reads only the 128-byte records.
 * Both structures are written once into pinned host memory and copied with one
 * cudaMemcpyAsync each.
 */
#ifndef H264B200_RECORDS_H
#define H264B200_RECORDS_H
#include <stdint.h>

/* mb_class */
#define H264B200_MB_INTER   0   /* any P macroblock incl. P_Skip */
#define H264B200_MB_I4x4    1
#define H264B200_MB_I16x16  2
#define H264B200_MB_IPCM    3
#define H264B200_MB_CONCEAL 4   /* lost macroblock filled by spatial interpolation from its neighbours (h264bsd_conceal.c:330-631);
                                  `avail` then holds H264B200_CN_*: the neighbours usable at its turn of the concealment order */
#define H264B200_MB_MISSING 255 /* never decoded and not concealable (no room for the concealment list): the samples of the frame slot stay as they are */

/* avail bits: neighbour usable for intra sample prediction (position, slice,
 * constrained_intra_pred all resolved on the host; h264bsd_neighbour.c:369-381,
 * h264bsd_intra_prediction.c:727-766) */
#define H264B200_AVAIL_A 1   /* left */
#define H264B200_AVAIL_B 2   /* up */
#define H264B200_AVAIL_C 4   /* up-right */
#define H264B200_AVAIL_D 8   /* up-left */

/* neighbours of a H264B200_MB_CONCEAL macroblock (decoded or already concealed) */
#define H264B200_CN_ABOVE 1
#define H264B200_CN_BELOW 2
#define H264B200_CN_LEFT  4
#define H264B200_CN_RIGHT 8

/* dbk_flags (h264bsd_deblocking.c:288-319 GetMbFilteringFlags, resolved on the host) */
#define H264B200_DBK_INNER 1 /* filter internal edges  (disable_deblocking_filter_idc != 1) */
#define H264B200_DBK_LEFT  2 /* filter left macroblock edge */
#define H264B200_DBK_TOP   4 /* filter top macroblock edge */

/* flags */
#define H264B200_MBF_DBK_AS_INTRA 1  /* concealed macroblock: the deblocking filter treats it as intra (h264bsd_conceal.c:300-306) */

/* resid_mask: bit b (0..15) luma4x4BlkIdx b has a slot; bits 16..19 Cb, 20..23 Cr;
 * bit 24: an Intra16x16 luma DC slot precedes the luma slots; bit 25: a chroma DC
 * slot (Cb dc[0..3], Cr dc[4..7]) precedes the chroma slots.  Slot order inside a
 * macroblock: [luma DC] [luma blocks, ascending blkIdx] [chroma DC] [Cb, Cr blocks].
 * I_PCM: resid_mask = 0 and 12 slots hold the 384 raw samples (Y 256, Cb 64, Cr 64). */
#define H264B200_RESID_LUMA_DC   (1u << 24)
#define H264B200_RESID_CHROMA_DC (1u << 25)

typedef struct {
    uint8_t  mb_class;        /* H264B200_MB_* */
    uint8_t  qp_y;            /* QP'Y used for dequantisation (running QP for P_Skip) */
    uint8_t  qp_c;            /* QPc used for chroma dequantisation */
    uint8_t  qp_dbk;          /* QPY seen by the deblocking filter (0 for I_PCM, h264bsd_macroblock_layer.c:1003) */
    int8_t   chroma_qp_off;   /* chroma_qp_index_offset of the macroblock's slice (h264bsd_deblocking.c:1489-1515) */
    uint8_t  dbk_flags;       /* H264B200_DBK_* */
    int8_t   dbk_off_a;       /* FilterOffsetA = 2*slice_alpha_c0_offset_div2 */
    int8_t   dbk_off_b;       /* FilterOffsetB */
    uint8_t  avail;           /* H264B200_AVAIL_* */
    uint8_t  i16_mode;        /* Intra16x16PredMode 0..3 */
    uint8_t  chroma_mode;     /* intra_chroma_pred_mode 0..3 */
    uint8_t  part_flags;      /* by partition syntax: bit q (0..3): 8x8 quadrant q is one partition; bit 4: the macroblock is one
                                 16x16 partition (informational: the kernels take the per-4x4 vectors) */
    uint32_t coef_offset;     /* first slot of this macroblock in the picture's slot array */
    uint32_t resid_mask;      /* see above */
    uint16_t nz_mask;         /* bit b: luma4x4BlkIdx b has TotalCoeff != 0 (bS=2 test; I_PCM: 0xffff) */
    uint16_t slice_id;        /* diagnostic only */
    uint8_t  ref_slot[4];     /* frame-pool slot of the reference picture per 8x8 quadrant (buffer identity for bS) */
    uint8_t  i4_mode[16];     /* Intra4x4PredMode by luma4x4BlkIdx */
    uint8_t  dbk_idc;         /* disable_deblocking_filter_idc of the macroblock's slice (diagnostic) */
    uint8_t  flags;           /* H264B200_MBF_* */
    uint8_t  reserved[18];
    int16_t  mv[16][2];       /* final motion vectors, quarter pel, by RASTER 4x4 position (by*4+bx): {hor, ver} */
} h264b200_mb_t;

#if defined(__cplusplus)
static_assert(sizeof(h264b200_mb_t) == 128, "h264b200_mb_t must be 128 bytes");
#else
_Static_assert(sizeof(h264b200_mb_t) == 128, "h264b200_mb_t must be 128 bytes");
#endif

#define H264B200_SLOT_I16 16   /* int16 values per coefficient slot */

#endif
