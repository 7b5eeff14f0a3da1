/* kp_types.h — plain-C data structures of the device-parse path (kernel Kp, kp_core.h / kp_parse.cuh): the look-up
 * table block the host builds once (h264_cavlc.c h264_kp_fill_tables), the per-macroblock parse context, the per-warp
 * staging area and the per-picture descriptors.  Shared by the host C sources, the CUDA engine and the CPU test build. */
#ifndef B200_KP_TYPES_H
#define B200_KP_TYPES_H
#include <stdint.h>
#include "h264b200_records.h"
#include "h264b200_slices.h"

/* the host parser's look-up tables (h264_cavlc.c) + the small constant tables of h264_consts.h, as one block */
typedef struct {
    uint8_t ct[3][128][4];       /* coeff_token, nC classes 0..2: [lz*8 + 3 bits] -> {len, TotalCoeff, TrailingOnes, -} */
    uint8_t ct_cdc[256][4];      /* chroma DC coeff_token by the next 8 bits */
    uint8_t tz[15][512][2];      /* total_zeros: [TotalCoeff-1][9 bits] -> {len, value} */
    uint8_t tz_cdc[3][8][2];
    uint8_t rb[7][8][2];         /* run_before: [min(zerosLeft,7)-1][3 bits] -> {len, run} (len 0: longer code) */
    int8_t  lvl[7][256][4];      /* level: [suffixLength][8 bits] -> {level, bits, next suffixLength, -} (bits 0: longer) */
    uint8_t cbp_map[48][2];      /* Table 9-4: codeNum -> coded_block_pattern {Intra4x4, Inter} */
    uint8_t zigzag[16];
    uint8_t raster_to_blk[16];
    uint8_t qpc[52];
    uint8_t lc_idx[16];          /* luma4x4BlkIdx -> index of the block in the luma TotalCoeff grid (KpStage.lc): 9 + x4 + 8 y4 */
    uint8_t ident4[4];           /* "scan" of a chroma DC block */
    uint8_t pad[8];
    uint16_t cbp_luma[16];       /* coded_block_pattern & 15 -> the luma4x4BlkIdx set of its 8x8 quadrants */
    uint32_t step_desc[28][2];   /* per residual step (kp_core.h kp_parse_residual): [0] grid cell | up distance << 8 | kind << 12 | mask bit << 16;
                                    [1] byte offset of the scan table in this struct | max coefficients << 16 | offset into the slot << 24 */
} KpTables;

typedef struct {                 /* == h264_mbctx_t (h264_internal.h) */
    uint8_t  tc[24];
    int8_t   ref_idx[4];
    uint8_t  kind, decoded;
    uint16_t slice_id;
} KpMbCtx;

/* per-warp staging (shared memory on the device) */
typedef struct {
    h264b200_mb_t rec;           /* the record being built */
    h264b200_mb_t nrec[4];       /* neighbour records A (left), B (up), C (up-right), D (up-left) */
    KpMbCtx ctx;                 /* the context being built */
    KpMbCtx nctx[4];
    KpMbCtx old;                 /* what ctx[addr] held before (a macroblock decoded twice is an error) */
    int16_t slots[27 * 16];      /* coefficient slots of this macroblock; ALL ZERO between macroblocks */
    int16_t lvl[16];             /* levels of the block being decoded, in decoding order */
    uint32_t mvd[16];            /* the vector differences {hor, ver} of the partitions (low 16 bits each: only those reach the int16 vector) */
    uint32_t mvg[60];            /* neighbour grid of motion vector prediction, 5 rows of 12 (kp_core.h KP_G); 16-byte aligned */
    int8_t   refg[64];           /* ... and its reference indices: -2 not available, -1 not inter */
    uint16_t part[16];           /* partition list of the macroblock (KP_PART) */
    uint8_t  grid[64];           /* TotalCoeff grids with guard row / column (nC): luma 5 rows of 8 (cells 0..39), then per chroma plane 3 rows of 4 */
    uint8_t  lvl_dummy;
    uint8_t  pad[15];
    uint16_t gsrc[64];           /* per cell of the TotalCoeff grid / of the vector grid: which neighbour value fills it (kp_core.h kp_stage_plan) */
    uint16_t msrc[64];
} KpStage;

typedef struct {                 /* outputs of one picture besides records / slots / contexts */
    uint32_t coef_used, n_intra, n_inter, any_deblock, n_conceal, conceal_offset;
    uint32_t pad[2];
    h264b200_picstat_t stat;
} KpResult;

typedef struct {                 /* one picture: where its input block and its outputs live */
    const uint8_t *block;        /* h264b200_pichdr_t + slices (include/h264b200_slices.h) */
    h264b200_mb_t *mbs;          /* width_mbs*height_mbs records */
    int16_t *coef;               /* coefficient slots, capacity coef_cap (worst case: 27 per macroblock + the concealment list) */
    KpMbCtx *ctx;                /* width_mbs*height_mbs contexts (scratch) */
    uint32_t coef_cap;
    uint32_t pad;
    KpResult *res;
} KpPic;

#define KP_COEF_CAP(n_mbs) ((n_mbs) * 27u + (n_mbs) / 8u + 8u)

#endif
