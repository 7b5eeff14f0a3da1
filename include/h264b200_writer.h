/* h264b200_writer.h — synthetic H.264 Baseline (CAVLC) Annex-B bitstream writer.
 *
 * BASELINE.json's north_star asks for "synthetic Baseline CAVLC bitstreams from
 * an in-repo bitstream writer (random residuals/MVs/intra modes at 1080p and
 * 4K)": the reference ships no encoder and the bundled Player mp4 clips are
 * absent from the mount, so this writer is the only source of test and bench
 * streams.  It is NOT an encoder (no analysis, no rate control): it draws
 * macroblock types, partitions, motion vectors, intra modes, CBPs and
 * coefficient levels from a seeded PRNG and serialises them with exactly the
 * syntax the reference parses (h264bsd_seq_param_set.c, _pic_param_set.c,
 * _slice_header.c, _slice_data.c, _macroblock_layer.c, _cavlc.c), tracking the
 * same neighbour state (nC, MV and intra-mode predictors) a decoder derives,
 * and keeping every stream inside the constraints under which the reference
 * decodes without concealment.
 */
#ifndef H264B200_WRITER_H
#define H264B200_WRITER_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t width_mbs;            /* picture width  in macroblocks (1080p: 120) */
    uint32_t height_mbs;           /* picture height in macroblocks (1080p: 68)  */
    uint32_t n_frames;             /* pictures to write */
    uint32_t idr_period;           /* IDR every n pictures; 0 = first picture only */
    uint32_t intra_only;           /* 1: every picture is an I picture */
    uint64_t seed;                 /* PRNG seed; same params => identical bytes */
    int32_t  qp;                   /* pic_init_qp / slice QP */
    int32_t  qp_jitter;            /* max |QP - qp| reached through mb_qp_delta (0 = constant QP) */
    uint32_t coded_blk_permille;   /* probability a 4x4 block carries coefficients */
    uint32_t max_coeffs;           /* 1..16 nonzero coefficients per coded block */
    int32_t  max_level;            /* |level| drawn from 1..max_level */
    uint32_t num_ref_frames;       /* 1..16 */
    uint32_t slices_per_pic;       /* >=1, macroblocks split evenly in raster order */
    uint32_t poc_type;             /* 0 or 2 */
    int32_t  chroma_qp_index_offset;
    uint32_t deblock_idc;          /* disable_deblocking_filter_idc 0/1/2 */
    int32_t  alpha_c0_offset_div2; /* -6..6 */
    int32_t  beta_offset_div2;     /* -6..6 */
    uint32_t constrained_intra_pred;
    uint32_t p_intra_permille;     /* intra macroblocks inside P pictures */
    uint32_t p_skip_permille;      /* P_Skip macroblocks */
    uint32_t ipcm_permille;        /* I_PCM among intra macroblocks */
    uint32_t i16_permille;         /* I16x16 (vs I4x4) among intra macroblocks */
    int32_t  mv_range_qpel;        /* final MVs uniform in +-range (quarter pels) */
    uint32_t far_mv_permille;      /* MVs pointing far outside the picture */
    uint32_t level_idc;            /* 40 for 1080p, 51 for 4K */
    uint32_t first_idr_ipcm;       /* 1: IDR pictures are all I_PCM noise (texture-rich reference) */
    uint32_t part_mix;             /* 0: 16x16 only; 1: all partition shapes */
    uint32_t crop;                 /* 1: signal frame cropping of 8 luma rows (1088 -> 1080) */
    uint32_t multi_slice_params;   /* 1: vary deblock idc/offsets and QP per slice */
    uint32_t dpb_stress;           /* 1: non-reference pictures, ref list reordering and MMCO 1 marking (needs num_ref_frames >= 2);
                                      2: also long-term pictures (MMCO 4+6, released by MMCO 2) */
    uint32_t fmo_type;             /* 0: one slice group; t+1: flexible macroblock ordering with slice_group_map_type t (0..6) */
    uint32_t fmo_groups;           /* slice groups 2..8 (map types 3..5 always use 2) */
} h264w_params_t;

/* Fill *p with the defaults used by BASELINE.json config 3 at the given size. */
void h264w_default_params(h264w_params_t *p, uint32_t width_mbs, uint32_t height_mbs, uint32_t n_frames);

/* Upper bound of the bytes h264w_generate may produce for *p. */
size_t h264w_bound(const h264w_params_t *p);

/* Write an Annex-B stream (SPS, PPS, then one access unit per picture) into
 * out[0..cap).  Returns the number of bytes written, or 0 on failure
 * (bad parameters or cap too small). */
size_t h264w_generate(const h264w_params_t *p, uint8_t *out, size_t cap);

#ifdef __cplusplus
}
#endif
#endif
