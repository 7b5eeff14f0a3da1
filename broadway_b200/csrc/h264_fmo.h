/* h264_fmo.h — macroblock to slice group map (flexible macroblock ordering, Baseline profile;
 * ITU-T H.264 8.2.2.1 - 8.2.2.7, frame_mbs_only_flag = 1 so map units are macroblocks).
 * Shared by the host parser and the synthetic bitstream writer.  The reference's counterpart is
 * h264bsdDecodeSliceGroupMap (Decoder/src/h264bsd_slice_group_map.c:503-589) with its six
 * per-type helpers (:120-500); this is written from the standard's pseudo-code. */
#ifndef B200_H264_FMO_H
#define B200_H264_FMO_H
#include <stdint.h>
#include <string.h>

typedef struct {
    uint32_t n_groups;                 /* 2..8 */
    uint32_t type;                     /* slice_group_map_type 0..6 */
    uint32_t run_length[8];            /* type 0: run_length_minus1 + 1 */
    uint32_t top_left[8], bottom_right[8];   /* type 2, macroblock addresses */
    uint32_t change_direction;         /* types 3..5 */
    uint32_t change_rate;              /* types 3..5: slice_group_change_rate_minus1 + 1 */
    const uint8_t *group_id;           /* type 6: one id per macroblock */
} h264_fmo_t;

/* units0 = Min(slice_group_change_cycle * SliceGroupChangeRate, PicSizeInMapUnits) (7-33), types 3..5 only */
static inline void h264_fmo_build_map(uint8_t *map, uint32_t W, uint32_t H, const h264_fmo_t *f, uint32_t units0)
{
    const uint32_t size = W * H, n = f->n_groups;
    uint32_t i, j, k, g;
    switch (f->type) {
    case 0:                                                    /* interleaved (8.2.2.1) */
        i = 0;
        do {
            for (g = 0; g < n && i < size; i += f->run_length[g++])
                for (j = 0; j < f->run_length[g] && i + j < size; j++) map[i + j] = (uint8_t)g;
        } while (i < size);
        break;
    case 1:                                                    /* dispersed (8.2.2.2) */
        for (i = 0; i < size; i++) map[i] = (uint8_t)(((i % W) + (((i / W) * n) / 2)) % n);
        break;
    case 2:                                                    /* foreground with left-over (8.2.2.3) */
        memset(map, (int)(n - 1), size);
        for (g = n - 1; g-- > 0;) {
            const uint32_t x0 = f->top_left[g] % W, y0 = f->top_left[g] / W, x1 = f->bottom_right[g] % W, y1 = f->bottom_right[g] / W;
            uint32_t x, y;
            for (y = y0; y <= y1 && y < H; y++) for (x = x0; x <= x1; x++) map[y * W + x] = (uint8_t)g;
        }
        break;
    case 3: {                                                  /* box-out (8.2.2.4) */
        const int dir = (int)f->change_direction;
        int x = ((int)W - dir) / 2, y = ((int)H - dir) / 2;
        int left = x, top = y, right = x, bottom = y, xd = dir - 1, yd = dir;
        uint32_t vacant;
        memset(map, 1, size);
        for (k = 0; k < units0; k += vacant) {
            vacant = map[(uint32_t)y * W + (uint32_t)x] == 1;
            if (vacant) map[(uint32_t)y * W + (uint32_t)x] = 0;
            if (xd == -1 && x == left) { left = left > 0 ? left - 1 : 0; x = left; xd = 0; yd = 2 * dir - 1; }
            else if (xd == 1 && x == right) { right = right + 1 < (int)W ? right + 1 : (int)W - 1; x = right; xd = 0; yd = 1 - 2 * dir; }
            else if (yd == -1 && y == top) { top = top > 0 ? top - 1 : 0; y = top; xd = 1 - 2 * dir; yd = 0; }
            else if (yd == 1 && y == bottom) { bottom = bottom + 1 < (int)H ? bottom + 1 : (int)H - 1; y = bottom; xd = 2 * dir - 1; yd = 0; }
            else { x += xd; y += yd; }
        }
        break; }
    case 4: {                                                  /* raster scan (8.2.2.5) */
        const uint32_t upper = f->change_direction ? size - units0 : units0;
        for (i = 0; i < size; i++) map[i] = (uint8_t)(i < upper ? f->change_direction : 1 - f->change_direction);
        break; }
    case 5: {                                                  /* wipe (8.2.2.6) */
        const uint32_t upper = f->change_direction ? size - units0 : units0;
        k = 0;
        for (j = 0; j < W; j++) for (i = 0; i < H; i++) map[i * W + j] = (uint8_t)(k++ < upper ? f->change_direction : 1 - f->change_direction);
        break; }
    default:                                                   /* explicit (8.2.2.7) */
        for (i = 0; i < size; i++) map[i] = f->group_id[i];
        break;
    }
}

/* bits of slice_group_change_cycle: Ceil(Log2(PicSizeInMapUnits / SliceGroupChangeRate + 1)) (7.4.3) */
static inline uint32_t h264_fmo_cycle_bits(uint32_t size, uint32_t rate)
{
    uint32_t v = size / rate + 1 + (size % rate ? 1 : 0), n = 0;
    while ((1u << n) < v) n++;
    return n;
}
#endif
