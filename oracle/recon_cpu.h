/* oracle/recon_cpu.h — TEST INFRASTRUCTURE ONLY (see recon_cpu.c). */
#ifndef ORACLE_RECON_CPU_H
#define ORACLE_RECON_CPU_H
#include <stdint.h>
#include <stddef.h>
#include "../broadway_b200/csrc/h264_internal.h"

/* K1: dequant + inverse transforms of one macroblock's coefficient slots, in place.
 * Returns 1 if a residual sample left [-512,511]. */
int  recon_cpu_residual_mb(const h264b200_mb_t *mb, int16_t *coef);
/* K2+K3: inter and intra prediction + residual add of a whole picture (slots already transformed). */
void recon_cpu_predict_picture(const h264b200_mb_t *mbs, const int16_t *coef, int wm, int hm, uint8_t *cur_frame, uint8_t *const *frames);
/* K3c: spatial concealment of the H264B200_MB_CONCEAL macroblocks, in the order of `list`. */
void recon_cpu_conceal_picture(const h264b200_mb_t *mbs, const uint32_t *list, uint32_t n, int wm, int hm, uint8_t *frame);
/* K4: in-loop deblocking of a whole picture, in place. */
void recon_cpu_deblock_picture(const h264b200_mb_t *mbs, int wm, int hm, uint8_t *frame);
/* single fractional samples (unit tests of the interpolators) */
int  recon_cpu_luma_sample(const uint8_t *plane, int w, int h, int x, int y, int fx, int fy);
int  recon_cpu_chroma_sample(const uint8_t *plane, int w, int h, int x, int y, int fx, int fy);

/* observation callbacks fired by the CPU backend for every picture */
typedef struct {
    void *user;
    void (*records)(void *user, const h264b200_mb_t *mbs, uint32_t n_mbs, const int16_t *coef, uint32_t n_slots);
    void (*residual)(void *user, const int16_t *coef, uint32_t n_slots);
    void (*predeblock)(void *user, const uint8_t *frame, size_t bytes);
} recon_cpu_tap_t;
void recon_cpu_set_tap(const recon_cpu_tap_t *t);

#endif
