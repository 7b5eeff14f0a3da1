/* oracle/null_backend.c — TEST/MEASUREMENT INFRASTRUCTURE ONLY.
 * A backend that accepts pictures and discards them, so the host-side cost of the
 * product decoder (NAL scan + CAVLC parse + MV/intra-mode prediction + record
 * emission + DPB) can be timed in isolation: the Amdahl term of SURVEY.md §7.3(1). */
#include <stdlib.h>
#include <string.h>
#include "../broadway_b200/csrc/h264_internal.h"
typedef struct { h264_pic_input_t pic; uint8_t *frame; } null_inst_t;
static void *n_create(h264_backend_t *be, uint32_t wm, uint32_t hm, uint32_t n) {
    null_inst_t *in = calloc(1, sizeof *in); (void)be; (void)n;
    in->pic.mbs = calloc((size_t)wm * hm, sizeof(h264b200_mb_t));
    in->pic.coef_cap = wm * hm * 27 + 64; in->pic.coef = malloc((size_t)in->pic.coef_cap * 32);
    in->frame = calloc((size_t)wm * hm, 384); return in; }
static void n_destroy(h264_backend_t *be, void *i) { null_inst_t *in = i; (void)be; free(in->pic.mbs); free(in->pic.coef); free(in->frame); free(in); }
static h264_pic_input_t *n_begin(h264_backend_t *be, void *i) { (void)be; return &((null_inst_t *)i)->pic; }
static int n_grow(h264_backend_t *be, void *i, h264_pic_input_t *p, uint32_t m) { (void)be; (void)i; (void)p; (void)m; return -1; }
static int n_submit(h264_backend_t *be, void *i, h264_pic_input_t *p) { (void)be; (void)i; (void)p; return 0; }
static uint8_t *n_frame(h264_backend_t *be, void *i, int s, uint32_t *e) { (void)be; (void)s; if (e) *e = 0; return ((null_inst_t *)i)->frame; }
static void n_del(h264_backend_t *be) { (void)be; }
#ifdef NULL_DEVICE_PARSE
/* device-parse flavour: the host only scans NAL units, parses slice headers and copies the slice payloads into the block */
static int n_block_grow(h264_backend_t *be, void *i, h264_pic_input_t *p, uint32_t m) { (void)be; (void)i; uint8_t *n = realloc(p->block, m + 4096); if (!n) return -1; p->block = n; p->block_cap = m + 4096; return 0; }
static h264_backend_t g = { n_create, n_destroy, n_begin, n_grow, n_submit, n_frame, n_del, NULL, NULL, NULL, NULL, n_block_grow, NULL, NULL, NULL, 1 };
#else
static h264_backend_t g = { n_create, n_destroy, n_begin, n_grow, n_submit, n_frame, n_del, NULL };
#endif
h264_backend_t *h264_default_backend(void) { return &g; }
