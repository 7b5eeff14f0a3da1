/* oracle/ref_tap.c — TEST INFRASTRUCTURE ONLY.
 *
 * Observation hooks linked into oracle/_ref/libh264ref.so next to the
 * unmodified reference objects.  `-Wl,--wrap=h264bsdFilterPicture` makes the
 * reference's call in h264bsd_decoder.c:489-491 land here first, so tests can
 * capture (a) the picture BEFORE in-loop deblocking and (b) the per-macroblock
 * state (mbStorage_t, h264bsd_macroblock_layer.h:166-188) the reference derived
 * for that picture.  That lets each kernel family (K1-K3 vs K4) and the host
 * parser be parity-checked in isolation.  Nothing here alters decode results.
 */
#include <string.h>
#include "h264bsd_image.h"
#include "h264bsd_macroblock_layer.h"
#include "ref_tap.h"

void __real_h264bsdFilterPicture(image_t *image, mbStorage_t *mb);

static unsigned char *g_pre_buf;
static size_t g_pre_cap;
static size_t g_pre_len;
static reftap_mb_t *g_mb_buf;
static size_t g_mb_cap;
static size_t g_mb_len;
static unsigned g_pics;

void reftap_set_predeblock_buffer(unsigned char *buf, size_t cap)
{ g_pre_buf = buf; g_pre_cap = cap; g_pre_len = 0; }

void reftap_set_mb_buffer(reftap_mb_t *buf, size_t cap_mbs)
{ g_mb_buf = buf; g_mb_cap = cap_mbs; g_mb_len = 0; }

size_t reftap_predeblock_len(void) { return g_pre_len; }
size_t reftap_mb_len(void) { return g_mb_len; }
unsigned reftap_pictures_filtered(void) { return g_pics; }

void __wrap_h264bsdFilterPicture(image_t *image, mbStorage_t *mb)
{
    size_t nmb = (size_t)image->width * image->height;
    size_t n = nmb * 384;
    g_pics++;
    if (g_pre_buf && n <= g_pre_cap) {
        memcpy(g_pre_buf, image->data, n);
        g_pre_len = n;
    }
    if (g_mb_buf && nmb <= g_mb_cap) {
        size_t i; unsigned k;
        for (i = 0; i < nmb; i++) {
            reftap_mb_t *o = &g_mb_buf[i];
            o->mb_type = (int)mb[i].mbType;
            o->slice_id = mb[i].sliceId;
            o->qp_y = mb[i].qpY;
            o->disable_deblock_idc = mb[i].disableDeblockingFilterIdc;
            o->filter_offset_a = mb[i].filterOffsetA;
            o->filter_offset_b = mb[i].filterOffsetB;
            o->chroma_qp_index_offset = mb[i].chromaQpIndexOffset;
            for (k = 0; k < 27; k++) o->total_coeff[k] = (short)mb[i].totalCoeff[k];
            for (k = 0; k < 16; k++) o->intra4x4_mode[k] = mb[i].intra4x4PredMode[k];
            for (k = 0; k < 16; k++) { o->mv[k][0] = mb[i].mv[k].hor; o->mv[k][1] = mb[i].mv[k].ver; }
            for (k = 0; k < 4; k++) o->ref_pic[k] = mb[i].refPic[k];
        }
        g_mb_len = nmb;
    }
    __real_h264bsdFilterPicture(image, mb);
}
