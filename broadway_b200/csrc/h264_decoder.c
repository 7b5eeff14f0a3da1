/* h264_decoder.c — the NAL-level state machine behind h264bsdInit / h264bsdDecode /
 * h264bsdNextOutputPicture / h264bsdShutdown and the accessors
 * (reference: h264bsd_decoder.c:99-122, :162-560, :579-618, :642-1006,
 * h264bsd_byte_stream.c:80-237, h264bsd_nal_unit.c:68-117,
 * h264bsd_storage.c:298-420 activation, :632-800 access-unit boundary).
 *
 * Same call protocol and return codes as the reference: one NAL unit consumed
 * per call; *readBytes = 0 when the same buffer must be presented again
 * (H264BSD_HDRS_RDY after a new SPS was activated); H264BSD_PIC_RDY as soon as
 * the last macroblock of a picture has been parsed.  What differs is what
 * "decode a slice" means: the slice is parsed into macroblock records and the
 * finished picture is handed to the backend (CUDA engine) for reconstruction;
 * pixels exist only on the GPU until h264bsdNextOutputPicture asks for them.
 * Error concealment (h264bsd_conceal.c) is out of scope: macroblocks of a lost
 * or corrupt slice are reported through numErrMbs and left unreconstructed.
 */
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include "h264b200.h"
#include "h264b200_batch.h"
#include "h264_internal.h"
#include "h264_consts.h"

static h264_decoder_t *DEC(storage_t *s) { return s ? (h264_decoder_t *)s->impl : NULL; }

/* ------------------------------------------------------------ NAL extraction */
#if defined(__SSE2__)
#include <emmintrin.h>
#endif
/* index of the first i < n - 1 with p[i] == 0 && p[i+1] == 0 — the only place a start code prefix or an emulation
 * prevention sequence can begin — or n when there is none.  16 bytes per step: the scan of a 1080p slice NAL
 * (~230 KB, a zero byte every 256 bytes in coefficient data) costs a few microseconds instead of one memchr call per
 * zero byte. */
static uint32_t find_zero_pair(const uint8_t *p, uint32_t n)
{
    uint32_t i = 0;
#if defined(__SSE2__)
    const __m128i z = _mm_setzero_si128();
    while (i + 17 <= n) {
        const __m128i a = _mm_loadu_si128((const __m128i *)(p + i)), b = _mm_loadu_si128((const __m128i *)(p + i + 1));
        const int m = _mm_movemask_epi8(_mm_and_si128(_mm_cmpeq_epi8(a, z), _mm_cmpeq_epi8(b, z)));
        if (m) return i + (uint32_t)__builtin_ctz((unsigned)m);
        i += 16;
    }
#endif
    while (i + 1 < n) {
        const uint8_t *q = (const uint8_t *)memchr(p + i, 0, n - 1 - i);
        if (!q) break;
        i = (uint32_t)(q - p);
        if (p[i + 1] == 0) return i;
        i++;
    }
    return n;
}

/* Copy src[0..n) to dst without its emulation prevention bytes (00 00 03 -> 00 00); returns the RBSP length or
 * (uint32_t)-1 on a forbidden sequence (h264bsd_byte_stream.c:192-234 reports the same streams as errors). */
static uint32_t unescape_copy(uint8_t *dst, const uint8_t *src, uint32_t n)
{
    uint32_t r = 0, w = 0;
    while (r < n) {
        uint32_t k = find_zero_pair(src + r, n - r);
        if (k >= n - r) { memmove(dst + w, src + r, n - r); w += n - r; break; }
        k += 2;                                        /* up to and including the two zero bytes */
        memmove(dst + w, src + r, k); w += k; r += k;
        if (r < n && src[r] == 0) return (uint32_t)-1;             /* 00 00 00 inside a NAL unit */
        if (r < n && src[r] == 3) {
            if (r + 1 == n || src[r + 1] > 3) return (uint32_t)-1;
            r++;                                       /* drop the emulation prevention byte */
        } else if (r < n && src[r] <= 2) return (uint32_t)-1;
    }
    return w;
}

/* Find the NAL unit at the head of buf: Annex-B (00 00 01 / 00 00 00 prefix) or a
 * bare NAL, and remove its emulation prevention bytes: IN PLACE (the reference does the
 * same to the caller's buffer, h264bsd_byte_stream.c:192-234), or — when `scratch` is given
 * (read-only input, h264b200SetReadOnlyInput) — by copying a NAL that contains any into
 * *scratch, so that the caller's bytes are never written.  Returns 0 and sets
 * nal/nal_len/consumed, or -1. */
typedef struct { uint8_t *p; uint32_t cap; } nal_scratch_t;
static int extract_nal(uint8_t *buf, uint32_t len, uint8_t **nal, uint32_t *nal_len, uint32_t *consumed, nal_scratch_t *scratch)
{
    uint32_t start = 0, end = len, trailing = 0, i, zeros;
    int has_epb = 0, invalid = 0;
    if (len > 3 && buf[0] == 0 && buf[1] == 0 && (buf[2] & 0xFE) == 0) {
        /* skip to the byte after the first start code prefix */
        i = 2; zeros = 2;
        for (;;) {
            uint8_t v = buf[i++];
            if (i == len) { *consumed = len; return -1; }
            if (!v) zeros++;
            else if (v == 1 && zeros >= 2) break;
            else zeros = 0;
        }
        start = i;
        /* the NAL ends at the next start code prefix or at the end of the buffer; trailing
         * zero bytes are not part of it */
        while (i < len) {
            uint32_t j, k = find_zero_pair(buf + i, len - i);
            if (k >= len - i) {                        /* no 00 00 any more: at most one trailing zero byte */
                if (buf[len - 1] == 0 && len - 1 >= i) { end = len - 1; trailing = 1; }
                break;
            }
            i += k;
            for (j = i; j < len && buf[j] == 0; j++) ;
            zeros = j - i;
            if (j == len) { end = i; trailing = zeros; break; }
            if (buf[j] == 1) { end = i; trailing = zeros > 3 ? zeros - 3 : 0; break; }
            if (zeros == 2 && buf[j] == 3) has_epb = 1;
            else if (zeros >= 3) invalid = 1;          /* 00 00 00 xx inside a NAL unit */
            i = j + 1;
        }
        *nal = buf + start; *nal_len = end - start; *consumed = end + trailing;
        if (invalid) return -1;
    } else {
        *nal = buf; *nal_len = len; *consumed = len; has_epb = 1;
    }
    if (has_epb) {
        uint8_t *dst = *nal;
        uint32_t n;
        if (scratch) {
            if (scratch->cap < *nal_len) {
                uint8_t *q = (uint8_t *)h264_malloc((size_t)*nal_len + 4096);
                if (!q) return -1;
                h264_free(scratch->p); scratch->p = q; scratch->cap = *nal_len + 4096;
            }
            dst = scratch->p;
        }
        n = unescape_copy(dst, *nal, *nal_len);        /* in place: the write position never passes the read position */
        if (n == (uint32_t)-1) return -1;
        *nal = dst; *nal_len = n;
    }
    return 0;
}

/* ------------------------------------------------------- access unit boundary */
static int check_au_boundary(h264_decoder_t *d, br_t b /* by value, after the NAL header */, int type, int ref_idc, int *boundary)
{
    uint32_t pps_id, v, frame_num;
    const h264_pps_t *pps; const h264_sps_t *sps;
    *boundary = 0;
    if ((type > 5 && type < 12) || (type > 12 && type <= 18)) { *boundary = 1; return 0; }
    if (type != NAL_SLICE && type != NAL_IDR) return 0;
    if (d->aub.first_call) { *boundary = 1; d->aub.first_call = 0; }
    if (h264_peek_pps_id(b, &pps_id)) return -1;
    pps = d->pps[pps_id];
    if (!pps || !d->sps[pps->sps_id] ||
        (d->active_sps_id != H264_MAX_SPS && pps->sps_id != d->active_sps_id && type != NAL_IDR)) return -2;
    sps = d->sps[pps->sps_id];
    if (d->aub.prev_ref_idc != ref_idc && (d->aub.prev_ref_idc == 0 || ref_idc == 0)) *boundary = 1;
    if ((d->aub.prev_type == NAL_IDR) != (type == NAL_IDR)) *boundary = 1;
    v = br_ue(&b); v = br_ue(&b); v = br_ue(&b);     /* first_mb, slice_type, pps_id */
    (void)v;
    frame_num = br_get(&b, sps->log2_max_frame_num);
    if (d->aub.prev_frame_num != frame_num) { d->aub.prev_frame_num = frame_num; *boundary = 1; }
    if (type == NAL_IDR) {
        v = br_ue(&b);
        if (v == 0xffffffffu) return -1;
        if (d->aub.prev_type == NAL_IDR && d->aub.prev_idr_pic_id != v) *boundary = 1;
        d->aub.prev_idr_pic_id = v;
    }
    if (sps->poc_type == 0) {
        v = br_get(&b, sps->log2_max_poc_lsb);
        if (d->aub.prev_poc_lsb != v) { d->aub.prev_poc_lsb = v; *boundary = 1; }
        if (pps->pic_order_present) {
            int32_t sv = br_se(&b);
            if (d->aub.prev_delta_poc_bottom != sv) { d->aub.prev_delta_poc_bottom = sv; *boundary = 1; }
        }
    } else if (sps->poc_type == 1 && !sps->delta_pic_order_always_zero) {
        int32_t sv = br_se(&b);
        if (d->aub.prev_delta_poc[0] != sv) { d->aub.prev_delta_poc[0] = sv; *boundary = 1; }
        if (pps->pic_order_present) {
            sv = br_se(&b);
            if (d->aub.prev_delta_poc[1] != sv) { d->aub.prev_delta_poc[1] = sv; *boundary = 1; }
        }
    }
    d->aub.prev_ref_idc = (uint8_t)ref_idc; d->aub.prev_type = (uint8_t)type;
    return br_overrun(&b) ? -1 : 0;
}

static int apply_output_format(h264_decoder_t *d);

/* --------------------------------------------------------- parameter sets */
static int sps_equal(const h264_sps_t *a, const h264_sps_t *b) { return memcmp(a, b, sizeof *a) == 0; }

static int store_sps(h264_decoder_t *d, const h264_sps_t *s)
{
    int id = s->sps_id;
    if (!d->sps[id]) { d->sps[id] = (h264_sps_t *)h264_malloc(sizeof *s); if (!d->sps[id]) return -1; }
    else if (id == d->active_sps_id) {
        if (sps_equal(s, d->active_sps)) return 0;
        d->active_sps_id = H264_MAX_SPS + 1; d->active_pps_id = H264_MAX_PPS + 1;
        d->active_sps = NULL; d->active_pps = NULL;
    }
    *d->sps[id] = *s;
    return 0;
}
static int store_pps(h264_decoder_t *d, const h264_pps_t *p)
{
    int id = p->pps_id;
    if (!d->pps[id]) { d->pps[id] = (h264_pps_t *)h264_malloc(sizeof *p); if (!d->pps[id]) return -1; }
    else {
        h264_free((void *)d->pps[id]->fmo.group_id);     /* the stored copy owns the explicit slice group map */
        if (id == d->active_pps_id && p->sps_id != d->active_sps_id) d->active_pps_id = H264_MAX_PPS + 1;
    }
    *d->pps[id] = *p;
    if (id == d->active_pps_id) d->active_pps = d->pps[id];
    return 0;
}

static void select_sets(h264_decoder_t *d, uint32_t pps_id)
{
    d->active_pps_id = (int)pps_id; d->active_pps = d->pps[pps_id];
    d->active_sps_id = d->active_pps->sps_id; d->active_sps = d->sps[d->active_sps_id];
    d->width_mbs = d->active_sps->width_mbs; d->height_mbs = d->active_sps->height_mbs;
    d->pic_size_mbs = d->width_mbs * d->height_mbs;
    d->pending_activation = 1;
}

/* slice group parameters against the picture size (h264bsd_storage.c:801-850 CheckPps) */
static int check_fmo(const h264_pps_t *pps, const h264_sps_t *sps)
{
    const uint32_t size = sps->width_mbs * sps->height_mbs, W = sps->width_mbs;
    const h264_fmo_t *f = &pps->fmo;
    uint32_t i;
    if (pps->num_slice_groups <= 1) return 0;
    if (f->type == 0) { for (i = 0; i < f->n_groups; i++) if (f->run_length[i] > size) return -1; }
    else if (f->type == 2) {
        for (i = 0; i + 1 < f->n_groups; i++)
            if (f->top_left[i] > f->bottom_right[i] || f->bottom_right[i] >= size || f->top_left[i] % W > f->bottom_right[i] % W) return -1;
    } else if (f->type >= 3 && f->type <= 5) { if (f->change_rate > size) return -1; }
    else if (f->type == 6 && pps->fmo_map_units != size) return -1;
    return 0;
}

/* 0 ok, -1 bad combination, -2 allocation failure (h264bsd_storage.c:298-420) */
static int activate_param_sets(h264_decoder_t *d, uint32_t pps_id, int is_idr)
{
    if (!d->pps[pps_id] || !d->sps[d->pps[pps_id]->sps_id]) return -1;
    if (check_fmo(d->pps[pps_id], d->sps[d->pps[pps_id]->sps_id])) return -1;
    if (d->active_pps_id == H264_MAX_PPS) select_sets(d, pps_id);
    else if (d->pending_activation) {
        const h264_sps_t *sps = d->active_sps;
        int no_reorder;
        d->pending_activation = 0;
        h264_free(d->mbctx);
        d->mbctx = (h264_mbctx_t *)h264_calloc(d->pic_size_mbs, sizeof(h264_mbctx_t));
        if (!d->mbctx) return -2;
        h264_free(d->slice_group_map);
        d->slice_group_map = (uint8_t *)h264_malloc(d->pic_size_mbs);
        if (!d->slice_group_map) return -2;
        no_reorder = d->no_reordering_app || sps->poc_type == 2 ||
                     (sps->vui_present && sps->bitstream_restriction && !sps->num_reorder_frames);
        h264_dpb_init(&d->dpb, sps->max_dpb_size, sps->num_ref_frames, sps->max_frame_num, no_reorder);
        d->n_slots = d->dpb.dpb_size + 2;        /* the DPB, the picture being decoded, and a spare (h264_dpb_rotate_spare) */
        if (!d->be) d->be = h264_default_backend();   /* CUDA engine; NULL (reason on stderr) without a usable GPU */
        if (!d->be) return -2;
        if (d->be_inst) { d->be->inst_destroy(d->be, d->be_inst); d->be_inst = NULL; }
        {
            const int host = d->force_host_parse && d->be->inst_create_ex;
            d->be_inst = host ? d->be->inst_create_ex(d->be, d->width_mbs, d->height_mbs, d->n_slots, 1)
                              : d->be->inst_create(d->be, d->width_mbs, d->height_mbs, d->n_slots);
            if (!d->be_inst) return -2;
            d->device_parse = d->be->block_grow && d->be->parse_mode && !host;
        }
        d->pic = NULL;
        if (d->out_format && apply_output_format(d)) return -2;
    } else if ((int)pps_id != d->active_pps_id) {
        if (d->pps[pps_id]->sps_id != d->active_sps_id) {
            if (!is_idr) return -1;
            select_sets(d, pps_id);
        } else { d->active_pps_id = (int)pps_id; d->active_pps = d->pps[pps_id]; }
    }
    return 0;
}

/* cropped output rectangle of the active SPS (h264bsd_decoder.c:886-917), whole frame without cropping */
static int apply_output_format(h264_decoder_t *d)
{
    const h264_sps_t *p = d->active_sps;
    int l = 0, t = 0, w, h;
    if (!d->be || !d->be_inst || !p) return 0;
    if (!d->be->set_output) return d->out_format ? -1 : 0;
    w = 16 * (int)p->width_mbs; h = 16 * (int)p->height_mbs;
    if (p->crop_flag) { l = 2 * (int)p->crop_left; t = 2 * (int)p->crop_top; w -= 2 * (int)(p->crop_left + p->crop_right); h -= 2 * (int)(p->crop_top + p->crop_bottom); }
    return d->be->set_output(d->be, d->be_inst, d->out_format, l, t, w, h);
}

/* Give the macroblocks of the slice that just failed back to "not decoded" (h264bsd_slice_data.c:302-358
 * h264bsdMarkSliceCorrupted): a P slice loses all of them; an I slice keeps what lies more than
 * max(picture width, 10) of its macroblocks before the last good one. */
static void mark_slice_corrupted(h264_decoder_t *d, uint32_t first_mb)
{
    const uint32_t sid = d->slice_id, N = d->pic_size_mbs;
    uint32_t cur = first_mb;
    if (d->slice_last_mb) {
        uint32_t i = d->slice_last_mb - 1, cnt = 0, lim = d->width_mbs > 10 ? d->width_mbs : 10;
        while (i > cur) {
            if (d->mbctx[i].slice_id == sid && ++cnt >= lim) break;
            i--;
        }
        cur = i;
    }
    do {
        if (d->mbctx[cur].slice_id != sid || !d->mbctx[cur].decoded) break;
        d->mbctx[cur].decoded = 0;
        if (d->active_pps->num_slice_groups > 1) {
            const uint8_t *map = d->slice_group_map, grp = map[cur];
            do cur++; while (cur < N && map[cur] != grp);
            if (cur >= N) cur = 0;
        } else cur = cur + 1 < N ? cur + 1 : 0;
    } while (cur);
}

/* ------------------------------------------------------------ concealment */
/* Macroblocks of lost slices (h264bsd_conceal.c:125-255 h264bsdConceal, :262-330 ConcealMb), expressed as ordinary
 * records so the kernels need nothing new:
 *   - P picture with a reference available: the co-located macroblock of the reference picture with the smallest
 *     list index (zero vector, no residual), QP 40, all edges filtered, treated as intra by the deblocking filter;
 *   - nothing of the picture decoded: a copy of that reference picture, or mid-grey (128) for an I picture /
 *     no reference, with deblocking off everywhere;
 *   - a PARTLY lost I picture (spatial interpolation, h264bsd_conceal.c:330-631) is NOT implemented: those
 *     macroblocks are flagged MISSING and keep whatever the frame buffer held.
 * Returns the number of macroblocks concealed or missing. */
static uint32_t conceal_picture(h264_decoder_t *d)
{
    h264_pic_input_t *pic = d->pic;
    const int is_p = !d->valid_slice_in_au || d->sh.slice_type == 0;
    const int whole = d->num_decoded_mbs == 0;
    const uint32_t W = d->width_mbs, H = d->height_mbs, N = d->pic_size_mbs;
    uint32_t i, n = 0;
    int ref = -1;
    if (is_p) for (i = 0; i < H264_MAX_REFS && ref < 0; i++) ref = h264_dpb_ref_slot(&d->dpb, i);
    pic->n_conceal = 0;
    if (ref >= 0 || whole) {
        for (i = 0; i < N; i++) {
            h264b200_mb_t *r = &pic->mbs[i];
            if (d->mbctx[i].decoded) continue;
            memset(r, 0, sizeof *r);
            n++;
            if (ref >= 0) {
                int q;
                r->mb_class = H264B200_MB_INTER; r->part_flags = 31;
                for (q = 0; q < 4; q++) r->ref_slot[q] = (uint8_t)ref;
                r->qp_y = r->qp_dbk = 40; r->qp_c = H264_QPC[40];
                r->flags = H264B200_MBF_DBK_AS_INTRA;
                if (!whole) {
                    r->dbk_flags = (uint8_t)(H264B200_DBK_INNER | (i % W ? H264B200_DBK_LEFT : 0) | (i >= W ? H264B200_DBK_TOP : 0));
                    pic->any_deblock = 1;
                }
                pic->n_inter++; pic->ref_slots_used[ref] = 1;
            } else {                                      /* grey picture: raw samples, like I_PCM */
                if (pic->coef_used + 12 > pic->coef_cap && d->be->coef_grow(d->be, d->be_inst, pic, pic->coef_used + 12 * (N - i))) { pic->mbs[i].mb_class = H264B200_MB_MISSING; continue; }
                r = &pic->mbs[i];                         /* growing may move the records */
                r->mb_class = H264B200_MB_IPCM; r->coef_offset = pic->coef_used; r->nz_mask = 0xffff;
                memset(pic->coef + (size_t)pic->coef_used * 16, 128, 384);
                pic->coef_used += 12; pic->n_intra++;
            }
        }
        if (whole) for (i = 0; i < N; i++) pic->mbs[i].dbk_flags = 0;   /* no filtering of a fully concealed picture */
        return n;
    }
    /* spatial interpolation, in the reference's order (h264bsd_conceal.c:186-252): the row of the first decoded
     * macroblock (leftwards from it, then rightwards), the rows above it column by column going up, the rows below in
     * raster order; every concealed macroblock counts as available for the ones after it */
    {
        uint32_t lost = N - d->num_decoded_mbs, need = (lost * 4 + 31) / 32, *list, row, col, j;
        if (pic->coef_used + need > pic->coef_cap && d->be->coef_grow(d->be, d->be_inst, pic, pic->coef_used + need)) {
            for (i = 0; i < N; i++) if (!d->mbctx[i].decoded) { memset(&pic->mbs[i], 0, sizeof pic->mbs[i]); pic->mbs[i].mb_class = H264B200_MB_MISSING; }
            return lost;
        }
        list = (uint32_t *)(pic->coef + (size_t)pic->coef_used * 16);
        pic->conceal_offset = pic->coef_used;
        for (i = 0; i < N && !d->mbctx[i].decoded; i++) ;
        row = i / W; col = i % W;
#define CONCEAL_ONE(addr) do { \
            const uint32_t a_ = (addr), y_ = a_ / W, x_ = a_ % W; \
            h264b200_mb_t *r_ = &pic->mbs[a_]; \
            memset(r_, 0, sizeof *r_); \
            r_->mb_class = H264B200_MB_CONCEAL; \
            r_->avail = (uint8_t)((y_ > 0 && d->mbctx[a_ - W].decoded ? H264B200_CN_ABOVE : 0) | (y_ + 1 < H && d->mbctx[a_ + W].decoded ? H264B200_CN_BELOW : 0) | \
                                  (x_ > 0 && d->mbctx[a_ - 1].decoded ? H264B200_CN_LEFT : 0) | (x_ + 1 < W && d->mbctx[a_ + 1].decoded ? H264B200_CN_RIGHT : 0)); \
            r_->qp_y = r_->qp_dbk = 40; r_->qp_c = H264_QPC[40]; \
            r_->dbk_flags = (uint8_t)(H264B200_DBK_INNER | (x_ ? H264B200_DBK_LEFT : 0) | (y_ ? H264B200_DBK_TOP : 0)); \
            d->mbctx[a_].decoded = 1; list[n++] = a_; } while (0)
        for (j = col; j-- > 0;) CONCEAL_ONE(row * W + j);
        for (j = col + 1; j < W; j++) if (!d->mbctx[row * W + j].decoded) CONCEAL_ONE(row * W + j);
        if (row) for (j = 0; j < W; j++) for (i = row; i-- > 0;) CONCEAL_ONE(i * W + j);
        for (i = row + 1; i < H; i++) for (j = 0; j < W; j++) if (!d->mbctx[i * W + j].decoded) CONCEAL_ONE(i * W + j);
#undef CONCEAL_ONE
        pic->n_conceal = n;
        pic->coef_used += (n * 4 + 31) / 32;
        pic->any_deblock = 1;
    }
    return n;
}

/* ------------------------------------------------ device-parse path: queue a slice */
/* Append the slice NAL (RBSP, emulation prevention already removed) and what the kernel needs from its header to the
 * picture's block (include/h264b200_slices.h).  `b` stands at the first bit of slice_data(). */
static int enqueue_slice(h264_decoder_t *d, const br_t *b, const h264_slice_hdr_t *sh)
{
    h264_pic_input_t *pic = d->pic;
    const uint32_t rbsp_len = (uint32_t)b->len, N = d->pic_size_mbs;
    const int fmo = d->active_pps->num_slice_groups > 1;
    const uint32_t rb_pad = (rbsp_len + 15u) & ~15u, map_pad = fmo ? (N + 15u) & ~15u : 0;
    const uint32_t need = (uint32_t)sizeof(h264b200_slice_t) + rb_pad + map_pad;
    h264b200_slice_t *sl; uint8_t *p; uint32_t i;
    if (pic->block_used + need > pic->block_cap && d->be->block_grow(d->be, d->be_inst, pic, pic->block_used + need)) return -1;
    p = pic->block + pic->block_used;
    sl = (h264b200_slice_t *)p;
    memset(sl, 0, sizeof *sl);
    sl->size = need; sl->rbsp_len = rbsp_len;
    sl->bit_off = (uint32_t)br_pos(b); sl->payload_bits = (uint32_t)b->payload_bits;
    sl->first_mb = sh->first_mb;
    sl->slice_id = (uint16_t)(++d->slice_id);
    sl->is_p = sh->slice_type == 0;
    sl->num_ref_idx_active = (uint8_t)sh->num_ref_idx_active;
    sl->slice_qp = (int8_t)sh->slice_qp;
    sl->chroma_qp_off = (int8_t)d->active_pps->chroma_qp_index_offset;
    sl->alpha_off = sh->alpha_off; sl->beta_off = sh->beta_off;
    sl->disable_deblocking_idc = sh->disable_deblocking_idc;
    sl->constrained_intra = d->active_pps->constrained_intra_pred;
    for (i = 0; i <= H264_MAX_REFS; i++) sl->ref_slot[i] = -1;
    if (sl->is_p) for (i = 0; i < sh->num_ref_idx_active && i <= H264_MAX_REFS; i++) sl->ref_slot[i] = (int8_t)h264_dpb_ref_slot(&d->dpb, i);
    memcpy(p + sizeof *sl, b->data, rbsp_len);
    memset(p + sizeof *sl + rbsp_len, 0, rb_pad - rbsp_len);
    if (fmo) {
        sl->map_off = (uint32_t)sizeof *sl + rb_pad;
        memcpy(p + sl->map_off, d->slice_group_map, N);
        memset(p + sl->map_off + N, 0, map_pad - N);
    }
    ((h264b200_pichdr_t *)pic->block)->n_slices++;
    pic->block_used += need;
    if (sl->is_p) pic->has_p_slice = 1;
    return 0;
}

/* ------------------------------------------------------------ picture end */
static void finish_picture(h264_decoder_t *d)
{
    int is_idr = d->pic_nal_type == NAL_IDR;
    int32_t poc;
    if (d->pic) {
        if (d->device_parse) {
            /* kernel Kp finds out which macroblocks the slices delivered and conceals the rest (kp_core.h
             * kp_conceal_picture); what it needs from the DPB goes into the picture header.  The number of
             * concealed macroblocks comes back with the frame (h264b200_picstat_t). */
            h264b200_pichdr_t *hdr = (h264b200_pichdr_t *)d->pic->block;
            const int is_p = !d->valid_slice_in_au || d->sh.slice_type == 0;
            int ref = -1; uint32_t i;
            if (is_p) for (i = 0; i < H264_MAX_REFS && ref < 0; i++) ref = h264_dpb_ref_slot(&d->dpb, i);
            hdr->conceal_as_p = (uint8_t)is_p; hdr->conceal_ref_slot = (int8_t)ref;
            hdr->total_bytes = d->pic->block_used;
            d->num_err_mbs = 0;
        } else
        if (d->num_decoded_mbs != d->pic_size_mbs) d->num_err_mbs = conceal_picture(d);    /* lost slices */
        d->pic->cur_slot = h264_dpb_current_slot(&d->dpb);
        d->be->pic_submit(d->be, d->be_inst, d->pic);
        d->pic = NULL;
    }
    d->num_decoded_mbs = 0; d->slice_id = 0;
    poc = h264_decode_poc(d, &d->sh, d->pic_nal_type, d->pic_nal_ref_idc);
    if (d->valid_slice_in_au)
        h264_dpb_mark(&d->dpb, &d->sh, d->pic_nal_ref_idc != 0, is_idr, poc, d->current_pic_id, d->num_err_mbs);
    d->pic_started = 0; d->valid_slice_in_au = 0;
}

static int begin_picture(h264_decoder_t *d)
{
    d->pic = d->be->pic_begin(d->be, d->be_inst);
    if (!d->pic) return -1;
    h264_dpb_rotate_spare(&d->dpb);
    d->pic->coef_used = 0; d->pic->n_intra = d->pic->n_inter = 0; d->pic->any_deblock = 0; d->pic->n_conceal = 0; d->pic->conceal_offset = 0;
    memset(d->pic->ref_slots_used, 0, sizeof d->pic->ref_slots_used);
    d->pic->block_used = 0; d->pic->has_p_slice = 0;
    d->num_decoded_mbs = 0; d->slice_id = 0;
    if (d->device_parse) {                      /* the block starts with the picture header; slices follow (enqueue_slice) */
        h264b200_pichdr_t *hdr;
        if (!d->pic->block || d->pic->block_cap < sizeof *hdr) {
            if (d->be->block_grow(d->be, d->be_inst, d->pic, 4096) || !d->pic->block) return -1;
        }
        hdr = (h264b200_pichdr_t *)d->pic->block;
        memset(hdr, 0, sizeof *hdr);
        hdr->magic = H264B200_PICHDR_MAGIC;
        hdr->width_mbs = (uint16_t)d->width_mbs; hdr->height_mbs = (uint16_t)d->height_mbs;
        hdr->conceal_ref_slot = -1;
        d->pic->block_used = sizeof *hdr;
        return 0;
    }
    memset(d->mbctx, 0, d->pic_size_mbs * sizeof(h264_mbctx_t));   /* records of unparsed macroblocks are flagged MISSING in finish_picture */
    return 0;
}

/* ------------------------------------------------------------------ API */
u32 h264_decoder_create(storage_t *pStorage, u32 noOutputReordering, h264_backend_t *be)
{
    h264_decoder_t *d;
    if (!pStorage) return HANTRO_NOK;
    memset(pStorage, 0, sizeof *pStorage);
    d = (h264_decoder_t *)h264_calloc(1, sizeof *d);
    if (!d) return HANTRO_NOK;
    h264_cavlc_init();
    d->be = be;
    d->active_sps_id = H264_MAX_SPS; d->active_pps_id = H264_MAX_PPS; d->old_sps_id = H264_MAX_SPS + 1;
    d->aub.first_call = 1;
    d->no_reordering_app = noOutputReordering ? 1 : 0;
    pStorage->impl = d;
    return HANTRO_OK;
}

u32 h264bsdInit(storage_t *pStorage, u32 noOutputReordering)
{
    return h264_decoder_create(pStorage, noOutputReordering, NULL);
}

u32 h264bsdDecode(storage_t *pStorage, u8 *byteStrm, u32 len, u32 picId, u32 *readBytes)
{
    h264_decoder_t *d = DEC(pStorage);
    uint8_t *nal; uint32_t nal_len, consumed;
    int type, ref_idc, boundary = 0, pic_ready = 0, rc;
    br_t b;
    if (!d || !byteStrm || !len || !readBytes) return H264BSD_ERROR;

    if (d->prev_buf_not_finished && byteStrm == d->prev_buf_ptr) {
        nal = (uint8_t *)d->nal_data; nal_len = (uint32_t)d->nal_len;
        *readBytes = d->prev_bytes_consumed;
    } else {
        if (extract_nal(byteStrm, len, &nal, &nal_len, &consumed, d->ro_input ? (nal_scratch_t *)&d->nal_scratch : NULL)) { *readBytes = consumed; return H264BSD_ERROR; }
        *readBytes = consumed;
        d->nal_data = nal; d->nal_len = nal_len;
        d->prev_bytes_consumed = consumed; d->prev_buf_ptr = byteStrm;
    }
    d->prev_buf_not_finished = 0;
    if (nal_len < 1) return H264BSD_ERROR;
    ref_idc = (nal[0] >> 5) & 3; type = nal[0] & 31;
    if (type == 2 || type == 3 || type == 4) return H264BSD_ERROR;       /* data partitioning: not Baseline decoder scope */
    if ((type == NAL_SPS || type == NAL_PPS || type == NAL_IDR) && ref_idc == 0) return H264BSD_ERROR;
    if ((type == NAL_SEI || type == NAL_AUD || type == NAL_EOSEQ || type == NAL_EOSTREAM || type == NAL_FILLER) && ref_idc != 0) return H264BSD_ERROR;
    if (type == 0 || type >= 13) return H264BSD_RDY;
    br_init(&b, nal + 1, nal_len - 1);

    rc = check_au_boundary(d, b, type, ref_idc, &boundary);
    if (rc) return rc == -2 ? H264BSD_PARAM_SET_ERROR : H264BSD_ERROR;
    if (boundary) {
        if (d->pic_started && d->active_sps) {
            /* a new access unit starts while the previous picture is incomplete: the reference
             * conceals here (h264bsd_decoder.c:236-270). Concealment is out of scope: the missing
             * macroblocks are counted and the picture is finished as it is. */
            if (d->pending_activation) return H264BSD_ERROR;
            if (!d->valid_slice_in_au) {
                if (begin_picture(d)) return H264BSD_MEMALLOC_ERROR;
                h264_dpb_init_ref_list(&d->dpb);
            }
            d->num_err_mbs = d->pic_size_mbs - d->num_decoded_mbs;
            pic_ready = 1;
            *readBytes = 0; d->prev_buf_not_finished = 1;
        } else d->valid_slice_in_au = 0;
        d->skip_redundant = 0;
    }

    if (!pic_ready) switch (type) {
    case NAL_SPS: {
        h264_sps_t *sps = (h264_sps_t *)h264_malloc(sizeof *sps);
        if (!sps) return H264BSD_MEMALLOC_ERROR;
        rc = h264_parse_sps(&b, sps);
        if (!rc) rc = store_sps(d, sps);
        h264_free(sps);
        if (rc) return H264BSD_ERROR;
        break; }
    case NAL_PPS: {
        h264_pps_t pps;
        if (h264_parse_pps(&b, &pps)) { h264_free((void *)pps.fmo.group_id); return H264BSD_ERROR; }
        if (store_pps(d, &pps)) { h264_free((void *)pps.fmo.group_id); return H264BSD_MEMALLOC_ERROR; }
        break; }
    case NAL_IDR:
    case NAL_SLICE: {
        h264_slice_hdr_t sh;
        int start_of_pic;
        if (d->skip_redundant) return H264BSD_RDY;
        d->pic_started = 1;
        start_of_pic = !d->valid_slice_in_au;
        if (start_of_pic) {
            uint32_t pps_id; int old_sps;
            d->num_err_mbs = 0; d->current_pic_id = picId;
            if (h264_peek_pps_id(b, &pps_id)) return H264BSD_ERROR;
            old_sps = d->active_sps_id;
            rc = activate_param_sets(d, pps_id, type == NAL_IDR);
            if (rc) {
                d->active_pps_id = H264_MAX_PPS; d->active_pps = NULL;
                d->active_sps_id = H264_MAX_SPS; d->active_sps = NULL;
                d->pending_activation = 0;
                return rc == -2 ? H264BSD_MEMALLOC_ERROR : H264BSD_PARAM_SET_ERROR;
            }
            if (old_sps != d->active_sps_id) {
                const h264_sps_t *old = d->old_sps_id < H264_MAX_SPS ? d->sps[d->old_sps_id] : NULL, *nw = d->active_sps;
                int no_output = 1, ok = 0;
                *readBytes = 0; d->prev_buf_not_finished = 1;
                if (type == NAL_IDR) {
                    h264_slice_hdr_t tmp;
                    if (!h264_parse_slice_header(&b, &tmp, nw, d->active_pps, type, ref_idc)) { ok = 1; no_output = tmp.no_output_of_prior_pics; }
                }
                if (!ok || no_output || d->dpb.no_reordering || !old || old->width_mbs != nw->width_mbs ||
                    old->height_mbs != nw->height_mbs || old->max_dpb_size != nw->max_dpb_size) d->dpb.flushed = 0;
                else h264_dpb_flush(&d->dpb);
                d->old_sps_id = d->active_sps_id;
                return H264BSD_HDRS_RDY;
            }
        }
        if (d->pending_activation) return H264BSD_ERROR;
        if (h264_parse_slice_header(&b, &sh, d->active_sps, d->active_pps, type, ref_idc)) return H264BSD_ERROR;
        if (start_of_pic) {
            if (type != NAL_IDR && h264_dpb_check_gaps(&d->dpb, sh.frame_num, ref_idc != 0, d->active_sps->gaps_allowed)) return H264BSD_ERROR;
            if (begin_picture(d)) return H264BSD_MEMALLOC_ERROR;
        }
        if (d->active_pps->num_slice_groups > 1) {
            const h264_fmo_t *f = &d->active_pps->fmo;
            uint32_t units0 = sh.slice_group_change_cycle * f->change_rate;
            if (units0 > d->pic_size_mbs) units0 = d->pic_size_mbs;
            h264_fmo_build_map(d->slice_group_map, d->width_mbs, d->height_mbs, f, units0);
        }
        d->sh = sh; d->valid_slice_in_au = 1;
        d->pic_nal_type = (uint8_t)type; d->pic_nal_ref_idc = (uint8_t)ref_idc;
        if (sh.redundant_pic_cnt) break;          /* redundant coded pictures are not decoded */
        h264_dpb_init_ref_list(&d->dpb);
        if (h264_dpb_reorder(&d->dpb, &d->sh)) return H264BSD_ERROR;
        if (d->device_parse) {
            /* the picture ends when the next access unit begins (or at h264bsdFlushBuffer): how many macroblocks
             * the slice holds is only known once kernel Kp has parsed it */
            if (enqueue_slice(d, &b, &d->sh)) return H264BSD_MEMALLOC_ERROR;
            break;
        }
        if (h264_decode_slice_data(d, &b, &d->sh)) {
            /* the macroblocks of a slice that failed are given back (they are concealed when the access unit ends) */
            mark_slice_corrupted(d, d->sh.first_mb);
            return H264BSD_ERROR;
        }
        if (d->num_decoded_mbs == d->pic_size_mbs) { pic_ready = 1; d->skip_redundant = 1; }
        break; }
    default: break;                                /* SEI, AUD, end of sequence/stream, filler: ignored */
    }

    if (pic_ready) { finish_picture(d); return H264BSD_PIC_RDY; }
    return H264BSD_RDY;
}

u8 *h264bsdNextOutputPicture(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs)
{
    h264_decoder_t *d = DEC(pStorage);
    const h264_out_t *o;
    uint32_t err = 0;
    uint8_t *p;
    if (!d || !d->dpb.allocated || !d->be_inst) return NULL;
    for (;;) {
        h264b200_picstat_t ps;
        o = h264_dpb_next_output(&d->dpb);
        if (!o) return NULL;
        p = d->be->frame_host(d->be, d->be_inst, o->slot, &err);
        if (picId) *picId = o->pic_id;
        if (isIdrPic) *isIdrPic = o->is_idr;
        if (numErrMbs) *numErrMbs = o->num_err_mbs;
        if (d->device_parse && p && d->be->frame_status && !d->be->frame_status(d->be, d->be_inst, o->slot, &ps)) {
            if (ps.flags & H264B200_PS_DROPPED) continue;     /* incomplete last picture of the stream: never became a picture */
            if (numErrMbs) *numErrMbs = ps.err_mbs;
        }
        return p;
    }
}

u32 h264b200SetOutputFormat(storage_t *pStorage, u32 format)
{
    h264_decoder_t *d = DEC(pStorage);
    if (!d || format > H264B200_OUT_RGBA) return HANTRO_NOK;
    d->out_format = (int)format;
    return apply_output_format(d) ? HANTRO_NOK : HANTRO_OK;
}

/* non-blocking pop + explicit wait (include/h264b200_batch.h) */
u8 *h264b200NextOutputPictureAsync(storage_t *pStorage, u32 *picId, u32 *isIdrPic, u32 *numErrMbs, u32 *ticket)
{
    h264_decoder_t *d = DEC(pStorage);
    const h264_out_t *o;
    uint32_t err = 0;
    if (!d || !d->dpb.allocated || !d->be_inst || !ticket) return NULL;
    o = h264_dpb_next_output(&d->dpb);
    if (!o) return NULL;
    *ticket = (u32)o->slot;
    if (picId) *picId = o->pic_id;
    if (isIdrPic) *isIdrPic = o->is_idr;
    if (numErrMbs) *numErrMbs = o->num_err_mbs;
    if (d->be->frame_host_async && d->be->frame_wait) {
        uint32_t gen = 0;
        u8 *p = d->be->frame_host_async(d->be, d->be_inst, o->slot, &gen);
        *ticket = (u32)o->slot | (gen << 8);          /* which picture of the slot: slots are re-used */
        return p;
    }
    return d->be->frame_host(d->be, d->be_inst, o->slot, &err);
}
u32 h264b200PictureWait(storage_t *pStorage, u32 ticket)
{
    h264_decoder_t *d = DEC(pStorage);
    uint32_t err = 0;
    if (!d || !d->be_inst) return 0xffffffffu;
    if (d->be->frame_host_async && d->be->frame_wait) {
        int rc = d->be->frame_wait(d->be, d->be_inst, (int)(ticket & 0xff), ticket >> 8, &err);
        if (rc < 0) return 0xffffffffu;
        if (rc == 2) return H264B200_WAIT_NOT_LAUNCHED; /* still queued in the engine (device-parse look-ahead) */
        if (rc > 0) return 0xfffffffeu;               /* waited too long: a later picture already occupies the slot */
        return err;
    }
    if (!d->be->frame_host(d->be, d->be_inst, (int)(ticket & 0xff), &err)) return 0xffffffffu;
    return err;
}

/* Never blocks: 0 the picture behind `ticket` is complete in its host buffer, 1 launched and in flight, 2 still queued in
 * the engine, 0xffffffff on an error.  A backend without the query answers 1 for anything launched (h264b200PictureWait decides). */
u32 h264b200PictureState(storage_t *pStorage, u32 ticket)
{
    h264_decoder_t *d = DEC(pStorage);
    int rc;
    if (!d || !d->be_inst) return 0xffffffffu;
    if (!d->be->frame_state) return 1;
    rc = d->be->frame_state(d->be, d->be_inst, (int)(ticket & 0xff), ticket >> 8);
    return rc < 0 ? 0xffffffffu : (u32)rc;
}

/* 0 and the status words of the picture behind `ticket` (after h264b200PictureWait returned 0); 1 if the backend keeps none */
u32 h264b200PictureStatus(storage_t *pStorage, u32 ticket, h264b200_picstat_t *out)
{
    h264_decoder_t *d = DEC(pStorage);
    if (!d || !d->be_inst || !out || !d->device_parse || !d->be->frame_status) return 1;
    return d->be->frame_status(d->be, d->be_inst, (int)(ticket & 0xff), out) ? 1 : 0;
}
void h264b200PictureRelease(storage_t *pStorage, u32 ticket)
{
    h264_decoder_t *d = DEC(pStorage);
    if (d && d->be_inst && d->be->frame_release) d->be->frame_release(d->be, d->be_inst, (int)(ticket & 0xff), ticket >> 8);
}
u32 h264b200PicturesPending(storage_t *pStorage)
{
    h264_decoder_t *d = DEC(pStorage);
    return d && d->be_inst && d->be->inst_pending ? d->be->inst_pending(d->be, d->be_inst) : 0;
}
/* The decoder stops editing the caller's buffers: a NAL unit that contains emulation prevention bytes is unescaped
 * into decoder-owned scratch memory instead of in place (the reference always edits in place,
 * h264bsd_byte_stream.c:192-234; callers that share one read-only copy of a stream between instances need this). */
void h264b200SetReadOnlyInput(storage_t *pStorage, u32 on) { h264_decoder_t *d = DEC(pStorage); if (d) d->ro_input = on != 0; }

/* This instance parses slice data on the host cores even on a device-parse engine (call before the first parameter
 * sets are activated).  Both parsers write the same records, so the pictures are the same either way; what changes is
 * who does the work: h264b200DecodeStreams gives the otherwise idle parser threads a share of the streams. */
void h264b200SetHostParse(storage_t *pStorage, u32 on) { h264_decoder_t *d = DEC(pStorage); if (d) d->force_host_parse = on != 0; }

u32 h264b200DeviceParse(storage_t *pStorage) { h264_decoder_t *d = DEC(pStorage); return d ? (u32)d->device_parse : 0; }

void h264bsdShutdown(storage_t *pStorage)
{
    h264_decoder_t *d = DEC(pStorage);
    int i;
    if (!d) return;
    if (d->be_inst) d->be->inst_destroy(d->be, d->be_inst);
    for (i = 0; i < H264_MAX_SPS; i++) h264_free(d->sps[i]);
    for (i = 0; i < H264_MAX_PPS; i++) { if (d->pps[i]) h264_free((void *)d->pps[i]->fmo.group_id); h264_free(d->pps[i]); }
    h264_free(d->mbctx); h264_free(d->slice_group_map); h264_free(d->nal_scratch.p);
    h264_free(d);
    pStorage->impl = NULL;
}

u32 h264bsdPicWidth(storage_t *s)  { h264_decoder_t *d = DEC(s); return d && d->active_sps ? d->active_sps->width_mbs : 0; }
u32 h264bsdPicHeight(storage_t *s) { h264_decoder_t *d = DEC(s); return d && d->active_sps ? d->active_sps->height_mbs : 0; }
u32 h264bsdVideoRange(storage_t *s)
{
    h264_decoder_t *d = DEC(s);
    return d && d->active_sps && d->active_sps->vui_present && d->active_sps->video_signal_present && d->active_sps->video_full_range;
}
u32 h264bsdMatrixCoefficients(storage_t *s)
{
    h264_decoder_t *d = DEC(s);
    if (d && d->active_sps && d->active_sps->vui_present && d->active_sps->video_signal_present && d->active_sps->colour_desc_present)
        return d->active_sps->matrix_coefficients;
    return 2;
}
void h264bsdCroppingParams(storage_t *s, u32 *croppingFlag, u32 *left, u32 *width, u32 *top, u32 *height)
{
    h264_decoder_t *d = DEC(s);
    if (d && d->active_sps && d->active_sps->crop_flag) {
        const h264_sps_t *p = d->active_sps;
        *croppingFlag = 1;
        *left = 2 * p->crop_left; *width = 16 * p->width_mbs - 2 * (p->crop_left + p->crop_right);
        *top = 2 * p->crop_top;   *height = 16 * p->height_mbs - 2 * (p->crop_top + p->crop_bottom);
    } else { *croppingFlag = 0; *left = 0; *width = 0; *top = 0; *height = 0; }
}
void h264bsdSampleAspectRatio(storage_t *s, u32 *sarWidth, u32 *sarHeight)
{
    static const uint8_t tab[14][2] = {{0,0},{1,1},{12,11},{10,11},{16,11},{40,33},{24,11},{20,11},{32,11},{80,33},{18,11},{15,11},{64,33},{160,99}};
    h264_decoder_t *d = DEC(s);
    u32 w = 1, h = 1;
    if (d && d->active_sps && d->active_sps->vui_present && d->active_sps->aspect_ratio_present) {
        const h264_sps_t *p = d->active_sps;
        if (p->aspect_ratio_idc < 14) { w = tab[p->aspect_ratio_idc][0]; h = tab[p->aspect_ratio_idc][1]; }
        else if (p->aspect_ratio_idc == 255) { w = p->sar_width; h = p->sar_height; if (!w || !h) w = h = 0; }
        else w = h = 0;
    }
    *sarWidth = w; *sarHeight = h;
}
u32 h264bsdCheckValidParamSets(storage_t *s)
{
    h264_decoder_t *d = DEC(s);
    int i;
    if (!d) return 0;
    for (i = 0; i < H264_MAX_PPS; i++) if (d->pps[i] && d->sps[d->pps[i]->sps_id]) return 1;
    return 0;
}
void h264bsdFlushBuffer(storage_t *s)
{
    h264_decoder_t *d = DEC(s);
    if (!d) return;
    /* device-parse: the last picture of the stream has no following access unit to end it */
    if (d->device_parse && d->pic_started && d->valid_slice_in_au && d->pic && !d->pending_activation) {
        ((h264b200_pichdr_t *)d->pic->block)->tentative = 1;
        finish_picture(d);
    }
    h264_dpb_flush(&d->dpb);
}
u32 h264bsdProfile(storage_t *s) { h264_decoder_t *d = DEC(s); return d && d->active_sps ? d->active_sps->profile_idc : 0; }

/* used by the H264SwDec layer (h264_swdec.c): DPB flags the reference's API pokes directly */
int h264_decoder_flushed_pending(storage_t *s)
{
    h264_decoder_t *d = DEC(s);
    if (d && d->dpb.flushed && d->dpb.num_out != d->dpb.out_index) { d->dpb.flushed = 0; return 1; }
    return 0;
}
