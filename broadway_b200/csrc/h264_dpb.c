/* h264_dpb.c — decoded picture buffer bookkeeping (ITU-T H.264 8.2.4, 8.2.5,
 * C.4) with the output policy of the reference (h264bsd_dpb.c): dpb_size+1
 * entries kept sorted so that RefPicList0 initialisation is a prefix copy
 * (:1099-1113), sliding window + MMCO 1-6 (:321-598, :628-830), gaps in
 * frame_num (:1244-1372), smallest-POC-first bumping when fullness exceeds
 * dpb_size and immediate output in no-reordering mode (:805-823, :1380-1455).
 *
 * Redesign: the reference moves dpbPicture_t structs and their malloc'ed `data`
 * pointers around and identifies reference pictures by pointer
 * (h264bsd_dpb.c:826, :1558-1584).  Frame storage here lives on the GPU, so an
 * entry carries a stable frame-pool SLOT id instead; macroblock records name
 * reference pictures by slot, and "same reference buffer" (bS derivation,
 * h264bsd_deblocking.c:348,402) becomes "same slot".
 */
#include <string.h>
#include "h264_internal.h"

#define IS_REF(p)   ((p).status > PIC_UNUSED)
#define IS_EXIST(p) ((p).status > PIC_NON_EXISTING)
#define IS_ST(p)    ((p).status == PIC_SHORT || (p).status == PIC_NON_EXISTING)

static void set_unused(h264_dpb_t *d, h264_dpb_pic_t *p)
{
    p->status = PIC_UNUSED;
    d->num_ref_frames--;
    if (!p->to_be_displayed) d->fullness--;
}

/* order: short-term by descending PicNum, long-term by ascending LongTermPicNum,
 * non-reference waiting for output, the rest */
static int before(const h264_dpb_pic_t *a, const h264_dpb_pic_t *b)
{
    int ra = IS_REF(*a), rb = IS_REF(*b);
    if (!ra && !rb) return a->to_be_displayed && !b->to_be_displayed;
    if (!rb) return 1;
    if (!ra) return 0;
    if (IS_ST(*a) && IS_ST(*b)) return a->pic_num > b->pic_num;
    if (IS_ST(*a)) return 1;
    if (IS_ST(*b)) return 0;
    return a->pic_num < b->pic_num;
}
static void sort_buf(h264_dpb_t *d)
{
    uint32_t n = d->dpb_size + 1, i, j;
    for (i = 1; i < n; i++) {
        h264_dpb_pic_t t = d->buf[i];
        for (j = i; j > 0 && before(&t, &d->buf[j - 1]); j--) d->buf[j] = d->buf[j - 1];
        d->buf[j] = t;
    }
}

void h264_dpb_init(h264_dpb_t *d, uint32_t dpb_size, uint32_t max_ref_frames, uint32_t max_frame_num, int no_reordering)
{
    uint32_t i;
    memset(d, 0, sizeof *d);
    d->max_long_term_idx = H264_NO_LONG_TERM;
    d->max_ref_frames = max_ref_frames ? max_ref_frames : 1;
    d->dpb_size = no_reordering ? d->max_ref_frames : dpb_size;
    if (d->dpb_size < d->max_ref_frames) d->dpb_size = d->max_ref_frames;
    d->max_frame_num = max_frame_num;
    d->no_reordering = (uint8_t)no_reordering;
    for (i = 0; i <= d->dpb_size; i++) d->buf[i].slot = (int)i;
    d->spare_slot = (int)d->dpb_size + 1;
    d->allocated = 1;
}

int h264_dpb_current_slot(h264_dpb_t *d) { return d->buf[d->dpb_size].slot; }

/* Called once per picture before anything names the current slot.  The reference decodes into the buffer the DPB just
 * freed (dpbSize + 1 buffers, h264bsd_dpb.c:1130-1180), which here would be the frame whose copy-out to the host — or whose
 * reader, after h264b200NextOutputPictureAsync — may still be busy with it, so the picture would have to wait.  With one
 * buffer more than the DPB needs the freed slot rests for a picture: the new picture takes the spare, the freed slot
 * becomes the spare.  Slots are storage identities only; which picture is a reference or waits for output is unchanged. */
void h264_dpb_rotate_spare(h264_dpb_t *d)
{
    const int s = d->buf[d->dpb_size].slot;
    if (d->spare_slot < 0) return;
    d->buf[d->dpb_size].slot = d->spare_slot;
    d->spare_slot = s;
}

static int output_one(h264_dpb_t *d)
{
    h264_dpb_pic_t *best = NULL;
    uint32_t i;
    if (d->no_reordering) return -1;
    for (i = 0; i <= d->dpb_size; i++)
        if (d->buf[i].to_be_displayed && (!best || d->buf[i].poc < best->poc)) best = &d->buf[i];
    if (!best) return -1;
    d->out[d->num_out].slot = best->slot; d->out[d->num_out].is_idr = best->is_idr;
    d->out[d->num_out].pic_id = best->pic_id; d->out[d->num_out].num_err_mbs = best->num_err_mbs;
    d->num_out++;
    best->to_be_displayed = 0;
    if (!IS_REF(*best)) d->fullness--;
    return 0;
}

static void set_pic_nums(h264_dpb_t *d, uint32_t cur_frame_num)
{
    uint32_t i;
    for (i = 0; i < d->num_ref_frames; i++) if (IS_ST(d->buf[i]))
        d->buf[i].pic_num = d->buf[i].frame_num > cur_frame_num ? (int32_t)d->buf[i].frame_num - (int32_t)d->max_frame_num
                                                                  : (int32_t)d->buf[i].frame_num;
}

static int sliding_window(h264_dpb_t *d)
{
    int idx = -1; int32_t pn = 0; uint32_t i;
    if (d->num_ref_frames < d->max_ref_frames) return 0;
    for (i = 0; i < d->num_ref_frames; i++)
        if (IS_ST(d->buf[i]) && (idx < 0 || d->buf[i].pic_num < pn)) { idx = (int)i; pn = d->buf[i].pic_num; }
    if (idx < 0) return -1;
    set_unused(d, &d->buf[idx]);
    return 0;
}

static int find_pic(h264_dpb_t *d, int32_t pic_num, int short_term)
{
    uint32_t i;
    for (i = 0; i < d->max_ref_frames; i++) {
        int st = d->buf[i].status;
        if (short_term ? (st == PIC_SHORT || st == PIC_NON_EXISTING) : st == PIC_LONG)
            if (d->buf[i].pic_num == pic_num) return (int)i;
    }
    return -1;
}

int h264_dpb_check_gaps(h264_dpb_t *d, uint32_t frame_num, int is_ref, int gaps_allowed)
{
    d->num_out = 0; d->out_index = 0;
    if (!gaps_allowed) return 0;
    if (frame_num != d->prev_ref_frame_num && frame_num != (d->prev_ref_frame_num + 1) % d->max_frame_num) {
        uint32_t unused = (d->prev_ref_frame_num + 1) % d->max_frame_num;
        int keep_slot = d->buf[d->dpb_size].slot;
        do {
            h264_dpb_pic_t *c;
            set_pic_nums(d, unused);
            if (sliding_window(d)) return -1;
            while (d->fullness >= d->dpb_size) if (output_one(d)) break;
            c = &d->buf[d->dpb_size];
            c->status = PIC_NON_EXISTING; c->frame_num = unused; c->pic_num = (int32_t)unused; c->poc = 0; c->to_be_displayed = 0;
            d->fullness++; d->num_ref_frames++;
            sort_buf(d);
            unused = (unused + 1) % d->max_frame_num;
        } while (unused != frame_num);
        /* the frame about to be decoded must not land in a slot that was just queued for output */
        if (d->num_out) {
            uint32_t i, k;
            for (i = 0; i < d->num_out; i++) if (d->out[i].slot == d->buf[d->dpb_size].slot) {
                for (k = 0; k < d->dpb_size; k++) if (d->buf[k].slot == keep_slot) {
                    d->buf[k].slot = d->buf[d->dpb_size].slot; d->buf[d->dpb_size].slot = keep_slot; break;
                }
                break;
            }
        }
    } else if (is_ref && frame_num == d->prev_ref_frame_num) return -1;
    if (is_ref) d->prev_ref_frame_num = frame_num;
    else if (frame_num != d->prev_ref_frame_num) d->prev_ref_frame_num = (frame_num + d->max_frame_num - 1) % d->max_frame_num;
    return 0;
}

void h264_dpb_init_ref_list(h264_dpb_t *d)
{
    uint32_t i;
    for (i = 0; i < d->num_ref_frames; i++) d->list[i] = &d->buf[i];
}

int h264_dpb_reorder(h264_dpb_t *d, const h264_slice_hdr_t *sh)
{
    uint32_t i, j, k, ref_idx = 0, pred = sh->frame_num, n = sh->num_ref_idx_active;
    set_pic_nums(d, sh->frame_num);
    if (!sh->reorder_flag) return 0;
    for (i = 0; i < sh->n_reorder; i++) {
        int32_t pic_num; int idx, short_term;
        if (sh->reorder[i].idc < 2) {
            int32_t nowrap;
            if (sh->reorder[i].idc == 0) { nowrap = (int32_t)pred - (int32_t)sh->reorder[i].val; if (nowrap < 0) nowrap += (int32_t)d->max_frame_num; }
            else { nowrap = (int32_t)(pred + sh->reorder[i].val); if (nowrap >= (int32_t)d->max_frame_num) nowrap -= (int32_t)d->max_frame_num; }
            pred = (uint32_t)nowrap;
            pic_num = nowrap;
            if ((uint32_t)nowrap > sh->frame_num) pic_num -= (int32_t)d->max_frame_num;
            short_term = 1;
        } else { pic_num = (int32_t)sh->reorder[i].val; short_term = 0; }
        idx = find_pic(d, pic_num, short_term);
        if (idx < 0 || !IS_EXIST(d->buf[idx])) return -1;
        if (n > H264_MAX_REFS) return -1;
        for (j = n; j > ref_idx; j--) d->list[j] = d->list[j - 1];
        d->list[ref_idx++] = &d->buf[idx];
        for (j = k = ref_idx; j <= n; j++) if (d->list[j] != &d->buf[idx]) d->list[k++] = d->list[j];
    }
    return 0;
}

int h264_dpb_ref_slot(const h264_dpb_t *d, uint32_t ref_idx)
{
    if (ref_idx > 16 || !d->list[ref_idx] || !IS_EXIST(*d->list[ref_idx])) return -1;
    return d->list[ref_idx]->slot;
}

static void mmco5(h264_dpb_t *d)
{
    uint32_t i;
    for (i = 0; i <= d->dpb_size && i < 16; i++) if (IS_REF(d->buf[i])) {
        d->buf[i].status = PIC_UNUSED;
        if (!d->buf[i].to_be_displayed) d->fullness--;
    }
    while (!output_one(d)) ;
    d->num_ref_frames = 0;
    d->max_long_term_idx = H264_NO_LONG_TERM;
    d->prev_ref_frame_num = 0;
}

static void drop_long_term_idx(h264_dpb_t *d, uint32_t idx)
{
    uint32_t i;
    for (i = 0; i < d->max_ref_frames; i++) if (d->buf[i].status == PIC_LONG && (uint32_t)d->buf[i].pic_num == idx) { set_unused(d, &d->buf[i]); break; }
}

int h264_dpb_mark(h264_dpb_t *d, const h264_slice_hdr_t *sh, int is_ref, int is_idr, int32_t poc, uint32_t pic_id, uint32_t num_err)
{
    h264_dpb_pic_t *cur = &d->buf[d->dpb_size];
    uint32_t frame_num = sh->frame_num, i;
    uint8_t disp = d->no_reordering ? 0 : 1;
    int status = 0;
    d->last_has_mmco5 = 0;
    if (!is_ref) {
        cur->status = PIC_UNUSED; cur->frame_num = frame_num; cur->pic_num = (int32_t)frame_num; cur->poc = poc; cur->to_be_displayed = disp;
        if (!d->no_reordering) d->fullness++;
    } else if (is_idr) {
        d->num_out = d->out_index = 0;
        mmco5(d);
        if (sh->no_output_of_prior_pics || d->no_reordering) { d->num_out = 0; d->out_index = 0; }
        if (sh->long_term_reference_flag) { cur->status = PIC_LONG; d->max_long_term_idx = 0; }
        else { cur->status = PIC_SHORT; d->max_long_term_idx = H264_NO_LONG_TERM; }
        cur->frame_num = 0; cur->pic_num = 0; cur->poc = 0; cur->to_be_displayed = disp;
        d->fullness = 1; d->num_ref_frames = 1;
    } else {
        int marked_long = 0;
        if (sh->adaptive_marking) {
            for (i = 0; i < sh->n_mmco && !status; i++) {
                const h264_mmco_t *m = &sh->mmco[i];
                int idx;
                switch (m->op) {
                case 1:
                    idx = find_pic(d, (int32_t)frame_num - (int32_t)m->diff_pic_nums, 1);
                    if (idx < 0) status = -1; else set_unused(d, &d->buf[idx]);
                    break;
                case 2:
                    idx = find_pic(d, (int32_t)m->long_term_pic_num, 0);
                    if (idx < 0) status = -1; else set_unused(d, &d->buf[idx]);
                    break;
                case 3:
                    if (d->max_long_term_idx == H264_NO_LONG_TERM || m->long_term_frame_idx > d->max_long_term_idx) { status = -1; break; }
                    drop_long_term_idx(d, m->long_term_frame_idx);
                    idx = find_pic(d, (int32_t)frame_num - (int32_t)m->diff_pic_nums, 1);
                    if (idx < 0 || !IS_EXIST(d->buf[idx])) { status = -1; break; }
                    d->buf[idx].status = PIC_LONG; d->buf[idx].pic_num = (int32_t)m->long_term_frame_idx;
                    break;
                case 4: {
                    uint32_t k;
                    d->max_long_term_idx = m->max_long_term_frame_idx;
                    for (k = 0; k < d->max_ref_frames; k++)
                        if (d->buf[k].status == PIC_LONG && ((uint32_t)d->buf[k].pic_num > d->max_long_term_idx || d->max_long_term_idx == H264_NO_LONG_TERM))
                            set_unused(d, &d->buf[k]);
                    break; }
                case 5:
                    mmco5(d); d->last_has_mmco5 = 1; frame_num = 0;
                    break;
                case 6:
                    if (d->max_long_term_idx == H264_NO_LONG_TERM || m->long_term_frame_idx > d->max_long_term_idx) { status = -1; break; }
                    drop_long_term_idx(d, m->long_term_frame_idx);
                    if (d->num_ref_frames < d->max_ref_frames) {
                        cur->frame_num = frame_num; cur->pic_num = (int32_t)m->long_term_frame_idx; cur->poc = poc;
                        cur->status = PIC_LONG; cur->to_be_displayed = disp;
                        d->num_ref_frames++; d->fullness++; marked_long = 1;
                    } else status = -1;
                    break;
                default: status = -1;
                }
            }
        } else status = sliding_window(d);
        if (!marked_long) {
            if (d->num_ref_frames < d->max_ref_frames) {
                cur->frame_num = frame_num; cur->pic_num = (int32_t)frame_num; cur->poc = poc;
                cur->status = PIC_SHORT; cur->to_be_displayed = disp;
                d->fullness++; d->num_ref_frames++;
            } else status = -1;
        }
    }
    cur->is_idr = (uint32_t)is_idr; cur->pic_id = pic_id; cur->num_err_mbs = num_err;
    if (d->no_reordering) {
        d->out[d->num_out].slot = cur->slot; d->out[d->num_out].is_idr = cur->is_idr;
        d->out[d->num_out].pic_id = cur->pic_id; d->out[d->num_out].num_err_mbs = cur->num_err_mbs;
        d->num_out++;
    } else {
        while (d->fullness > d->dpb_size) if (output_one(d)) break;
    }
    sort_buf(d);
    return status;
}

void h264_dpb_flush(h264_dpb_t *d)
{
    if (!d->allocated) return;
    d->flushed = 1;
    while (!output_one(d)) ;
}

const h264_out_t *h264_dpb_next_output(h264_dpb_t *d)
{
    if (d->out_index < d->num_out) return &d->out[d->out_index++];
    return NULL;
}
