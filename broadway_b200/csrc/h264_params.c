/* h264_params.c — sequence/picture parameter sets, VUI subset, slice header and
 * picture order count (ITU-T H.264 7.3.2.1, 7.3.2.2, E.1.1, 7.3.3, 8.2.1) with
 * the Baseline restrictions the reference enforces: frame_mbs_only
 * (h264bsd_seq_param_set.c:250-258), CAVLC only (h264bsd_pic_param_set.c:
 * 125-131), no weighted prediction (:264-270), I and P slices only
 * (h264bsd_slice_header.c:135-144).  Return 0 on success, <0 on error. */
#include <string.h>
#include <stdlib.h>
#include "h264_internal.h"

#define CHECK_UE(v) do { if ((v) == 0xffffffffu) return -1; } while (0)

/* MaxDPB (bytes) and max frame size (MBs) per level_idc (Table A-1) */
static uint32_t dpb_size_for_level(uint32_t pic_size_mbs, uint32_t level_idc)
{
    uint32_t bytes, max_mbs;
    switch (level_idc) {
    case 10: bytes = 152064; max_mbs = 99; break;
    case 11: bytes = 345600; max_mbs = 396; break;
    case 12: case 13: case 20: bytes = 912384; max_mbs = 396; break;
    case 21: bytes = 1824768; max_mbs = 792; break;
    case 22: case 30: bytes = 3110400; max_mbs = 1620; break;
    case 31: bytes = 6912000; max_mbs = 3600; break;
    case 32: bytes = 7864320; max_mbs = 5120; break;
    case 40: case 41: bytes = 12582912; max_mbs = 8192; break;
    case 42: bytes = 34816u * 384u; max_mbs = 8704; break;
    case 50: bytes = 42393600; max_mbs = 22080; break;
    case 51: bytes = 70778880; max_mbs = 36864; break;
    default: return 0xffffffffu;
    }
    if (pic_size_mbs > max_mbs) return 0xffffffffu;
    bytes /= pic_size_mbs * 384u;
    return bytes < 16 ? bytes : 16;
}

static int parse_hrd(br_t *b)
{
    uint32_t cnt = br_ue(b), i;
    CHECK_UE(cnt);
    if (cnt > 31) return -1;
    br_get(b, 8);                              /* bit_rate_scale, cpb_size_scale */
    for (i = 0; i <= cnt; i++) { CHECK_UE(br_ue(b)); CHECK_UE(br_ue(b)); br_get1(b); }
    br_get(b, 20);                             /* four 5-bit length fields */
    return 0;
}

static int parse_vui(br_t *b, h264_sps_t *s)
{
    int nal_hrd, vcl_hrd;
    s->aspect_ratio_present = (uint8_t)br_get1(b);
    if (s->aspect_ratio_present) {
        s->aspect_ratio_idc = (uint8_t)br_get(b, 8);
        if (s->aspect_ratio_idc == 255) { s->sar_width = br_get(b, 16); s->sar_height = br_get(b, 16); }
    }
    if (br_get1(b)) br_get1(b);                /* overscan */
    s->video_signal_present = (uint8_t)br_get1(b);
    s->matrix_coefficients = 2;
    if (s->video_signal_present) {
        br_get(b, 3);
        s->video_full_range = (uint8_t)br_get1(b);
        s->colour_desc_present = (uint8_t)br_get1(b);
        if (s->colour_desc_present) { br_get(b, 16); s->matrix_coefficients = (uint8_t)br_get(b, 8); }
    }
    if (br_get1(b)) { CHECK_UE(br_ue(b)); CHECK_UE(br_ue(b)); }      /* chroma loc */
    if (br_get1(b)) { br_get(b, 32); br_get(b, 32); br_get1(b); }    /* timing */
    nal_hrd = (int)br_get1(b);
    if (nal_hrd && parse_hrd(b)) return -1;
    vcl_hrd = (int)br_get1(b);
    if (vcl_hrd && parse_hrd(b)) return -1;
    if (nal_hrd || vcl_hrd) br_get1(b);
    br_get1(b);                                /* pic_struct_present_flag */
    s->bitstream_restriction = (uint8_t)br_get1(b);
    if (s->bitstream_restriction) {
        uint32_t v;
        br_get1(b);
        CHECK_UE(br_ue(b)); CHECK_UE(br_ue(b)); CHECK_UE(br_ue(b)); CHECK_UE(br_ue(b));
        v = br_ue(b); CHECK_UE(v); s->num_reorder_frames = v;
        v = br_ue(b); CHECK_UE(v); s->max_dec_frame_buffering = v;
    }
    return br_overrun(b) ? -1 : 0;
}

int h264_parse_sps(br_t *b, h264_sps_t *s)
{
    uint32_t v, i;
    memset(s, 0, sizeof *s);
    s->profile_idc = (uint8_t)br_get(b, 8);    /* non-Baseline profiles are attempted, as in the reference */
    br_get(b, 8);                              /* constraint flags + reserved */
    s->level_idc = (uint8_t)br_get(b, 8);
    v = br_ue(b); CHECK_UE(v); if (v >= H264_MAX_SPS) return -1; s->sps_id = (uint8_t)v;
    v = br_ue(b); CHECK_UE(v); if (v > 12) return -1;
    s->log2_max_frame_num = (uint8_t)(v + 4); s->max_frame_num = 1u << (v + 4);
    v = br_ue(b); CHECK_UE(v); if (v > 2) return -1; s->poc_type = (uint8_t)v;
    if (s->poc_type == 0) {
        v = br_ue(b); CHECK_UE(v); if (v > 12) return -1;
        s->log2_max_poc_lsb = (uint8_t)(v + 4); s->max_poc_lsb = 1u << (v + 4);
    } else if (s->poc_type == 1) {
        int32_t sv;
        s->delta_pic_order_always_zero = (uint8_t)br_get1(b);
        sv = br_se(b); if (sv == INT32_MIN) return -1; s->offset_for_non_ref_pic = sv;
        sv = br_se(b); if (sv == INT32_MIN) return -1; s->offset_for_top_to_bottom = sv;
        v = br_ue(b); CHECK_UE(v); if (v > 255) return -1; s->num_ref_frames_in_poc_cycle = v;
        for (i = 0; i < v; i++) { sv = br_se(b); if (sv == INT32_MIN) return -1; s->offset_for_ref_frame[i] = sv; }
    }
    v = br_ue(b); CHECK_UE(v); if (v > H264_MAX_REFS) return -1; s->num_ref_frames = v;
    s->gaps_allowed = (uint8_t)br_get1(b);
    v = br_ue(b); CHECK_UE(v); s->width_mbs = v + 1;
    v = br_ue(b); CHECK_UE(v); s->height_mbs = v + 1;
    if (s->width_mbs > 1024 || s->height_mbs > 1024) return -1;
    if (!br_get1(b)) return -1;                /* frame_mbs_only_flag must be 1 */
    br_get1(b);                                /* direct_8x8_inference_flag */
    s->crop_flag = (uint8_t)br_get1(b);
    if (s->crop_flag) {
        v = br_ue(b); CHECK_UE(v); s->crop_left = v;
        v = br_ue(b); CHECK_UE(v); s->crop_right = v;
        v = br_ue(b); CHECK_UE(v); s->crop_top = v;
        v = br_ue(b); CHECK_UE(v); s->crop_bottom = v;
        if ((int64_t)s->crop_left > 8 * (int64_t)s->width_mbs - ((int64_t)s->crop_right + 1) ||
            (int64_t)s->crop_top > 8 * (int64_t)s->height_mbs - ((int64_t)s->crop_bottom + 1)) return -1;
    }
    v = dpb_size_for_level(s->width_mbs * s->height_mbs, s->level_idc);
    if (v == 0xffffffffu || s->num_ref_frames > v) v = s->num_ref_frames;   /* h264bsd_seq_param_set.c:302-313 */
    s->max_dpb_size = v;
    s->vui_present = (uint8_t)br_get1(b);
    s->matrix_coefficients = 2;
    if (s->vui_present) {
        if (parse_vui(b, s)) return -1;
        if (s->bitstream_restriction) {
            if (s->num_reorder_frames > s->max_dec_frame_buffering || s->max_dec_frame_buffering < s->num_ref_frames ||
                s->max_dec_frame_buffering > s->max_dpb_size) return -1;
            s->max_dpb_size = s->max_dec_frame_buffering ? s->max_dec_frame_buffering : 1;
        }
    }
    if (br_overrun(b)) return -1;
    s->valid = 1;
    return 0;
}

int h264_parse_pps(br_t *b, h264_pps_t *p)
{
    uint32_t v; int32_t sv;
    memset(p, 0, sizeof *p);
    v = br_ue(b); CHECK_UE(v); if (v >= H264_MAX_PPS) return -1; p->pps_id = (uint8_t)v;
    v = br_ue(b); CHECK_UE(v); if (v >= H264_MAX_SPS) return -1; p->sps_id = (uint8_t)v;
    if (br_get1(b)) return -1;                 /* entropy_coding_mode_flag: CABAC is not Baseline */
    p->pic_order_present = (uint8_t)br_get1(b);
    v = br_ue(b); CHECK_UE(v); if (v > 7) return -1; p->num_slice_groups = v + 1;
    if (p->num_slice_groups > 1) {             /* flexible macroblock ordering (h264bsd_pic_param_set.c:150-262) */
        h264_fmo_t *f = &p->fmo;
        uint32_t i;
        f->n_groups = p->num_slice_groups;
        v = br_ue(b); CHECK_UE(v); if (v > 6) return -1; f->type = v;
        if (f->type == 0) {
            for (i = 0; i < f->n_groups; i++) { v = br_ue(b); CHECK_UE(v); f->run_length[i] = v + 1; }
        } else if (f->type == 2) {
            for (i = 0; i + 1 < f->n_groups; i++) { v = br_ue(b); CHECK_UE(v); f->top_left[i] = v; v = br_ue(b); CHECK_UE(v); f->bottom_right[i] = v; }
        } else if (f->type >= 3 && f->type <= 5) {
            f->change_direction = br_get1(b);
            v = br_ue(b); CHECK_UE(v); f->change_rate = v + 1;
        } else if (f->type == 6) {
            uint32_t bits = 0, n;
            uint8_t *ids;
            v = br_ue(b); CHECK_UE(v); if (v >= 36864u) return -1;
            n = v + 1; p->fmo_map_units = n;
            while ((1u << bits) < f->n_groups) bits++;
            ids = (uint8_t *)h264_malloc(n);
            if (!ids) return -1;
            for (i = 0; i < n; i++) { ids[i] = (uint8_t)br_get(b, (int)bits); if (ids[i] >= f->n_groups) { h264_free(ids); return -1; } }
            f->group_id = ids;
        }
    }
    v = br_ue(b); CHECK_UE(v); if (v > 31) return -1; p->num_ref_idx_l0_default = v + 1;
    v = br_ue(b); CHECK_UE(v); if (v > 31) return -1;
    if (br_get1(b)) return -1;                 /* weighted_pred_flag */
    if (br_get(b, 2) > 2) return -1;           /* weighted_bipred_idc */
    sv = br_se(b); if (sv < -26 || sv > 25) return -1; p->pic_init_qp = sv + 26;
    sv = br_se(b); if (sv < -26 || sv > 25) return -1;
    sv = br_se(b); if (sv < -12 || sv > 12) return -1; p->chroma_qp_index_offset = sv;
    p->deblocking_control_present = (uint8_t)br_get1(b);
    p->constrained_intra_pred = (uint8_t)br_get1(b);
    p->redundant_pic_cnt_present = (uint8_t)br_get1(b);
    if (br_overrun(b)) return -1;
    p->valid = 1;
    return 0;
}

int h264_peek_pps_id(br_t b, uint32_t *pps_id)
{
    uint32_t v;
    v = br_ue(&b); CHECK_UE(v);               /* first_mb_in_slice */
    v = br_ue(&b); CHECK_UE(v);               /* slice_type */
    v = br_ue(&b); CHECK_UE(v);
    if (v >= H264_MAX_PPS) return -1;
    *pps_id = v;
    return 0;
}

int h264_parse_slice_header(br_t *b, h264_slice_hdr_t *sh, const h264_sps_t *sps, const h264_pps_t *pps, int nal_type, int nal_ref_idc)
{
    uint32_t v; int32_t sv;
    int idr = nal_type == NAL_IDR;
    memset(sh, 0, sizeof *sh);
    v = br_ue(b); CHECK_UE(v); if (v >= sps->width_mbs * sps->height_mbs) return -1; sh->first_mb = v;
    v = br_ue(b); CHECK_UE(v); if (v > 9) return -1;
    sh->slice_type = (uint8_t)(v % 5);
    if (sh->slice_type != 2 && (sh->slice_type != 0 || idr || !sps->num_ref_frames)) return -1;
    v = br_ue(b); CHECK_UE(v); if (v != pps->pps_id) return -1; sh->pps_id = v;
    sh->frame_num = br_get(b, sps->log2_max_frame_num);
    if (idr && sh->frame_num != 0) return -1;
    if (idr) { v = br_ue(b); CHECK_UE(v); if (v > 65535) return -1; sh->idr_pic_id = v; }
    if (sps->poc_type == 0) {
        sh->poc_lsb = br_get(b, sps->log2_max_poc_lsb);
        if (pps->pic_order_present) { sv = br_se(b); if (sv == INT32_MIN) return -1; sh->delta_poc_bottom = sv; }
        if (idr) {
            int32_t bot = (int32_t)sh->poc_lsb + sh->delta_poc_bottom;
            if (sh->poc_lsb > sps->max_poc_lsb / 2 || ((int32_t)sh->poc_lsb < bot ? (int32_t)sh->poc_lsb : bot) != 0) return -1;
        }
    } else if (sps->poc_type == 1 && !sps->delta_pic_order_always_zero) {
        sv = br_se(b); if (sv == INT32_MIN) return -1; sh->delta_poc[0] = sv;
        if (pps->pic_order_present) { sv = br_se(b); if (sv == INT32_MIN) return -1; sh->delta_poc[1] = sv; }
        if (idr) {
            int32_t bot = sh->delta_poc[0] + sps->offset_for_top_to_bottom + sh->delta_poc[1];
            if ((sh->delta_poc[0] < bot ? sh->delta_poc[0] : bot) != 0) return -1;
        }
    }
    if (pps->redundant_pic_cnt_present) { v = br_ue(b); CHECK_UE(v); if (v > 127) return -1; sh->redundant_pic_cnt = v; }
    if (sh->slice_type == 0) {
        if (br_get1(b)) { v = br_ue(b); CHECK_UE(v); if (v > 15) return -1; sh->num_ref_idx_active = v + 1; }
        else { if (pps->num_ref_idx_l0_default > 16) return -1; sh->num_ref_idx_active = pps->num_ref_idx_l0_default; }
        /* ref_pic_list_reordering (7.3.3.1) */
        sh->reorder_flag = (uint8_t)br_get1(b);
        if (sh->reorder_flag) {
            for (;;) {
                uint32_t idc = br_ue(b); CHECK_UE(idc);
                if (idc > 3) return -1;
                if (idc == 3) break;
                if (sh->n_reorder > sh->num_ref_idx_active) return -1;
                v = br_ue(b); CHECK_UE(v);
                if (idc < 2) { if (v >= sps->max_frame_num) return -1; v += 1; }
                sh->reorder[sh->n_reorder].idc = (uint8_t)idc; sh->reorder[sh->n_reorder].val = v; sh->n_reorder++;
            }
            if (sh->n_reorder == 0) return -1;
        }
    }
    if (nal_ref_idc != 0) {                     /* dec_ref_pic_marking (7.3.3.3) */
        if (idr) { sh->no_output_of_prior_pics = (uint8_t)br_get1(b); sh->long_term_reference_flag = (uint8_t)br_get1(b); }
        else {
            sh->adaptive_marking = (uint8_t)br_get1(b);
            if (sh->adaptive_marking) {
                uint32_t n4 = 0, n5 = 0, n6 = 0, n123 = 0;
                for (;;) {
                    h264_mmco_t *m;
                    uint32_t op = br_ue(b); CHECK_UE(op);
                    if (op > 6) return -1;
                    if (op == 0) break;
                    if (sh->n_mmco >= 35) return -1;
                    m = &sh->mmco[sh->n_mmco++];
                    m->op = (uint8_t)op;
                    if (op == 1 || op == 3) { v = br_ue(b); CHECK_UE(v); m->diff_pic_nums = v + 1; }
                    if (op == 2) { v = br_ue(b); CHECK_UE(v); m->long_term_pic_num = v; }
                    if (op == 3 || op == 6) { v = br_ue(b); CHECK_UE(v); m->long_term_frame_idx = v; }
                    if (op == 4) {
                        v = br_ue(b); CHECK_UE(v);
                        if (v > sps->num_ref_frames) return -1;
                        m->max_long_term_frame_idx = v ? v - 1 : H264_NO_LONG_TERM;
                        n4++;
                    }
                    if (op == 5) n5++;
                    if (op == 6) n6++;
                    if (op >= 1 && op <= 3) n123++;
                }
                /* at most one each of 4, 5, 6; 5 excludes 1..3 (h264bsd_slice_header.c DecRefPicMarking) */
                if (n4 > 1 || n5 > 1 || n6 > 1 || (n123 && n5)) return -1;
            }
        }
    }
    sv = br_se(b); if (sv == INT32_MIN) return -1;
    sh->slice_qp = pps->pic_init_qp + sv;
    if (sh->slice_qp < 0 || sh->slice_qp > 51) return -1;
    if (pps->deblocking_control_present) {
        v = br_ue(b); CHECK_UE(v); if (v > 2) return -1; sh->disable_deblocking_idc = (uint8_t)v;
        if (v != 1) {
            sv = br_se(b); if (sv < -6 || sv > 6) return -1; sh->alpha_off = (int8_t)(sv * 2);
            sv = br_se(b); if (sv < -6 || sv > 6) return -1; sh->beta_off = (int8_t)(sv * 2);
        }
    }
    if (pps->num_slice_groups > 1 && pps->fmo.type >= 3 && pps->fmo.type <= 5) {
        const uint32_t size = sps->width_mbs * sps->height_mbs, rate = pps->fmo.change_rate;
        sh->slice_group_change_cycle = br_get(b, (int)h264_fmo_cycle_bits(size, rate));
        if (sh->slice_group_change_cycle > (size + rate - 1) / rate) return -1;      /* h264bsd_slice_header.c:371-380 */
    }
    return br_overrun(b) ? -1 : 0;
}

/* 8.2.1, frames only; mirrors the reference's handling of mmco5 (h264bsd_pic_order_cnt.c:77-350) */
int32_t h264_decode_poc(h264_decoder_t *d, const h264_slice_hdr_t *sh, int nal_type, int nal_ref_idc)
{
    const h264_sps_t *sps = d->active_sps;
    int idr = nal_type == NAL_IDR, mmco5 = 0;
    int32_t poc = 0;
    uint32_t i, frame_num_offset;
    if (sh->adaptive_marking) for (i = 0; i < sh->n_mmco; i++) if (sh->mmco[i].op == 5) mmco5 = 1;
    if (sps->poc_type == 0) {
        int32_t msb;
        if (idr) { d->poc.prev_poc_msb = 0; d->poc.prev_poc_lsb = 0; }
        if (sh->poc_lsb < d->poc.prev_poc_lsb && d->poc.prev_poc_lsb - sh->poc_lsb >= sps->max_poc_lsb / 2)
            msb = d->poc.prev_poc_msb + (int32_t)sps->max_poc_lsb;
        else if (sh->poc_lsb > d->poc.prev_poc_lsb && sh->poc_lsb - d->poc.prev_poc_lsb > sps->max_poc_lsb / 2)
            msb = d->poc.prev_poc_msb - (int32_t)sps->max_poc_lsb;
        else msb = d->poc.prev_poc_msb;
        if (nal_ref_idc) d->poc.prev_poc_msb = msb;
        poc = msb + (int32_t)sh->poc_lsb;
        if (sh->delta_poc_bottom < 0) poc += sh->delta_poc_bottom;
        if (nal_ref_idc) {
            if (mmco5) {
                d->poc.prev_poc_msb = 0;
                d->poc.prev_poc_lsb = sh->delta_poc_bottom < 0 ? (uint32_t)(-sh->delta_poc_bottom) : 0;
                poc = 0;
            } else d->poc.prev_poc_lsb = sh->poc_lsb;
        }
        return poc;
    }
    if (idr) frame_num_offset = 0;
    else if (d->poc.prev_frame_num > sh->frame_num) frame_num_offset = d->poc.prev_frame_num_offset + sps->max_frame_num;
    else frame_num_offset = d->poc.prev_frame_num_offset;
    if (sps->poc_type == 1) {
        uint32_t abs_frame_num = sps->num_ref_frames_in_poc_cycle ? frame_num_offset + sh->frame_num : 0;
        int32_t cycle_delta = 0;
        if (nal_ref_idc == 0 && abs_frame_num > 0) abs_frame_num--;
        for (i = 0; i < sps->num_ref_frames_in_poc_cycle; i++) cycle_delta += sps->offset_for_ref_frame[i];
        if (abs_frame_num > 0) {
            uint32_t cnt = (abs_frame_num - 1) / sps->num_ref_frames_in_poc_cycle;
            uint32_t in_cycle = (abs_frame_num - 1) % sps->num_ref_frames_in_poc_cycle;
            poc = (int32_t)cnt * cycle_delta;
            for (i = 0; i <= in_cycle; i++) poc += sps->offset_for_ref_frame[i];
        }
        if (nal_ref_idc == 0) poc += sps->offset_for_non_ref_pic;
        poc += sh->delta_poc[0];
        if (sps->offset_for_top_to_bottom + sh->delta_poc[1] < 0) poc += sps->offset_for_top_to_bottom + sh->delta_poc[1];
    } else {
        if (idr) poc = 0;
        else if (nal_ref_idc == 0) poc = 2 * (int32_t)(frame_num_offset + sh->frame_num) - 1;
        else poc = 2 * (int32_t)(frame_num_offset + sh->frame_num);
    }
    if (!mmco5) { d->poc.prev_frame_num_offset = frame_num_offset; d->poc.prev_frame_num = sh->frame_num; }
    else { d->poc.prev_frame_num_offset = 0; d->poc.prev_frame_num = 0; poc = 0; }
    return poc;
}
