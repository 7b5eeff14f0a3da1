/* tests/native/dpb_spare_check.c — the spare frame slot of the DPB (h264_dpb_rotate_spare, csrc/h264_dpb.c).
 * Drives the DPB the way h264_decoder.c does (rotate at picture begin, current slot at picture end, mark) through IPPP
 * sequences with 1..4 reference frames and checks the two properties the engine relies on:
 *   1. the slot a picture is decoded into is never a slot the DPB still holds as a reference or for output, and
 *   2. it is never the slot of the picture decoded just before, nor the one before that when that one has been freed in
 *      between: a freed slot rests for (at least) one picture before it is decoded into again. */
#include <stdio.h>
#include <string.h>
#include "h264_internal.h"

static int run(uint32_t num_ref, int no_reorder, int n_pics)
{
    h264_dpb_t d; h264_slice_hdr_t sh; int prev1 = -1, prev2 = -1, i, k, bad = 0;
    uint32_t max_slot = 0;
    memset(&sh, 0, sizeof sh);
    h264_dpb_init(&d, num_ref, num_ref, 16, no_reorder);
    for (i = 0; i < n_pics; i++) {
        int cur;
        sh.frame_num = (uint32_t)i % 16;
        if (i && h264_dpb_check_gaps(&d, sh.frame_num, 1, 0)) return 100;
        h264_dpb_rotate_spare(&d);
        cur = h264_dpb_current_slot(&d);
        if ((uint32_t)cur > max_slot) max_slot = (uint32_t)cur;
        for (k = 0; k < (int)d.dpb_size; k++)                      /* property 1 */
            if (d.buf[k].slot == cur && (d.buf[k].status != PIC_UNUSED || d.buf[k].to_be_displayed)) bad |= 1;
        if (cur == prev1) bad |= 2;                               /* property 2 */
        if (cur == prev2 && num_ref == 1) bad |= 4;               /* with one reference frame the slots go round in threes */
        if (h264_dpb_mark(&d, &sh, 1, i == 0, 2 * i, (uint32_t)i, 0)) return 101;
        while (h264_dpb_next_output(&d)) ;
        prev2 = prev1; prev1 = cur;
    }
    if (max_slot > d.dpb_size + 1) bad |= 8;                      /* n_slots = dpb_size + 2 (h264_decoder.c) */
    return bad;
}

int main(void)
{
    uint32_t r; int nr, rc = 0;
    for (r = 1; r <= 4; r++) for (nr = 0; nr < 2; nr++) {
        int b = run(r, nr, 40);
        if (b) { printf("num_ref %u no_reorder %d: violation mask %d\n", r, nr, b); rc = 1; }
    }
    if (!rc) printf("ok\n");
    return rc;
}
