/* oracle/ref_tap.h — TEST INFRASTRUCTURE ONLY (see ref_tap.c). */
#ifndef ORACLE_REF_TAP_H
#define ORACLE_REF_TAP_H
#include <stddef.h>

/* Plain copy of the fields of the reference's mbStorage_t
 * (h264bsd_macroblock_layer.h:166-188) that the parity tests compare. */
typedef struct {
    int mb_type;
    unsigned slice_id;
    unsigned qp_y;
    unsigned disable_deblock_idc;
    int filter_offset_a;
    int filter_offset_b;
    int chroma_qp_index_offset;
    short total_coeff[27];
    unsigned char intra4x4_mode[16];
    short mv[16][2];
    unsigned ref_pic[4];
} reftap_mb_t;

void reftap_set_predeblock_buffer(unsigned char *buf, size_t cap);
void reftap_set_mb_buffer(reftap_mb_t *buf, size_t cap_mbs);
size_t reftap_predeblock_len(void);
size_t reftap_mb_len(void);
unsigned reftap_pictures_filtered(void);

#endif
