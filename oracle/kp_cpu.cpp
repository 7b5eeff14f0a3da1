/* oracle/kp_cpu.cpp — TEST INFRASTRUCTURE ONLY (never linked into the product).
 * The device-side slice parser (broadway_b200/csrc/kp_core.h, the body of CUDA kernel Kp) compiled as plain C++
 * with a warp of one lane, so that the CPU test-suite can run the device-parse path end to end — host decoder in
 * device-parse mode -> picture blocks (include/h264b200_slices.h) -> kp_core -> records -> recon_cpu.c — and compare
 * the records with the host parser's and the frames with the reference goldens, without a GPU. */
#include <stdlib.h>
#include "../broadway_b200/csrc/kp_core.h"

extern "C" void kp_cpu_parse_picture(const KpPic *pic, const KpTables *tables)
{
    KpStage *st = (KpStage *)aligned_alloc(16, (sizeof(KpStage) + 15) & ~(size_t)15);
    memset(st, 0, sizeof *st);
    kp_parse_picture(0, *pic, st, tables);
    free(st);
}
