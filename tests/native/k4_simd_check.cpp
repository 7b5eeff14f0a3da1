// Host check of broadway_b200/csrc/k4_simd.cuh (built and run by tests/test_k4_simd_cpu.py).
// A scalar statement of the edge filter (8.7.2.3 / 8.7.2.4; h264bsd_deblocking.c:649-1121) against the
// two-lines-per-register version, on random samples, extreme samples, every bS and every table index.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "h264_consts.h"
#include "k4_simd.cuh"

static int clip3(int lo, int hi, int v) { return v < lo ? lo : v > hi ? hi : v; }
static void scalar_edge(int *v, int bs, int alpha, int beta, const uint8_t *tc0tab, bool luma)
{
    const int p3 = v[0], p2 = v[1], p1 = v[2], p0 = v[3], q0 = v[4], q1 = v[5], q2 = v[6], q3 = v[7];
    if (!bs) return;
    if (!(abs(p0 - q0) < alpha && abs(p1 - p0) < beta && abs(q1 - q0) < beta)) return;
    const bool ap = luma && abs(p2 - p0) < beta, aq = luma && abs(q2 - q0) < beta;
    if (bs < 4) {
        const int tc0 = tc0tab[bs - 1], tc = luma ? tc0 + ap + aq : tc0 + 1;
        const int d = clip3(-tc, tc, (((q0 - p0) << 2) + (p1 - q1) + 4) >> 3);
        v[3] = clip3(0, 255, p0 + d); v[4] = clip3(0, 255, q0 - d);
        if (ap) v[2] = p1 + clip3(-tc0, tc0, (p2 + ((p0 + q0 + 1) >> 1) - (p1 << 1)) >> 1);
        if (aq) v[5] = q1 + clip3(-tc0, tc0, (q2 + ((p0 + q0 + 1) >> 1) - (q1 << 1)) >> 1);
    } else {
        const bool small = abs(p0 - q0) < ((alpha >> 2) + 2);
        if (ap && small) { v[3] = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3; v[2] = (p2 + p1 + p0 + q0 + 2) >> 2; v[1] = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3; }
        else v[3] = (2 * p1 + p0 + q1 + 2) >> 2;
        if (aq && small) { v[4] = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3; v[5] = (p0 + q0 + q1 + q2 + 2) >> 2; v[6] = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3; }
        else v[4] = (2 * q1 + q0 + p1 + 2) >> 2;
    }
}

static uint32_t rng = 12345;
static uint32_t rnd() { rng = rng * 1664525u + 1013904223u; return rng >> 8; }

int main()
{
    long n = 0, bad = 0, changed = 0;
    for (int ia = 0; ia < 52; ia++) for (int ib = 0; ib < 52; ib += 3) for (int bs = 0; bs <= 4; bs++) for (int lu = 0; lu < 2; lu++) for (int rep = 0; rep < 120; rep++) {
        const int alpha = H264_ALPHA[ia], beta = H264_BETA[ib];
        const uint32_t thr = (uint32_t)alpha | ((uint32_t)beta << 8);
        const uint32_t tcw = (uint32_t)H264_TC0[ia][0] | ((uint32_t)H264_TC0[ia][1] << 8) | ((uint32_t)H264_TC0[ia][2] << 16);
        int a[8], b[8];
        const int mode = rep % 6, base = rnd() & 255, spread = mode < 2 ? 4 : mode < 4 ? 24 : 255;
        for (int i = 0; i < 8; i++) {
            if (mode == 5) { a[i] = (rnd() & 1) ? 255 : 0; b[i] = (rnd() & 1) ? 255 : 0; }
            else { a[i] = clip3(0, 255, base + (int)(rnd() % (2 * spread + 1)) - spread); b[i] = clip3(0, 255, base + (int)(rnd() % (2 * spread + 1)) - spread); }
        }
        uint32_t v[8];
        for (int i = 0; i < 8; i++) v[i] = (uint32_t)a[i] | ((uint32_t)b[i] << 16);
        for (int variant = 0; variant < 3; variant++) {           /* hints: exact, both on */
            uint32_t w[8]; for (int i = 0; i < 8; i++) w[i] = v[i];
            const bool aw = variant == 0 ? (bs > 0 && bs < 4) : true, as = variant == 0 ? bs == 4 : variant == 1;
            if (variant == 2 && bs == 4) continue;                /* strong lanes always come with any_strong */
            dbk_edge2(w, bs, thr, tcw, lu != 0, aw, as);
            int ra[8], rb[8]; for (int i = 0; i < 8; i++) { ra[i] = a[i]; rb[i] = b[i]; }
            scalar_edge(ra, bs, alpha, beta, H264_TC0[ia], lu != 0); scalar_edge(rb, bs, alpha, beta, H264_TC0[ia], lu != 0);
            for (int i = 0; i < 8; i++) { n++; changed += ((uint32_t)ra[i] | ((uint32_t)rb[i] << 16)) != v[i]; if (w[i] != ((uint32_t)ra[i] | ((uint32_t)rb[i] << 16))) { if (bad++ < 10) printf("MISMATCH ia %d ib %d bs %d luma %d i %d: got %08x want %04x%04x\n", ia, ib, bs, lu, i, w[i], rb[i], ra[i]); } }
        }
    }
    /* byte <-> two-line register shuffles */
    for (int rep = 0; rep < 1000; rep++) {
        uint32_t a = rnd() * 2654435761u, b = rnd() * 40503u + rnd(), o[4], a2, b2;
        k4s_unpack_rows(a, b, o);
        for (int j = 0; j < 4; j++) { n++; if (o[j] != (((a >> (8 * j)) & 0xff) | (((b >> (8 * j)) & 0xff) << 16))) bad++; }
        k4s_pack_rows(o, &a2, &b2); n += 2; if (a2 != a) bad++; if (b2 != b) bad++;
        uint32_t h = rnd() & 0xffff; n += 2;
        if (k4s_unpack_pair(h) != ((h & 0xff) | ((h >> 8) << 16))) bad++;
        if (k4s_pack_pair(k4s_unpack_pair(h)) != h) bad++;
    }
    printf("%ld checks, %ld mismatches, %ld registers changed by the filter\n", n, bad, changed);
    if (changed < n / 50) { printf("too few filtered cases\n"); return 2; }
    return bad ? 1 : 0;
}
