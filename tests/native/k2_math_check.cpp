// Checks broadway_b200/csrc/k2_math.cuh (compiled as plain C++) against a scalar statement of the luma sample
// interpolation process, H.264 8.4.2.2.1 (the reference's h264bsd_reconstruct.c:491-1791 implements the same
// formulas position by position).  TEST INFRASTRUCTURE ONLY.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "k2_math.cuh"

static int clip1(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
static int t6(int a, int b, int c, int d, int e, int f) { return a - 5 * b + 20 * c + 20 * d - 5 * e + f; }

// win[r][c]: 9x9 window whose sample (2,2) is the integer sample G of output position (0,0)
struct Ref {
    const uint8_t (*w)[9];
    int G(int x, int y) const { return w[y + 2][x + 2]; }
    int b1(int x, int y) const { return t6(w[y + 2][x], w[y + 2][x + 1], w[y + 2][x + 2], w[y + 2][x + 3], w[y + 2][x + 4], w[y + 2][x + 5]); }   // unclipped
    int h1(int x, int y) const { return t6(w[y][x + 2], w[y + 1][x + 2], w[y + 2][x + 2], w[y + 3][x + 2], w[y + 4][x + 2], w[y + 5][x + 2]); }
    int b(int x, int y) const { return clip1((b1(x, y) + 16) >> 5); }
    int h(int x, int y) const { return clip1((h1(x, y) + 16) >> 5); }
    int j(int x, int y) const { return clip1((t6(b1(x, y - 2), b1(x, y - 1), b1(x, y), b1(x, y + 1), b1(x, y + 2), b1(x, y + 3)) + 512) >> 10); }
    int sample(int x, int y, int fx, int fy) const {
        const int m = h(x + 1, y), s = b(x, y + 1);
        switch (fy * 4 + fx) {
        case 0: return G(x, y);
        case 1: return (G(x, y) + b(x, y) + 1) >> 1;          // a
        case 2: return b(x, y);
        case 3: return (G(x + 1, y) + b(x, y) + 1) >> 1;      // c
        case 4: return (G(x, y) + h(x, y) + 1) >> 1;          // d
        case 5: return (b(x, y) + h(x, y) + 1) >> 1;          // e
        case 6: return (b(x, y) + j(x, y) + 1) >> 1;          // f
        case 7: return (b(x, y) + m + 1) >> 1;                // g
        case 8: return h(x, y);
        case 9: return (h(x, y) + j(x, y) + 1) >> 1;          // i
        case 10: return j(x, y);
        case 11: return (j(x, y) + m + 1) >> 1;               // k
        case 12: return (G(x, y + 1) + h(x, y) + 1) >> 1;     // n
        case 13: return (h(x, y) + s + 1) >> 1;               // p
        case 14: return (j(x, y) + s + 1) >> 1;               // q
        default: return (m + s + 1) >> 1;                     // r
        }
    }
};

static uint32_t rng = 12345;
static uint32_t rnd() { rng = rng * 1664525u + 1013904223u; return rng >> 8; }

int main()
{
    long checked = 0, bad = 0;
    for (int iter = 0; iter < 60000; iter++) {
        uint8_t w[9][9];
        const int kind = iter % 6;
        for (int r = 0; r < 9; r++) for (int c = 0; c < 9; c++) {
            switch (kind) {
            case 0: w[r][c] = (uint8_t)rnd(); break;
            case 1: w[r][c] = (rnd() & 1) ? 255 : 0; break;                          // overshoot / undershoot of every filter
            case 2: w[r][c] = ((r + c) & 1) ? 255 : 0; break;
            case 3: w[r][c] = (iter & 64) ? 255 : 0; break;                          // flat extremes
            case 4: w[r][c] = (uint8_t)(((c % 3) == (iter % 3)) ? 255 : (rnd() & 3)); break;
            default: w[r][c] = (uint8_t)(128 + (int)(rnd() % 7) - 3); break;        // near-flat
            }
        }
        uint32_t r0[9], r1[9], r2[9];
        for (int r = 0; r < 9; r++) {
            r0[r] = w[r][0] | (w[r][1] << 8) | (w[r][2] << 16) | ((uint32_t)w[r][3] << 24);
            r1[r] = w[r][4] | (w[r][5] << 8) | (w[r][6] << 16) | ((uint32_t)w[r][7] << 24);
            r2[r] = w[r][8];
        }
        Ref ref; ref.w = w;
        for (int fy = 0; fy < 4; fy++) for (int fx = 0; fx < 4; fx++) {
            // operand selection exactly as k2_inter.cuh derives it
            const bool useJ = (fx == 2 && fy != 0) || (fy == 2 && fx != 0);
            const bool useB = fx != 0 && fy != 2;
            const bool useH = fy != 0 && fx != 2;
            const bool useG = !useJ && !(useB && useH) && !(fx == 2) && !(fy == 2);
            const int n_ops = (int)useJ + (int)useB + (int)useH + (int)useG;
            // the warp-level "any" flags are supersets of the block's own: every superset must give the same block
            for (int extra = 0; extra < 8; extra++) {
                const bool anyB = useB || (extra & 1), anyH = useH || (extra & 2), anyJ = useJ || (extra & 4);
                uint32_t out[4];
                k2m_luma4x4(r0, r1, r2, fx, fy, useG, useB, useH, useJ, n_ops, anyB, anyH, anyJ, out);
                for (int y = 0; y < 4; y++) for (int x = 0; x < 4; x++) {
                    const int got = (out[y] >> (8 * x)) & 0xff, want = ref.sample(x, y, fx, fy);
                    checked++;
                    if (got != want) { if (bad < 10) printf("MISMATCH kind %d frac (%d,%d) extra %d at (%d,%d): got %d want %d\n", kind, fx, fy, extra, x, y, got, want); bad++; }
                }
            }
        }
    }
    printf("%ld samples checked, %ld mismatches\n", checked, bad);
    return bad ? 1 : 0;
}
