/* h264_bits.h — MSB-first bit reader over an RBSP (emulation prevention already
 * removed) with a 64-bit cache, Exp-Golomb ue/se/te and more_rbsp_data().
 * Role of the reference's h264bsd_stream.c:72-242 and h264bsd_vlc.c:103-391,
 * redesigned: the reference re-assembles 32 bits from bytes on every
 * h264bsdShowBits32 call; here one unaligned 64-bit load refills the cache. */
#ifndef B200_H264_BITS_H
#define B200_H264_BITS_H
#include <stdint.h>
#include <stddef.h>
#include <string.h>

typedef struct {
    const uint8_t *data;   /* RBSP bytes */
    size_t   len;
    size_t   pos;          /* next byte to load; may run past len (virtual zero bytes) */
    uint64_t cache;        /* next bits, MSB first */
    int      bits;         /* valid bits in cache */
    uint64_t payload_bits; /* bits before the rbsp_stop_one_bit */
} br_t;

static inline void br_refill(br_t *b)
{
    if (b->pos + 8 <= b->len) {
        uint64_t w;
        int n;
        memcpy(&w, b->data + b->pos, 8);
        w = __builtin_bswap64(w);
        b->cache |= (w >> b->bits);
        n = (63 - b->bits) >> 3;
        b->pos += (size_t)n;
        b->bits += n * 8;
    } else {
        /* tail (and past the end: zeros; consumers detect overrun through br_overrun) */
        while (b->bits <= 56) {
            uint64_t v = b->pos < b->len ? b->data[b->pos] : 0;
            b->cache |= v << (56 - b->bits);
            b->pos++;
            b->bits += 8;
        }
    }
}

static inline void br_init(br_t *b, const uint8_t *data, size_t len)
{
    size_t n = len;
    b->data = data; b->len = len; b->pos = 0; b->cache = 0; b->bits = 0;
    /* locate rbsp_stop_one_bit: last set bit of the last non-zero byte */
    while (n > 0 && data[n - 1] == 0) n--;
    if (n == 0) b->payload_bits = 0;
    else b->payload_bits = (uint64_t)(n - 1) * 8 + (uint64_t)(7 - __builtin_ctz(data[n - 1]));
    br_refill(b);
}

/* bits consumed so far */
static inline uint64_t br_pos(const br_t *b) { return (uint64_t)b->pos * 8 - (uint64_t)b->bits; }

static inline uint32_t br_peek(br_t *b, int n)  /* 1 <= n <= 32 */
{
    if (b->bits < n) br_refill(b);
    return (uint32_t)(b->cache >> (64 - n));
}
static inline void br_skip(br_t *b, int n) { b->cache <<= n; b->bits -= n; }
static inline uint32_t br_get(br_t *b, int n)   /* 0 <= n <= 32 */
{
    uint32_t v;
    if (n == 0) return 0;
    v = br_peek(b, n);
    br_skip(b, n);
    return v;
}
static inline uint32_t br_get1(br_t *b)
{
    uint32_t v;
    if (b->bits < 1) br_refill(b);
    v = (uint32_t)(b->cache >> 63);
    b->cache <<= 1; b->bits--;
    return v;
}

/* ue(v); returns 0xffffffff on a malformed (over-long) code */
static inline uint32_t br_ue(br_t *b)
{
    uint32_t v;
    int lz;
    if (b->bits < 32) br_refill(b);
    v = (uint32_t)(b->cache >> 32);
    if (v & 0x80000000u) { br_skip(b, 1); return 0; }
    if (v == 0) {
        /* 32+ leading zeros: only 2^32-1 is representable (codeNum 2^32-1); treat as malformed */
        br_skip(b, 32);
        return 0xffffffffu;
    }
    lz = __builtin_clz(v);
    if (lz <= 15) {
        v >>= (31 - 2 * lz);
        br_skip(b, 2 * lz + 1);
        return v - 1;
    }
    br_skip(b, lz);
    v = br_get(b, lz + 1);
    return v - 1;
}
static inline int32_t br_se(br_t *b)
{
    uint32_t k = br_ue(b);
    if (k == 0xffffffffu) return INT32_MIN;
    return (k & 1) ? (int32_t)((k + 1) >> 1) : -(int32_t)(k >> 1);
}
/* te(v) with range cMax (9.1): 1 bit inverted when cMax == 1 */
static inline uint32_t br_te(br_t *b, uint32_t cmax) { return cmax > 1 ? br_ue(b) : !br_get1(b); }

static inline int br_more_data(const br_t *b) { return br_pos(b) < b->payload_bits; }
static inline int br_overrun(const br_t *b) { return br_pos(b) > (uint64_t)b->len * 8; }
static inline void br_align(br_t *b) { int r = (int)(br_pos(b) & 7); if (r) { if (b->bits < 8) br_refill(b); br_skip(b, 8 - r); } }
/* reposition at an absolute byte offset (used after raw I_PCM bytes) */
static inline void br_seek_bytes(br_t *b, size_t byte_pos) { b->pos = byte_pos; b->cache = 0; b->bits = 0; br_refill(b); }

#endif
