/* h264_mp4.c — ISO-BMFF (MP4) -> Annex-B for the decoder's input side.
 *
 * The reference's Player demuxes in JavaScript (Player/mp4.js): it walks
 * moov/trak/mdia/minf/stbl, takes SPS[0] and PPS[0] from the avcC box (:414-431),
 * and feeds, per sample, each length-prefixed NAL unit to the decoder
 * (getSampleNALUnits :711-723, play loop :867-876; sample location from
 * stsc/stco/stsz).  This is the same walk in C, producing one Annex-B byte
 * stream (00 00 00 01 before every NAL) that h264bsdDecode / H264SwDecDecode /
 * h264b200DecodeStreams / h264b200SplitGops consume.  Differences: every SPS and
 * PPS of the avcC box is emitted (the Player only sends the first of each), NAL
 * length fields of 1, 2 or 4 bytes are accepted (the Player asserts 4), co64 is
 * understood, and the first track with an avc1 sample entry is used (the Player
 * hard-codes track 1).
 */
#include <string.h>
#include "h264b200_batch.h"

static uint32_t rd32(const uint8_t *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
static uint64_t rd64(const uint8_t *p) { return ((uint64_t)rd32(p) << 32) | rd32(p + 4); }

typedef struct { const uint8_t *p; size_t len; } span_t;

/* find the first child box `type` inside [p, p+len); returns its payload */
static int find_box(span_t in, const char *type, span_t *out)
{
    size_t off = 0;
    while (off + 8 <= in.len) {
        uint64_t size = rd32(in.p + off);
        size_t hdr = 8;
        if (size == 1) { if (in.len - off < 16) return -1; size = rd64(in.p + off + 8); hdr = 16; }
        else if (size == 0) size = in.len - off;
        if (size < hdr || size > in.len - off) return -1;      /* compared without a sum that could wrap (64-bit largesize) */
        if (!memcmp(in.p + off + 4, type, 4)) { out->p = in.p + off + hdr; out->len = (size_t)size - hdr; return 0; }
        off += (size_t)size;
    }
    return -1;
}

typedef struct {
    span_t avcc, stsz, stsc, stco, co64;
} track_t;

static int open_track(span_t trak, track_t *t)
{
    span_t mdia, minf, stbl, stsd, e;
    memset(t, 0, sizeof *t);
    if (find_box(trak, "mdia", &mdia) || find_box(mdia, "minf", &minf) || find_box(minf, "stbl", &stbl) || find_box(stbl, "stsd", &stsd)) return -1;
    if (stsd.len < 8) return -1;
    e.p = stsd.p + 8; e.len = stsd.len - 8;                    /* version/flags + entry_count */
    {
        span_t avc1;
        if (find_box(e, "avc1", &avc1) || avc1.len < 78) return -1;
        avc1.p += 78; avc1.len -= 78;                          /* VisualSampleEntry fields */
        if (find_box(avc1, "avcC", &t->avcc)) return -1;
    }
    if (find_box(stbl, "stsz", &t->stsz) || find_box(stbl, "stsc", &t->stsc)) return -1;
    if (find_box(stbl, "stco", &t->stco) && find_box(stbl, "co64", &t->co64)) return -1;
    return 0;
}

#define PUT(src, n) do { if (o + 4 + (n) > cap) return -2; out[o] = 0; out[o + 1] = 0; out[o + 2] = 0; out[o + 3] = 1; memcpy(out + o + 4, (src), (n)); o += 4 + (n); } while (0)

long h264b200Mp4ToAnnexB(const uint8_t *mp4, size_t len, uint8_t *out, size_t cap, size_t *out_len)
{
    span_t file, moov, trak_area;
    track_t t;
    size_t o = 0, off;
    int found = 0, lsz;
    uint32_t n_samples, fixed_size, n_chunks, n_stsc, i, sample = 0, chunk;
    if (!mp4 || !out || !out_len) return -1;
    file.p = mp4; file.len = len;
    if (find_box(file, "moov", &moov)) return -1;
    /* first trak with an avc1 sample entry */
    trak_area = moov; off = 0;
    while (off + 8 <= trak_area.len) {
        uint64_t size = rd32(trak_area.p + off);
        size_t hdr = 8;
        if (size == 1) { if (trak_area.len - off < 16) return -1; size = rd64(trak_area.p + off + 8); hdr = 16; }
        else if (size == 0) size = trak_area.len - off;
        if (size < hdr || size > trak_area.len - off) return -1;
        if (!memcmp(trak_area.p + off + 4, "trak", 4)) {
            span_t trak; trak.p = trak_area.p + off + hdr; trak.len = (size_t)size - hdr;
            if (!open_track(trak, &t)) { found = 1; break; }
        }
        off += (size_t)size;
    }
    if (!found) return -1;

    /* avcC: parameter sets (Player/mp4.js:414-431) */
    {
        const uint8_t *p = t.avcc.p, *e = p + t.avcc.len;
        int cnt, k;
        if (t.avcc.len < 7) return -1;
        lsz = (p[4] & 3) + 1;
        if (lsz == 3) return -1;
        cnt = p[5] & 31; p += 6;
        for (k = 0; k < cnt; k++) { uint32_t n; if (p + 2 > e) return -1; n = ((uint32_t)p[0] << 8) | p[1]; p += 2; if (p + n > e) return -1; PUT(p, n); p += n; }
        if (p + 1 > e) return -1;
        cnt = *p++;
        for (k = 0; k < cnt; k++) { uint32_t n; if (p + 2 > e) return -1; n = ((uint32_t)p[0] << 8) | p[1]; p += 2; if (p + n > e) return -1; PUT(p, n); p += n; }
    }

    /* sample tables */
    if (t.stsz.len < 12) return -1;
    fixed_size = rd32(t.stsz.p + 4); n_samples = rd32(t.stsz.p + 8);
    if (!fixed_size && t.stsz.len < 12 + (size_t)n_samples * 4) return -1;
    if (t.stsc.len < 8) return -1;
    n_stsc = rd32(t.stsc.p + 4);
    if (t.stsc.len < 8 + (size_t)n_stsc * 12) return -1;
    if (t.stco.p) { if (t.stco.len < 8) return -1; n_chunks = rd32(t.stco.p + 4); if (t.stco.len < 8 + (size_t)n_chunks * 4) return -1; }
    else { if (t.co64.len < 8) return -1; n_chunks = rd32(t.co64.p + 4); if (t.co64.len < 8 + (size_t)n_chunks * 8) return -1; }

    for (chunk = 1; chunk <= n_chunks && sample < n_samples; chunk++) {
        /* samples per chunk: last stsc entry whose first_chunk <= chunk */
        uint32_t per = 0;
        uint64_t pos;
        for (i = 0; i < n_stsc; i++) { if (rd32(t.stsc.p + 8 + i * 12) <= chunk) per = rd32(t.stsc.p + 8 + i * 12 + 4); else break; }
        pos = t.stco.p ? rd32(t.stco.p + 8 + (size_t)(chunk - 1) * 4) : rd64(t.co64.p + 8 + (size_t)(chunk - 1) * 8);
        for (i = 0; i < per && sample < n_samples; i++, sample++) {
            uint32_t ssz = fixed_size ? fixed_size : rd32(t.stsz.p + 12 + (size_t)sample * 4);
            uint64_t end;
            if (pos > len || ssz > len - pos) return -1;        /* a chunk offset near 2^64 must not wrap pos + ssz */
            end = pos + ssz;
            while (pos + (uint64_t)lsz <= end) {                /* length-prefixed NAL units (mp4.js:711-723) */
                uint32_t n = lsz == 4 ? rd32(mp4 + pos) : lsz == 2 ? (((uint32_t)mp4[pos] << 8) | mp4[pos + 1]) : mp4[pos];
                pos += (uint64_t)lsz;
                if (n > end - pos) return -1;
                if (n) PUT(mp4 + pos, n);
                pos += n;
            }
            pos = end;
        }
    }
    *out_len = o;
    return (long)sample;
}
