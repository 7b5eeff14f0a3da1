"""Minimal MP4 (ISO-BMFF) muxer for tests: wraps an Annex-B H.264 stream (one access unit per
picture, as the in-repo writer produces it) into ftyp + moov(one avc1 video track) + mdat with
4-byte NAL length prefixes.  Test infrastructure only — the reference's bundled Player/*.mp4 clips
are absent from the mount, so this is how the demuxer gets inputs."""
import struct


def box(kind, payload):
    return struct.pack(">I4s", 8 + len(payload), kind) + payload


def full(kind, version, flags, payload):
    return box(kind, struct.pack(">I", (version << 24) | flags) + payload)


def split_annexb(data):
    nals, i, n = [], 0, len(data)
    while True:
        j = data.find(b"\x00\x00\x01", i)
        if j < 0:
            break
        k = data.find(b"\x00\x00\x01", j + 3)
        end = n if k < 0 else (k - 1 if data[k - 1] == 0 else k)
        nals.append(data[j + 3:end])
        i = j + 3 if k < 0 else k
        if k < 0:
            break
    return nals


def mux(annexb, width, height, samples_per_chunk=3, length_size=4, use_co64=False):
    nals = split_annexb(annexb)
    sps = [x for x in nals if x[0] & 31 == 7]
    pps = [x for x in nals if x[0] & 31 == 8]
    samples, cur = [], []
    for x in nals:
        t = x[0] & 31
        if t in (7, 8):
            continue
        if t in (1, 5) and (x[1] & 0x80) and cur and any(c[0] & 31 in (1, 5) for c in cur):
            samples.append(cur); cur = []          # first_mb_in_slice == 0 opens a new picture
        cur.append(x)
    if cur:
        samples.append(cur)
    pack = {4: ">I", 2: ">H", 1: ">B"}[length_size]
    blobs = [b"".join(struct.pack(pack, len(x)) + x for x in s) for s in samples]
    ftyp = box(b"ftyp", b"isom" + struct.pack(">I", 0x200) + b"isomavc1")
    avcc = box(b"avcC", bytes([1, sps[0][1], sps[0][2], sps[0][3], 0xFC | (length_size - 1), 0xE0 | len(sps)]) +
               b"".join(struct.pack(">H", len(x)) + x for x in sps) + bytes([len(pps)]) + b"".join(struct.pack(">H", len(x)) + x for x in pps))
    avc1 = box(b"avc1", b"\x00" * 6 + struct.pack(">H", 1) + b"\x00" * 16 + struct.pack(">HH", width, height) +
               struct.pack(">II", 0x480000, 0x480000) + b"\x00" * 4 + struct.pack(">H", 1) + b"\x00" * 32 + struct.pack(">Hh", 24, -1) + avcc)
    stsd = full(b"stsd", 0, 0, struct.pack(">I", 1) + avc1)
    stts = full(b"stts", 0, 0, struct.pack(">III", 1, len(blobs), 1000))
    stsz = full(b"stsz", 0, 0, struct.pack(">II", 0, len(blobs)) + b"".join(struct.pack(">I", len(b_)) for b_ in blobs))
    n_chunks = (len(blobs) + samples_per_chunk - 1) // samples_per_chunk
    last = len(blobs) - (n_chunks - 1) * samples_per_chunk
    entries = [(1, samples_per_chunk, 1)] + ([(n_chunks, last, 1)] if last != samples_per_chunk and n_chunks > 1 else [])
    if n_chunks == 1:
        entries = [(1, len(blobs), 1)]
    stsc = full(b"stsc", 0, 0, struct.pack(">I", len(entries)) + b"".join(struct.pack(">III", *e) for e in entries))

    def moov_with(offsets):
        co = (full(b"co64", 0, 0, struct.pack(">I", len(offsets)) + b"".join(struct.pack(">Q", o) for o in offsets)) if use_co64 else
              full(b"stco", 0, 0, struct.pack(">I", len(offsets)) + b"".join(struct.pack(">I", o) for o in offsets)))
        stbl = box(b"stbl", stsd + stts + stsc + stsz + co)
        minf = box(b"minf", full(b"vmhd", 0, 1, b"\x00" * 8) + stbl)
        mdia = box(b"mdia", full(b"mdhd", 0, 0, struct.pack(">IIIIHH", 0, 0, 30000, 1000 * len(blobs), 0x55C4, 0)) +
                   full(b"hdlr", 0, 0, b"\x00" * 4 + b"vide" + b"\x00" * 12 + b"v\x00") + minf)
        trak = box(b"trak", full(b"tkhd", 0, 3, b"\x00" * 80) + mdia)
        return box(b"moov", full(b"mvhd", 0, 0, b"\x00" * 96) + trak)
    probe = moov_with([0] * n_chunks)
    base = len(ftyp) + len(probe) + 8
    offsets, pos = [], base
    for c in range(n_chunks):
        offsets.append(pos)
        pos += sum(len(b_) for b_ in blobs[c * samples_per_chunk:(c + 1) * samples_per_chunk])
    return ftyp + moov_with(offsets) + box(b"mdat", b"".join(blobs)), len(blobs)
