/* h264_shim.c — the `broadway*` entry points of the reference's C shim
 * (Decoder/src/Decoder.c:58-184, exported to JavaScript by Decoder/make.py:39):
 * one global decoder instance, a caller-filled stream buffer, and two
 * notifications — headers decoded, picture decoded (I420 pointer, width,
 * height: Decoder.c:113-147, templates/DecoderPost.js:68-72).
 *
 * In the reference the notifications are JavaScript functions supplied through
 * library.js; natively they are either ordinary C functions of the embedding
 * program named broadwayOnHeadersDecoded / broadwayOnPictureDecoded (picked up
 * as weak references when the program exports them) or callbacks registered
 * with broadwaySetCallbacks (this library only).
 */
#include <stdlib.h>
#include "h264b200_swdec.h"
#include "h264b200_shim.h"

extern void broadwayOnHeadersDecoded(void) __attribute__((weak));
extern void broadwayOnPictureDecoded(u8 *buffer, u32 width, u32 height) __attribute__((weak));

static struct {
    H264SwDecInst inst;
    H264SwDecInfo info;
    u8 *stream; u32 stream_cap;
    u32 decode_number, display_number;
    broadway_headers_cb on_headers; broadway_picture_cb on_picture; void *user;
} g;

void broadwaySetCallbacks(broadway_headers_cb on_headers, broadway_picture_cb on_picture, void *user)
{ g.on_headers = on_headers; g.on_picture = on_picture; g.user = user; }

u32 broadwayInit(void)
{
    if (g.inst) { H264SwDecRelease(g.inst); g.inst = NULL; }
    if (H264SwDecInit(&g.inst, 0) != H264SWDEC_OK) { g.inst = NULL; return (u32)-1; }
    g.decode_number = g.display_number = 1;
    return 0;
}

void broadwayExit(void)
{
    if (g.inst) { H264SwDecRelease(g.inst); g.inst = NULL; }
    free(g.stream); g.stream = NULL; g.stream_cap = 0;
}

u8 *broadwayCreateStream(u32 length)
{
    if (length > g.stream_cap) {
        u8 *n = (u8 *)realloc(g.stream, (size_t)length + 16);
        if (!n) return NULL;
        g.stream = n; g.stream_cap = length;
    }
    return g.stream;
}

static void headers_ready(void)
{
    if (g.on_headers) g.on_headers(g.user);
    else if (broadwayOnHeadersDecoded) broadwayOnHeadersDecoded();
}
static void picture_ready(u8 *p)
{
    if (g.on_picture) g.on_picture(g.user, p, g.info.picWidth, g.info.picHeight);
    else if (broadwayOnPictureDecoded) broadwayOnPictureDecoded(p, g.info.picWidth, g.info.picHeight);
}

/* decode everything in the stream buffer (Decoder.c:44-53 playStream + :100-162 broadwayDecode) */
void broadwayPlayStream(u32 length)
{
    H264SwDecInput in; H264SwDecOutput out; H264SwDecPicture pic;
    if (!g.inst || !g.stream || length > g.stream_cap) return;
    in.pStream = g.stream; in.dataLen = length; in.intraConcealmentMethod = 0;
    while (in.dataLen > 0) {
        H264SwDecRet ret;
        in.picId = g.decode_number;
        ret = H264SwDecDecode(g.inst, &in, &out);
        switch (ret) {
        case H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY:
            if (H264SwDecGetInfo(g.inst, &g.info) != H264SWDEC_OK) return;
            headers_ready();
            in.dataLen -= (u32)(out.pStrmCurrPos - in.pStream); in.pStream = out.pStrmCurrPos;
            break;
        case H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY:
            in.dataLen -= (u32)(out.pStrmCurrPos - in.pStream); in.pStream = out.pStrmCurrPos;
            /* fall through */
        case H264SWDEC_PIC_RDY:
            if (ret == H264SWDEC_PIC_RDY) in.dataLen = 0;
            g.decode_number++;
            while (H264SwDecNextPicture(g.inst, &pic, 0) == H264SWDEC_PIC_RDY) { g.display_number++; picture_ready((u8 *)pic.pOutputPicture); }
            break;
        default:                                   /* stream processed / error: nothing more in this buffer */
            in.dataLen = 0;
            break;
        }
    }
}

u32 broadwayGetMajorVersion(void) { return H264SwDecGetAPIVersion().major; }
u32 broadwayGetMinorVersion(void) { return H264SwDecGetAPIVersion().minor; }
