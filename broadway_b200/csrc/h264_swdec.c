/* h264_swdec.c — the H264SwDec* instance API over the h264bsd core
 * (reference: Decoder/src/H264SwDecApi.c:124-569, Decoder/inc/H264SwDecApi.h).
 * Same state machine: H264SwDecDecode loops h264bsdDecode until the buffer is
 * consumed or a picture / new headers are ready (:391-463) and translates the
 * core's return codes; H264SwDecNextPicture pops the display queue (:524-569). */
#include <stdlib.h>
#include <string.h>
#include "h264b200_swdec.h"

int h264_decoder_flushed_pending(storage_t *s);

/* H264SwDecApi.c:78-96: the embedder hooks, here as weak defaults over the C library */
__attribute__((weak)) void H264SwDecTrace(char *string) { (void)string; }
__attribute__((weak)) void *H264SwDecMalloc(u32 size) { return malloc(size); }
__attribute__((weak)) void H264SwDecFree(void *ptr) { free(ptr); }
__attribute__((weak)) void H264SwDecMemcpy(void *dest, void *src, u32 count) { memcpy(dest, src, count); }
__attribute__((weak)) void H264SwDecMemset(void *ptr, i32 value, u32 count) { memset(ptr, value, count); }

enum { ST_UNINITIALIZED = 0, ST_INITIALIZED, ST_NEW_HEADERS };
typedef struct { int stat; u32 pic_number; storage_t storage; } container_t;

H264SwDecRet H264SwDecInit(H264SwDecInst *decInst, u32 noOutputReordering)
{
    container_t *c;
    if (!decInst) return H264SWDEC_PARAM_ERR;
    *decInst = NULL;
    if (((-1) >> 1) != (-1)) return H264SWDEC_INITFAIL;       /* arithmetic right shift required (H264SwDecApi.c:134) */
    c = (container_t *)H264SwDecMalloc((u32)sizeof *c);
    if (!c) return H264SWDEC_MEMFAIL;
    H264SwDecMemset(c, 0, (u32)sizeof *c);
    if (h264bsdInit(&c->storage, noOutputReordering) != HANTRO_OK) { H264SwDecFree(c); return H264SWDEC_INITFAIL; }
    c->stat = ST_INITIALIZED;
    *decInst = c;
    return H264SWDEC_OK;
}

H264SwDecRet H264SwDecGetInfo(H264SwDecInst decInst, H264SwDecInfo *info)
{
    container_t *c = (container_t *)decInst;
    storage_t *s;
    if (!c || !info) return H264SWDEC_PARAM_ERR;
    s = &c->storage;
    if (!h264bsdPicWidth(s)) return H264SWDEC_HDRS_NOT_RDY;
    info->profile = h264bsdProfile(s);
    info->picWidth = h264bsdPicWidth(s) << 4;
    info->picHeight = h264bsdPicHeight(s) << 4;
    info->videoRange = h264bsdVideoRange(s);
    info->matrixCoefficients = h264bsdMatrixCoefficients(s);
    h264bsdCroppingParams(s, &info->croppingFlag, &info->cropParams.cropLeftOffset, &info->cropParams.cropOutWidth,
                          &info->cropParams.cropTopOffset, &info->cropParams.cropOutHeight);
    h264bsdSampleAspectRatio(s, &info->parWidth, &info->parHeight);
    return H264SWDEC_OK;
}

void H264SwDecRelease(H264SwDecInst decInst)
{
    container_t *c = (container_t *)decInst;
    if (!c) return;
    h264bsdShutdown(&c->storage);
    H264SwDecFree(c);
}

H264SwDecRet H264SwDecDecode(H264SwDecInst decInst, H264SwDecInput *in, H264SwDecOutput *out)
{
    container_t *c = (container_t *)decInst;
    u32 len, nread = 0, res;
    u8 *p;
    H264SwDecRet ret = H264SWDEC_STRM_PROCESSED;
    if (!in || !out || !in->pStream || !in->dataLen) return H264SWDEC_PARAM_ERR;
    if (!c || c->stat == ST_UNINITIALIZED) return H264SWDEC_NOT_INITIALIZED;
    out->pStrmCurrPos = NULL;
    len = in->dataLen; p = in->pStream;
    do {
        if (c->stat == ST_NEW_HEADERS) { res = H264BSD_HDRS_RDY; c->stat = ST_INITIALIZED; nread = 0; }
        else res = h264bsdDecode(&c->storage, p, len, in->picId, &nread);
        p += nread;
        len = nread <= len ? len - nread : 0;
        out->pStrmCurrPos = p;
        switch (res) {
        case H264BSD_HDRS_RDY:
            if (h264_decoder_flushed_pending(&c->storage)) { c->stat = ST_NEW_HEADERS; ret = H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY; }
            else ret = H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY;
            len = 0;
            break;
        case H264BSD_PIC_RDY:
            c->pic_number++;
            ret = len == 0 ? H264SWDEC_PIC_RDY : H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY;
            len = 0;
            break;
        case H264BSD_PARAM_SET_ERROR:
            if (!h264bsdCheckValidParamSets(&c->storage) && len == 0) ret = H264SWDEC_STRM_ERR;
            break;
        case H264BSD_MEMALLOC_ERROR:
            ret = H264SWDEC_MEMFAIL; len = 0;
            break;
        default: break;
        }
    } while (len);
    return ret;
}

H264SwDecApiVersion H264SwDecGetAPIVersion(void)
{
    H264SwDecApiVersion v; v.major = 2; v.minor = 3; return v;
}

H264SwDecRet H264SwDecNextPicture(H264SwDecInst decInst, H264SwDecPicture *pic, u32 flushBuffer)
{
    container_t *c = (container_t *)decInst;
    u32 id, idr, err; u8 *p;
    if (!c || !pic) return H264SWDEC_PARAM_ERR;
    if (flushBuffer) h264bsdFlushBuffer(&c->storage);
    p = h264bsdNextOutputPicture(&c->storage, &id, &idr, &err);
    if (!p) return H264SWDEC_OK;
    pic->pOutputPicture = (u32 *)p; pic->picId = id; pic->isIdrPicture = idr; pic->nbrOfErrMBs = err;
    return H264SWDEC_PIC_RDY;
}
