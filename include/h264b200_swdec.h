/* h264b200_swdec.h — the instance-level decoder API, exported by libh264b200.so
 * with the names, types and return codes of the reference's public header
 * (Decoder/inc/H264SwDecApi.h:52-173) so that DecTestBench.c, Decoder.c (the
 * broadway* shim), SoftAVC.cpp or any other caller of that API links against
 * this library unchanged.  Picture memory returned by H264SwDecNextPicture is
 * pinned host memory filled by an asynchronous device-to-host copy that has
 * completed by the time the call returns. */
#ifndef H264B200_SWDEC_H
#define H264B200_SWDEC_H
#include "h264b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {                                   /* H264SwDecApi.h:52-67 */
    H264SWDEC_OK = 0,
    H264SWDEC_STRM_PROCESSED = 1,
    H264SWDEC_PIC_RDY,
    H264SWDEC_PIC_RDY_BUFF_NOT_EMPTY,
    H264SWDEC_HDRS_RDY_BUFF_NOT_EMPTY,
    H264SWDEC_PARAM_ERR = -1,
    H264SWDEC_STRM_ERR = -2,
    H264SWDEC_NOT_INITIALIZED = -3,
    H264SWDEC_MEMFAIL = -4,
    H264SWDEC_INITFAIL = -5,
    H264SWDEC_HDRS_NOT_RDY = -6,
    H264SWDEC_EVALUATION_LIMIT_EXCEEDED = -7
} H264SwDecRet;

typedef void *H264SwDecInst;

typedef struct {                                 /* H264SwDecApi.h:77-85 */
    u8 *pStream;                 /* stream to decode; the decoder edits it in place */
    u32 dataLen;
    u32 picId;                   /* caller's tag, returned with the picture */
    u32 intraConcealmentMethod;  /* accepted; lost macroblocks of I pictures are always concealed the way the reference's default (0) does, csrc/k3c_conceal.cuh */
} H264SwDecInput;

typedef struct { u8 *pStrmCurrPos; } H264SwDecOutput;

typedef struct {                                 /* H264SwDecApi.h:96-103 */
    u32 *pOutputPicture;         /* planar I420, MB aligned, pinned host memory */
    u32 picId;
    u32 isIdrPicture;
    u32 nbrOfErrMBs;
} H264SwDecPicture;

typedef struct { u32 cropLeftOffset, cropOutWidth, cropTopOffset, cropOutHeight; } CropParams;

typedef struct {                                 /* H264SwDecApi.h:118-129 */
    u32 profile, picWidth, picHeight, videoRange, matrixCoefficients, parWidth, parHeight, croppingFlag;
    CropParams cropParams;
} H264SwDecInfo;

typedef struct { u32 major, minor; } H264SwDecApiVersion;

H264SwDecRet H264SwDecInit(H264SwDecInst *decInst, u32 noOutputReordering);
H264SwDecRet H264SwDecDecode(H264SwDecInst decInst, H264SwDecInput *pInput, H264SwDecOutput *pOutput);
H264SwDecRet H264SwDecNextPicture(H264SwDecInst decInst, H264SwDecPicture *pOutput, u32 endOfStream);
H264SwDecRet H264SwDecGetInfo(H264SwDecInst decInst, H264SwDecInfo *pDecInfo);
void H264SwDecRelease(H264SwDecInst decInst);
H264SwDecApiVersion H264SwDecGetAPIVersion(void);

/* Embedder hooks (H264SwDecApi.h:158-173).  The reference's API layer defines them itself with the C library
 * (H264SwDecApi.c:78-96); so does this library, as WEAK symbols: an embedder that defines its own (every test bench of
 * the reference does) overrides them at link / load time.  Every allocation of the host-side decoder state (parameter
 * sets, macroblock contexts, slice group maps, instance containers) goes through H264SwDecMalloc / H264SwDecFree.
 * Frame storage is NOT host heap memory here — it is CUDA device memory with pinned host mirrors — and is not routed
 * through the hooks. */
void H264SwDecTrace(char *string);
void *H264SwDecMalloc(u32 size);
void H264SwDecFree(void *ptr);
void H264SwDecMemcpy(void *dest, void *src, u32 count);
void H264SwDecMemset(void *ptr, i32 value, u32 count);

#ifdef __cplusplus
}
#endif
#endif
