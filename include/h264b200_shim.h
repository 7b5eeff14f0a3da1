/* h264b200_shim.h — the `broadway*` functions of the reference's emscripten shim
 * (Decoder/src/Decoder.c:58-184; export list Decoder/make.py:39), exported natively by
 * libh264b200.so so that code written against that surface (one global decoder, fill the
 * stream buffer, play it, get called back per picture) runs on the B200 engine. */
#ifndef H264B200_SHIM_H
#define H264B200_SHIM_H
#include "h264b200.h"
#ifdef __cplusplus
extern "C" {
#endif

u32  broadwayInit(void);                       /* Decoder.c:74-93: 0 on success */
void broadwayExit(void);                       /* Decoder.c:164-168 */
u8  *broadwayCreateStream(u32 length);         /* Decoder.c:58-61: buffer the caller fills with NAL units / Annex-B */
void broadwayPlayStream(u32 length);           /* Decoder.c:67-70: decode `length` bytes of that buffer */
/* One deliberate difference: after a picture the reference's loop zeroes the remaining length even when the decoder
 * reported "buffer not empty" (Decoder.c:129-134, the test on PIC_RDY is commented out), i.e. it plays at most one
 * picture per call and drops the rest of the buffer.  The Player feeds one NAL unit per call, where both behave the
 * same; this library decodes everything the caller put in the buffer. */
u32  broadwayGetMajorVersion(void);            /* Decoder.c:178-184 */
u32  broadwayGetMinorVersion(void);

/* Notifications.  The reference resolves broadwayOnHeadersDecoded() and
 * broadwayOnPictureDecoded(buffer,width,height) (Decoder.c:95-97) against JavaScript; here a program
 * may define functions of those names (weak references) or register callbacks: */
typedef void (*broadway_headers_cb)(void *user);
typedef void (*broadway_picture_cb)(void *user, u8 *i420, u32 width, u32 height);
void broadwaySetCallbacks(broadway_headers_cb on_headers, broadway_picture_cb on_picture, void *user);

#ifdef __cplusplus
}
#endif
#endif
