/* h264_consts.h — constants of ITU-T H.264 (Baseline subset) shared by the host
 * parser, the synthetic bitstream writer and (as __constant__ copies) the CUDA
 * kernels.  Values are the standard's; the reference keeps the same numbers in
 * h264bsd_vlc.c:57-63 (CBP map), h264bsd_util.c:53-55 (QPc),
 * h264bsd_transform.c:55-56 (LevelScale), h264bsd_deblocking.c:77-98
 * (alpha/beta/tc0) and h264bsd_intra_prediction.c:87-90 (block x/y).
 */
#ifndef B200_H264_CONSTS_H
#define B200_H264_CONSTS_H
#include <stdint.h>

#ifdef __CUDACC__
#define H264_TBL static __device__ __constant__ const
#else
#define H264_TBL static const
#endif

/* Table 9-4, ChromaArrayType 1: codeNum -> coded_block_pattern, [0]=Intra_4x4, [1]=Inter */
static const uint8_t H264_CBP_MAP[48][2] = {
 {47, 0},{31,16},{15, 1},{ 0, 2},{23, 4},{27, 8},{29,32},{30, 3},{ 7, 5},{11,10},{13,12},{14,15},
 {39,47},{43, 7},{45,11},{46,13},{16,14},{ 3, 6},{ 5, 9},{10,31},{12,35},{19,37},{21,42},{26,44},
 {28,33},{35,34},{37,36},{42,40},{44,39},{ 1,43},{ 2,45},{ 4,46},{ 8,17},{17,18},{18,20},{20,24},
 {24,19},{ 6,21},{ 9,26},{22,28},{25,23},{32,27},{33,29},{34,30},{36,22},{40,25},{38,38},{41,41}};

/* 4x4 zig-zag (frame) scan: scan index -> raster index (row*4+col) */
static const uint8_t H264_ZIGZAG4x4[16] = {0,1,4,8,5,2,3,6,9,12,13,10,7,11,14,15};

/* luma4x4BlkIdx -> top-left pel inside the macroblock (6.4.3) */
static const uint8_t H264_BLK_X[16] = {0,4,0,4,8,12,8,12,0,4,0,4,8,12,8,12};
static const uint8_t H264_BLK_Y[16] = {0,0,4,4,0,0,4,4,8,8,12,12,8,8,12,12};
/* raster 4x4 position (by*4+bx) -> luma4x4BlkIdx, and back */
static const uint8_t H264_RASTER_TO_BLK[16] = {0,1,4,5,2,3,6,7,8,9,12,13,10,11,14,15};
#define H264_BLK_TO_RASTER H264_RASTER_TO_BLK   /* the permutation is an involution */

/* Table 8-15: qPI -> QPc */
H264_TBL uint8_t H264_QPC[52] = {0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15,16,17,18,19,20,21,22,23,24,25,26,27,28,29,
 29,30,31,32,32,33,34,34,35,35,36,36,37,37,37,38,38,38,39,39,39,39};

/* LevelScale(qP%6, class): class 0 = positions (0,0)(0,2)(2,0)(2,2); 1 = (1,1)(1,3)(3,1)(3,3); 2 = rest */
H264_TBL uint8_t H264_LEVEL_SCALE[6][3] = {{10,16,13},{11,18,14},{13,20,16},{14,23,18},{16,25,20},{18,29,23}};
/* raster position -> class above */
H264_TBL uint8_t H264_POS_CLASS[16] = {0,2,0,2, 2,1,2,1, 0,2,0,2, 2,1,2,1};

/* Tables 8-16/8-17: deblocking thresholds */
H264_TBL uint8_t H264_ALPHA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,4,4,5,6,7,8,9,10,12,13,15,17,20,22,25,28,32,36,40,45,
 50,56,63,71,80,90,101,113,127,144,162,182,203,226,255,255};
H264_TBL uint8_t H264_BETA[52] = {0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,0,2,2,2,3,3,3,3,4,4,4,6,6,7,7,8,8,9,9,10,10,11,11,12,12,
 13,13,14,14,15,15,16,16,17,17,18,18};
/* tC0[indexA][bS-1], bS in 1..3 */
H264_TBL uint8_t H264_TC0[52][3] = {
 {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,0},
 {0,0,0},{0,0,0},{0,0,0},{0,0,0},{0,0,1},{0,0,1},{0,0,1},{0,0,1},{0,1,1},{0,1,1},{1,1,1},{1,1,1},{1,1,1},
 {1,1,1},{1,1,2},{1,1,2},{1,1,2},{1,1,2},{1,2,3},{1,2,3},{2,2,3},{2,2,4},{2,3,4},{2,3,4},{3,3,5},{3,4,6},
 {3,4,6},{4,5,7},{4,5,8},{4,6,9},{5,7,10},{6,8,11},{6,8,13},{7,10,14},{8,11,16},{9,12,18},{10,13,20},
 {11,15,23},{13,17,25}};

#endif
