#!/bin/sh
# Build an experimental variant of libh264b200.so with the CUDA engine recompiled with extra nvcc flags (host objects reused):
#   tools/build_cuda_variant.sh NAME -DK4_PUBLISH=4 ...   ->  build/variants/libh264b200_NAME.so
# Run it with H264B200_LIB=build/variants/libh264b200_NAME.so python bench.py --skip-e2e --no-cpu-baseline ...
set -e
name=$1; shift
C=broadway_b200/csrc; out=build/variants; mkdir -p $out/$name
objs=""
for f in h264_decoder.c h264_params.c h264_dpb.c h264_slice.c h264_cavlc.c h264_swdec.c h264_runner.c h264_mp4.c h264_shim.c; do objs="$objs build/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Iinclude -I$C "$@" -c $C/h264_engine.cu -o $out/$name/h264_engine.cu.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libh264b200_$name.so $objs $out/$name/h264_engine.cu.o -cudart static -lpthread -ldl -lrt
echo built $out/libh264b200_$name.so
