#!/bin/sh
# Build an experimental variant of libh264b200.so (host C recompiled with extra flags, CUDA object reused):
#   tools/build_variant.sh NAME "extra gcc flags"   ->  build/variants/libh264b200_NAME.so
# Run it with H264B200_LIB=build/variants/libh264b200_NAME.so python bench.py --e2e-only ...
set -e
name=$1; shift
C=broadway_b200/csrc; out=build/variants; mkdir -p $out/$name
objs=""
for f in h264_decoder.c h264_params.c h264_dpb.c h264_slice.c h264_cavlc.c h264_swdec.c h264_runner.c h264_mp4.c h264_shim.c; do
  gcc -O3 -g -fPIC -pthread -Iinclude -I$C $@ -c $C/$f -o $out/$name/$f.o
  objs="$objs $out/$name/$f.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libh264b200_$name.so $objs build/h264_engine.cu.o -cudart static -lpthread -ldl -lrt
echo built $out/libh264b200_$name.so
