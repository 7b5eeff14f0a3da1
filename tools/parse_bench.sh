#!/bin/sh
# Host-parse-only throughput of the product decoder (null backend: pictures are parsed into records and dropped).
# usage: tools/parse_bench.sh file.264 [reps] ; prints the best of 5 runs
C=broadway_b200/csrc
gcc -O3 -g -Iinclude -I$C -DUSE_B200 oracle/null_backend.c $C/h264_decoder.c $C/h264_params.c $C/h264_dpb.c $C/h264_slice.c $C/h264_cavlc.c $C/h264_swdec.c tools/swdec_cli.c -o /tmp/nulldec || exit 1
for i in 1 2 3 4 5; do taskset -c 3 /tmp/nulldec -r ${2:-5} $1 | tail -1 | sed 's/.*"fps": \([0-9.]*\).*/\1/'; done | sort -n | tail -1
