#!/bin/sh
# Go / no-go measurement for "motion vector prediction off the host" (DESIGN.md section 9, item 1).
#
# Builds two libraries from a scratch copy of the host sources: the tree as it is (base) and one whose
# predict_mv() returns (0,0) at once (nopred: WRONG vectors, same bitstream parse — a measurement aid, never a
# product build), then prints the command that times both on the GPU box's own cores:
#   tools/box_parser_ab.sh && gpurun --timeout 200 -- 'sh build/variants/run_parser_ab.sh'
# Compare parse_core_s (host parse core-seconds per step; repeats to +-0.3 % on that box).
set -e
C=broadway_b200/csrc; out=build/variants; tmp=$out/nopred_src
[ -f build/h264_engine.cu.o ] || { echo "build the product first (python -c 'import __graft_entry__ as g; g.build()')"; exit 1; }
rm -rf $tmp; mkdir -p $tmp $out/base $out/nopred
cp $C/*.c $C/*.h $tmp/
python - "$tmp/h264_slice.c" <<'PY'
import sys
p = sys.argv[1]; s = open(p).read()
key = "static void predict_mv(const sl_t *s, int x4, int y4, int w4, int ref, unsigned done, int dir, int *px, int *py)\n{\n"
assert key in s, "predict_mv signature changed: update tools/box_parser_ab.sh"
s = s.replace(key, key + "    *px = 0; *py = 0; (void)s; (void)x4; (void)y4; (void)w4; (void)ref; (void)done; (void)dir; return;\n", 1)
open(p, "w").write(s)
PY
for v in base nopred; do
  src=$C; [ $v = nopred ] && src=$tmp
  objs=""
  for f in h264_decoder.c h264_params.c h264_dpb.c h264_slice.c h264_cavlc.c h264_swdec.c h264_runner.c h264_mp4.c h264_shim.c; do
    gcc -O3 -g -fPIC -pthread -w -Iinclude -I$src -c $src/$f -o $out/$v/$f.o; objs="$objs $out/$v/$f.o"
  done
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libh264b200_$v.so $objs build/h264_engine.cu.o -cudart static -lpthread -ldl -lrt
done
cat > $out/run_parser_ab.sh <<'EOS'
for v in base nopred base nopred; do
  H264B200_LIB=build/variants/libh264b200_$v.so timeout 80 python bench.py --e2e-only --no-check --steps 3 --warmup 2 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'e2e', round(d['e2e_fps']), 'parse_core_s', round(d['parse_core_s'],3))"
done
EOS
echo "built $out/libh264b200_base.so and $out/libh264b200_nopred.so; run:  gpurun --timeout 200 -- 'sh build/variants/run_parser_ab.sh'"
