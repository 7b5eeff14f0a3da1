#!/bin/sh
# run every broadway_b200/bin/nulldec_* variant on a 1080p stream, 4 times each, interleaved; print best fps
python - <<'PY'
import sys; sys.path.insert(0,'.')
from broadway_b200 import bitstream as bs
open('/tmp/p1080.264','wb').write(bs.synth(120,68,31,seed=1234))
PY
for round in 1 2 3 4; do for v in broadway_b200/bin/nulldec_*; do echo "$v $(taskset -c 5 $v -r 4 /tmp/p1080.264 | tail -1 | sed 's/.*"fps": \([0-9.]*\).*/\1/')"; done; done | sort | awk '{if($2>m[$1])m[$1]=$2} END{for(k in m)print k, m[k]}' | sort
