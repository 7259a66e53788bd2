#!/bin/bash
# A/B of the strip kernel geometry on the attitude workloads (run on the GPU box)
out=gpurun_out/strip_sweep.log; : > $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "attitude or chain or window or lean" >> $out 2>&1
run() { echo "== $*" >> $out; env "$@" python bench.py --workload $W --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('   ms/stage %.4f  hbm frac %.3f  kernel %s' % (d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel']))
" >> $out; }
for W in ${WORKLOADS:-attitude_x16_3x16000x4800x3 attitude_x4_3x4000x1200x3}; do
  echo "#### $W" >> $out
  while read -r cfg; do [ -n "$cfg" ] && run $cfg; done <<< "${CONFIGS:-BELLMAN_WIN_NOSTRIP=1
BELLMAN_STRIP_NW=4 BELLMAN_STRIP_R=8}"
done
cat $out
