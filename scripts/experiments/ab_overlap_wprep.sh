#!/bin/bash
# A/B on one box: grouped weight preparation on its own stream (a fused.OVERLAP_WPREP switch that existed for this
# experiment only and was removed after it showed no difference) vs on the main stream; plain iterations.
for rep in 1 2; do
for v in True False; do
python -c "
import sys, runpy
import textboxgan_b200.fused as F
F.OVERLAP_WPREP = $v
sys.argv = ['bench.py', '--plain-only', '--no-roofline', '--no-cpu-baseline', '--no-scaling-base', '--steps', '96']
runpy.run_path('bench.py', run_name='__main__')
" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OVERLAP_WPREP=$v', round(d['ms_per_step'],4), 'ms', round(d['value'],1), 'img/s')"
done; done
