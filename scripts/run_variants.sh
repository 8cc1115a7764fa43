#!/bin/bash
# usage: scripts/run_variants.sh "<tag> <tag> ..." [extra bench args]   ("-" = the default library); one line per variant (cold L2)
tags=$1; shift
for v in $tags; do
  [ "$v" = "-" ] && v=""
  export B2POINTS_LIB=$PWD/gtsam_points_b200/lib/libb2points$v.so
  timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extra "$@" 2>gpurun_out/variant_err$v.log | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); print('VARIANT[$v]', round(d['ms_per_step']*1e3,2),'us  e2e',round(d['e2e']['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3), 'inliers', d['detail']['inliers'])
except Exception as e:
    print('VARIANT[$v] FAILED', e)"
done
