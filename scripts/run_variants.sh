#!/bin/bash
# usage: scripts/run_variants.sh "<tag> <tag> ..."   ("-" = the default library); prints one line per variant x {cold, warm L2}
for v in $1; do
  [ "$v" = "-" ] && v=""
  export B2POINTS_LIB=$PWD/gtsam_points_b200/lib/libb2points$v.so
  for fl in "" "--no-flush"; do
    timeout 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $fl 2>&1 | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); print('VARIANT[$v] [$fl]', round(d['ms_per_step']*1e3,2),'us  e2e',round(d['e2e']['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3), 'inliers', d['config']['inliers'])
except Exception as e:
    print('VARIANT[$v] [$fl] FAILED', e)"
  done
done
