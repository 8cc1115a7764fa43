#!/bin/bash
# usage: scripts/run_gicp_variants.sh "<tag> ..."  ("-" = default library): cfg3 (GICP 500k<->500k) per library variant
for v in $1; do
  [ "$v" = "-" ] && v=""
  export B2POINTS_LIB=$PWD/gtsam_points_b200/lib/libb2points$v.so
  timeout 200 python scripts/bench_configs.py --configs cfg3 --steps 6 --warmup 2 2>&1 | tail -1 | python -c "import sys,json
try:
    d=json.loads(sys.stdin.read()); print('GICP VARIANT[$v]', round(d['ms_per_step']*1e3,1),'us', 'inliers', d['inliers'])
except Exception as e:
    print('GICP VARIANT[$v] FAILED', e)"
done
