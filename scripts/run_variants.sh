for v in "" _u2 _u3 _u4 _u6; do
  export B2POINTS_LIB=$PWD/gtsam_points_b200/lib/libb2points$v.so
  for fl in "" "--no-flush"; do
    timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline $fl 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('VARIANT[$v] [$fl]', round(d['ms_per_step']*1e3,2),'us  e2e',round(d['e2e']['ms_per_step']*1e3,2),'us frac',round(d['roofline']['frac'],3))"
  done
done
