cd $GRAFT_REPO_ROOT
for v in "$@"; do
  export ORBX_LIB=$GRAFT_REPO_ROOT/tools/ab/liborbx_$v.so
  echo "== $v"
  timeout 200 python tools/time_stages.py rgbd_1080p 1 2>&1 | tail -1 | sed 's/.*liborbx_//'
  ORBX_NO_GRAPHS=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 130 -c 13 --csv --log-file gpurun_out/s21_$v.csv python tools/time_total.py rgbd_1080p 1 > /dev/null 2>&1
  python - <<PY
import csv
rows = list(csv.reader(open('gpurun_out/s21_$v.csv')))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
for r in rows[hdr+1:]:
    print('   ', r[ki].split('(')[0][-34:], r[vi], r[gi])
PY
done 2>&1 | tee gpurun_out/s21.log
