cd $GRAFT_REPO_ROOT
timeout 200 python tools/time_stages.py rgbd_1080p 32 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/s3_launches.csv python tools/time_stages.py rgbd_1080p 32 > gpurun_out/s3_ncu.log 2>&1
tail -2 gpurun_out/s3_ncu.log
python - <<'PY'
import csv
rows = list(csv.reader(open('gpurun_out/s3_launches.csv')))
hdr = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hdr]
ki, vi, ui = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
for r in rows[hdr+1:hdr+40]:
    print(r[ki][:40], r[vi], r[ui])
PY
