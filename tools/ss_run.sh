#!/bin/bash
# sort tests, the two benchmark shapes, and a per-kernel launch list of each (tools/gpu.sh 900 'bash tools/ss_run.sh tag')
tag=$1
python -m pytest tests/test_swd_gpu.py -m gpu -x -q -k "sort" > gpurun_out/pytest_ss_$tag.log 2>&1; tail -n 4 gpurun_out/pytest_ss_$tag.log
python tools/sort_bench.py 512 262144; python tools/sort_bench.py 128 2097152
for s in "a 512 262144" "b 128 2097152"; do set -- $s
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12 --csv --log-file gpurun_out/ss_launches_$1.csv python tools/sort_bench.py $2 $3 > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ss_launches_$1.csv')) if len(r)>5]
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
print('== $1', ' '.join('%s=%s' % (r[ki].split('(')[0].split('::')[-1].replace('_kernel',''), r[vi]) for r in rows[1:] if 'ss_' in r[ki] or 'radix' in r[ki]))
PY
done
