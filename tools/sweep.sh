#!/bin/bash
# usage: tools/sweep.sh "<np list>" "<ctas list>" "<iconv list>" [workload]
wl=${4:-airfoil}
for np in $1; do for ct in $2; do for ic in $3; do
  FLUIDGRID_THREADS=$np FLUIDGRID_CTAS=$ct FLUIDGRID_ICONV=$ic python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('threads',$np,'ctas',$ct,'iconv',$ic,'$wl', round(d['value']/1e6,3),'Mf/s frac', round(d['roofline']['frac'],4))"
done; done; done
