#!/bin/bash
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --rows ${1:-125000}"
$CMD > $O/r2_plain_pair.json 2> $O/r2_plain_pair.err && \
ncu --set full --clock-control none --import-source on -k regex:tensor_filter2 -s 3 -c 1 -o /tmp/prof_tf2 $CMD > $O/r2_ncu_tf2.log 2>&1
ncu -i /tmp/prof_tf2.ncu-rep --page raw --csv > $O/r2_ncu_tensor_filter2_raw_${1:-125000}.csv 2>/dev/null
ncu -i /tmp/prof_tf2.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r2_ncu_tensor_filter2_source_${1:-125000}.csv.gz
ls -la $O/r2_ncu_tensor_filter2*
