#!/bin/bash
# Final round-2 record on one B200: smoke, GPU parity suite, default bench (C2 + C3 + build blocks), reference arm, ncu.
O=gpurun_out
python __graft_entry__.py --smoke > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02_smoke.log
python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > $O/r02_plain_c2.json 2> $O/r02_plain_c2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_ncu_launches_c2.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tensor_filter2 -s 3 -c 1 -o /tmp/prof_tf $CMD > $O/r02_ncu_tf.log 2>&1
ncu -i /tmp/prof_tf.ncu-rep --page raw --csv > $O/r02_ncu_tensor_filter2_raw.csv 2>/dev/null
ncu -i /tmp/prof_tf.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_ncu_tensor_filter2_source.csv.gz
ls -la $O | grep r02_ | awk '{print $5, $9}' | head -30
