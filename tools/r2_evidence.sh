#!/bin/bash
# Round-2 evidence run on one B200: parity suite, default bench, ncu launch lists and full captures.
O=gpurun_out
python -m pytest tests -m gpu -q > $O/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02_pytest_gpu.log
python bench.py --steps 20 --warmup 5 > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_bench_reference.json 2>/dev/null; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
$CMD > $O/r02_plain_c2.json 2> $O/r02_plain_c2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_ncu_launches_c2.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tensor_filter2 -s 3 -c 1 -o /tmp/prof_tf $CMD > $O/r02_ncu_tf.log 2>&1
ncu -i /tmp/prof_tf.ncu-rep --page raw --csv > $O/r02_ncu_tensor_filter2_raw.csv 2>/dev/null
ncu -i /tmp/prof_tf.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_ncu_tensor_filter2_source.csv.gz
CMD3="python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD3 > $O/r02_plain_c3.json 2> $O/r02_plain_c3.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_ncu_launches_c3.csv $CMD3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel -s 6 -c 1 -o /tmp/prof_hnsw $CMD3 > $O/r02_ncu_hnsw.log 2>&1
ncu -i /tmp/prof_hnsw.ncu-rep --page raw --csv > $O/r02_ncu_hnsw_search_raw.csv 2>/dev/null
ncu -i /tmp/prof_hnsw.ncu-rep --page source --csv 2>/dev/null | gzip > $O/r02_ncu_hnsw_search_source.csv.gz
python tools/sweep_c5.py --dims 128,768,1536 --metric 3 --out $O/r02_c5_sweep_ip.json > /dev/null 2>&1; echo "c5 ip rc=$?"
ls -la $O | grep r02_ | awk '{print $5, $9}'
