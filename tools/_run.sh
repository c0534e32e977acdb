python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_batcher.py tests/test_gpu_rdb.py tests/test_gpu_flat.py -m gpu -x -q 2>&1 | tail -3
for o in "hnsw_global=1" "hnsw_global=1 --opt hnsw_gather=3" "hnsw_global=1 --opt hnsw_gather=3 --opt hnsw_per_sm=12" "hnsw_global=1 --opt hnsw_gather=3 --opt hnsw_hash=8192" "hnsw_gather=3 --opt hnsw_hash=6144"; do
  echo "== $o"
  python bench.py --workload c3 --no-cpu-baseline --opt $o 2>gpurun_out/err.log | python tools/fmt_bench.py
done
for o in "hnsw_global=1" "hnsw_global=1 --opt hnsw_gather=3"; do
python bench.py --workload c1 --no-cpu-baseline --opt $o 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --nq 10000 --no-cpu-baseline --opt $o  2>>gpurun_out/err.log | python tools/fmt_bench.py
done
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/prof_hnsw_r1h python bench.py --workload c3 --no-cpu-baseline --steps 2 --warmup 1 --opt hnsw_global=1 --opt hnsw_gather=3 > gpurun_out/ncu_hnsw_h.log 2>&1
