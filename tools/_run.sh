python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_flat.py -m gpu -x -q 2>&1 | tail -3
for o in "hnsw_gather=3" "hnsw_gather=4" "hnsw_gather=4 --opt hnsw_hash=8192" "hnsw_gather=4 --opt hnsw_per_sm=20" "hnsw_gather=4 --opt hnsw_per_sm=20 --opt hnsw_hash=8192"; do
  echo "== $o"
  python bench.py --workload c3 --no-cpu-baseline --opt $o 2>gpurun_out/err.log | python tools/fmt_bench.py
done
python bench.py --workload c1 --no-cpu-baseline --opt hnsw_gather=4 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --nq 10000 --no-cpu-baseline --opt hnsw_gather=4 2>>gpurun_out/err.log | python tools/fmt_bench.py
