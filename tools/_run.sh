python -m pytest tests/test_gpu_hnsw.py tests/test_gpu_flat.py tests/test_gpu_tensor.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c3 --no-cpu-baseline 2>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c3 --no-cpu-baseline --opt hnsw_gather=1 2>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --nq 10000 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c2 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
