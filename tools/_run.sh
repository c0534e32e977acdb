python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_sharded.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c2 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c2 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
