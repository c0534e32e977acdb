python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_sharded.py -m gpu -x -q 2>&1 | tail -3
python tools/sweep_c5.py --dims 128,256,512 --batches 1024,16384 --out gpurun_out/c5_x.json 2>&1 | python tools/fmt_sweep.py
python tools/sweep_c5.py --dims 1536 --batches 16384 --out gpurun_out/c5_y.json 2>&1 | python tools/fmt_sweep.py
python bench.py --workload c2 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c2 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
