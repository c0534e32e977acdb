python -m pytest tests/test_gpu_tensor.py tests/test_gpu_flat.py tests/test_gpu_sharded.py tests/test_gpu_batcher.py -m gpu -x -q 2>&1 | tail -3
python tools/sweep_c5.py --dims 128,768,1536 --batches 1,16,128,1024 --out gpurun_out/c5_x.json 2>&1 | python tools/fmt_sweep.py
