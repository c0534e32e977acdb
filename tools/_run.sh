python -m pytest tests -m gpu -x -q -s 2>&1 | tail -8
python bench.py --workload c3 --no-cpu-baseline 2>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
python bench.py --workload c1 --nq 10000 --no-cpu-baseline 2>>gpurun_out/err.log | python tools/fmt_bench.py
