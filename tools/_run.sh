set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_pytest_gpu.log
python bench.py > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_c2_ref.json 2> gpurun_out/final_c2_ref.err
python bench.py --workload c3 > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err
python bench.py --workload c3 --impl reference --steps 3 --warmup 1 > gpurun_out/final_c3_ref.json 2> gpurun_out/final_c3_ref.err
python bench.py --workload c1 > gpurun_out/final_c1.json 2> gpurun_out/final_c1.err
python bench.py --workload c1 --nq 10000 --no-cpu-baseline > gpurun_out/final_c1_10k.json 2>> gpurun_out/final_c1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_c2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_c3.csv python bench.py --workload c3 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_c3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tensor_filter_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/final_prof_tf python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_f_tf.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hnsw_search_kernel --launch-skip 4 --launch-count 1 -f -o gpurun_out/final_prof_hnsw python bench.py --workload c3 --no-cpu-baseline --steps 2 --warmup 1 > gpurun_out/ncu_f_hnsw.log 2>&1
timeout 500 python tools/sweep_c5.py --out gpurun_out/final_c5.json > gpurun_out/final_c5.log 2>&1
cat gpurun_out/final_c2.json gpurun_out/final_c3.json gpurun_out/final_c1.json gpurun_out/final_c1_10k.json | python tools/fmt_bench.py
cat gpurun_out/final_pytest_gpu.log gpurun_out/final_smoke.log
