set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/g8_c2.json 2> gpurun_out/g8_c2.err
timeout 300 $TR bench.py --gpus 8 --steps 5 --warmup 3 --workload c3 > gpurun_out/g8_c3.json 2> gpurun_out/g8_c3.err
timeout 400 $TR bench.py --gpus 8 --steps 3 --warmup 3 --workload c4 > gpurun_out/g8_c4.json 2> gpurun_out/g8_c4.err
cat gpurun_out/g8_c2.json gpurun_out/g8_c3.json gpurun_out/g8_c4.json | python tools/fmt_bench.py
tail -3 gpurun_out/g8_c2.err gpurun_out/g8_c3.err gpurun_out/g8_c4.err
