# 8-GPU runs of round 2 (gpurun --gpus 8): default workload, cfg 5 mix strong and weak; results in profiles/r02_bench_lines.jsonl
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611"
timeout 240 $T bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-saturation > gpurun_out/n8_bench.log 2>&1
timeout 240 $T bench.py --gpus 8 --workload mix16384 --scaling strong --steps 6 --warmup 3 > gpurun_out/n8_mix_strong.log 2>&1
timeout 200 $T bench.py --gpus 8 --workload mix16384 --scaling weak --steps 6 --warmup 3 > gpurun_out/n8_mix_weak.log 2>&1
