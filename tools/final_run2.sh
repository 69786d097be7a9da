O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > $O/r2_final_gputest.log
timeout 400 python bench.py --steps 20 --warmup 3 > $O/r2_final_bench.log 2>&1
B="python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-e2e --no-saturation"
timeout 200 ncu --clock-control none --set full --import-source on -k regex:cfe_cluster_kernel -s 30 -c 1 -f -o $O/r02_cfe_cluster_kernel $B > $O/ncu_k2.log 2>&1
