# Round-end evidence run (one GPU): full GPU test suite, the four workloads, the ncu launch list and the --set full captures that profiles/r02_* are made from
set -u
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -8 > $O/r2_final_gputest.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2_final_bench.log 2>&1
timeout 600 python bench.py --workload msk1200 --steps 10 > $O/r2_final_msk.log 2>&1
timeout 900 python bench.py --workload burst1200x2048 --steps 2 --warmup 1 > $O/r2_final_burst.log 2>&1
timeout 600 python bench.py --workload mix16384 --scaling strong --steps 8 --warmup 4 > $O/r2_final_mix1.log 2>&1
B="python bench.py --steps 2 --warmup 4 --no-cpu-baseline --no-e2e --no-saturation"
ncu --clock-control none --metrics gpu__time_duration.sum -k regex:'oqpsk|cfe|viterbi|pchan|soft_reset|peak|regroup' -s 100 -c 200 --csv --log-file $O/r02_launches_bench_4096ch.csv $B > $O/ncu_1.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:oqpsk_pipe_kernel -s 60 -c 1 -f -o $O/r02_oqpsk_pipe_kernel $B > $O/ncu_2.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:pchan_frame_kernel -s 3 -c 1 -f -o $O/r02_pchan_frame_kernel $B > $O/ncu_3.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:pchan_su_kernel -s 3 -c 1 -f -o $O/r02_pchan_su_kernel $B > $O/ncu_4.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:viterbi_k7_kernel -s 18 -c 1 -f -o $O/r02_viterbi_k7_kernel $B > $O/ncu_5.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:cfe_cluster_kernel -s 30 -c 1 -f -o $O/r02_cfe_cluster_kernel $B > $O/ncu_6.log 2>&1
