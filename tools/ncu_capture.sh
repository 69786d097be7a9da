#!/bin/bash
# ncu evidence for every kernel of the path (run under gpurun, ONE GPU). Reports land in gpurun_out/; summaries are made
# here afterwards with tools/ncu_lines.py and committed under profiles/ (r02_*).
set -u
O=gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-saturation"
NCU="ncu --clock-control none"
# 1. launch list of the bench command (shares of the step)
$NCU --metrics gpu__time_duration.sum -k regex:'oqpsk|cfe|viterbi|pchan|soft_reset|peak' -c 260 --csv --log-file $O/r02_launches_bench_4096ch.csv $B > $O/ncu_1.log 2>&1
full() { # name regex skip command...
  local name=$1 rx=$2 skip=$3; shift 3
  $NCU --set full --import-source on -k regex:$rx -s $skip -c 1 -f -o $O/r02_$name "$@" > $O/ncu_$name.log 2>&1
}
full oqpsk_pipe_kernel oqpsk_pipe_kernel 20 $B
full cfe_cluster_kernel cfe_cluster_kernel 10 $B
full viterbi_k7_kernel viterbi_k7_kernel 5 $B
full pchan_frame_kernel pchan_frame_kernel 1 $B
M="python bench.py --workload msk1200 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e"
full msk_pipe_kernel msk_pipe_kernel 30 $M
full cfe_row_logmag_kernel cfe_row_logmag_kernel 10 $M
U="python bench.py --workload burst1200x2048 --channels 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
full burst_front_kernel burst_front_kernel 4 $U
full burst_back_kernel burst_back_kernel 4 $U
full trident_fft_kernel trident_fft_kernel 1 $U
full hilbert_block_kernel hilbert_block_kernel 3 $U
X="python bench.py --workload mix16384 --channels 2048 --steps 2 --warmup 1"
full oqpsk_segment_kernel oqpsk_segment_kernel 6 $X
full fir_block_kernel fir_block_kernel 6 $X
full cchan_frame_kernel cchan_frame_kernel 1 $X
ls -la $O/*.ncu-rep | awk '{print $5, $9}' > $O/ncu_sizes.txt
