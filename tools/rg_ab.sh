python -m pytest tests -m gpu -q -x -k "regroup or parity or golden or write_batch" 2>&1 | tail -3 > gpurun_out/rg_t.log
rm -f gpurun_out/rg_ab.log
for r in 1 2; do for l in search2 rg34; do echo -n "$l " >> gpurun_out/rg_ab.log; JAERO_DEBUG=1 JAERO_B200_LIB=$PWD/_abl/$l.so python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-saturation 2>gpurun_out/rg_$l.err | tail -1 >> gpurun_out/rg_ab.log; done; done
grep -h seating gpurun_out/rg_rg34.err > gpurun_out/rg_seat_new.log
