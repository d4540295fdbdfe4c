mkdir -p gpurun_out
nvidia-smi --query-gpu=timestamp,clocks.sm,clocks.mem,power.draw,pstate,clocks_event_reasons.active,utilization.gpu --format=csv,noheader -lms 50 > gpurun_out/smi_trace.csv 2>&1 &
SMI=$!
for i in 1 2 3 4; do
  date +"run $i start %s.%N" >> gpurun_out/default_wall.txt
  MF_B200_TRACE=1 MF_B200_TRACE_WALL=1 ./oracle/_ref/dropin_benchmark_snark_default > gpurun_out/defaultw_$i.out 2> gpurun_out/defaultw_$i.err
  grep -E "^(setup|prover)" gpurun_out/defaultw_$i.out | tr "\n" " "; echo
done
kill $SMI
