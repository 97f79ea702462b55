# usage: bash tools/ab_ncu.sh <variant...>   ('' or "cur" = in-tree library)
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum
for v in "$@"; do
  if [ "$v" != "cur" ]; then export PTC_LIB=$PWD/build/$v/pkg/csrc/libptcore.so; else unset PTC_LIB; fi
  ncu --metrics $M --clock-control none -k regex:k_trace -s 10 -c 5 --csv --log-file gpurun_out/ncu_$v.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$v.log 2>&1
done
