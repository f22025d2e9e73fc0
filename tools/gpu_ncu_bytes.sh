set -u
mkdir -p gpurun_out
for OP in f16 bf16; do
  CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --operands $OP"
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:gemm_tcgen05_2cta -s 12 -c 8 --csv --log-file gpurun_out/r02h_bytes_$OP.csv $CMD > gpurun_out/r02h_bytes_$OP.log 2>&1
  echo "$OP rc=$?"
done
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu --format=csv
