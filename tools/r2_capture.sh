#!/bin/bash
# ncu captures of round 2 (run on the GPU box from the repo root): every bench command first runs to completion without ncu,
# then once under `ncu --set full` restricted to the second launch of the hot kernel.  Reports stay in /tmp; the raw / source CSV
# pages land in gpurun_out/ and are condensed with tools/ncu_summary.py into profiles/.
mkdir -p gpurun_out
cap() {   # cap <name> <kernel regex> <bench args...>
  local name=$1 kreg=$2; shift 2
  timeout 900 python bench.py "$@" --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r2_cap_${name}_plain.json 2> gpurun_out/r2_cap_${name}_plain.err
  local rc=$?
  echo "$name plain rc=$rc"
  [ $rc -ne 0 ] && return
  timeout 1500 ncu --set full --clock-control none --import-source on -k regex:$kreg --launch-skip 1 -c 1 -f -o /tmp/r2_${name} \
    python bench.py "$@" --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_cap_${name}_ncu.log 2>&1
  echo "$name ncu rc=$?"
  ncu -i /tmp/r2_${name}.ncu-rep --page raw --csv > gpurun_out/r2_${name}.raw.csv 2>/dev/null
  ncu -i /tmp/r2_${name}.ncu-rep --page source --csv > gpurun_out/r2_${name}.source.csv 2>/dev/null
}
cap lin14_200k_1chunk cf_kernel --cells 200000 --chunks 1
cap cfg2_ideal2d cf_kernel --workload cfg2
cap cfg5_vah cf_kernel --workload cfg5 --cells 20000
cap cfg4jonah cf_kernel --workload cfg4jonah --cells 20000
cap op0_cfg3 cf_kernel --operation 0 --cells 100000
# DRAM traffic of the full-size cfg3 launch (1 M cells; the plain run of this command is r2_bench_cfg3_1gpu.json)
timeout 1500 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,lts__t_bytes.sum --clock-control none \
  -k regex:cf_kernel --launch-skip 1 -c 1 --csv --log-file gpurun_out/r2_traffic_cfg3_full.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_traffic_cfg3_full.log 2>&1
echo "traffic rc=$?"
