#!/bin/bash
mkdir -p gpurun_out
for v in default norun; do
lib=""; [ "$v" != default ] && lib=$PWD/tools/exp/variants/$v.so
VXL_LIB=$lib ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_reflection|k_local_lights' -s 4 -c 2 --csv --log-file gpurun_out/r2run_$v.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open("gpurun_out/r2run_$v.csv") if not l.startswith("=="))]
for r in rows: print("$v", r["Kernel Name"][:30], r["Metric Name"], r["Metric Value"])
PY
done
