set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/h13_tests.log 2>&1; tail -15 gpurun_out/h13_tests.log
bash tools/bench_sweep.sh "AMX_LEAN2=1" > gpurun_out/h13_sweep.log 2>&1
cat gpurun_out/h13_sweep.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:k_noddi_stage3 -c 1 --csv --log-file gpurun_out/h13_stage.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-pipeline --no-configs > /dev/null 2>&1
python - <<'P'
import csv
for r in csv.reader(open('gpurun_out/h13_stage.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[4][:50], r[7], r[-3], r[-1])
P
