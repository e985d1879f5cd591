set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "noddi" > gpurun_out/h11_tests.log 2>&1; tail -3 gpurun_out/h11_tests.log
bash tools/bench_sweep.sh "AMX_LEAN2=1" "AMX_STAGE2_WARPS=32" "AMX_STAGE2_WARPS=24" "AMX_STAGE1_WARPS=28" > gpurun_out/h11_sweep.log 2>&1
cat gpurun_out/h11_sweep.log
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-pipeline --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('checksum', repr(d['maps_checksum']))"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:k_noddi_stage -c 4 --csv --log-file gpurun_out/h11_stage.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-pipeline --no-configs > /dev/null 2>&1
python - <<'P'
import csv
for r in csv.reader(open('gpurun_out/h11_stage.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[4][:50], r[7], r[-3], r[-1])
P
