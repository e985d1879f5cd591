set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "second_generation or known or at_scale or slow_path or variants" 2>&1 | tail -2
bash tools/bench_sweep.sh "AMX_LEAN2=1" 2>&1 | cut -c1-60
timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu --no-e2e --no-pipeline --no-configs 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('checksum', repr(d['maps_checksum']), d['roofline']['kernel_ms_per_launch'], d['roofline']['frac'])"
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_noddi_stage2 -c 1 --csv --log-file gpurun_out/h23.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-pipeline --no-configs > /dev/null 2>&1
python - <<'P'
import csv
for r in csv.reader(open('gpurun_out/h23.csv')):
    if len(r)>10 and r[0].isdigit(): print(r[4][:50], r[-3], r[-1])
P
