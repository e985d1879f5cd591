set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/g2_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g2_tests.log
tail -5 gpurun_out/g2_tests.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "known_answers" > gpurun_out/g2_parity.log 2>&1
grep -E "known answers|passed|failed|Error|error" gpurun_out/g2_parity.log | head
python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 5 > gpurun_out/g2_short.json 2>gpurun_out/g2_short.err; python -c "
import json;d=json.load(open('gpurun_out/g2_short.json'));print('value',d['value'],d['roofline']['kernel_ms_per_launch'], d['fit'])"
AMX_EXACT_TOL=0 python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 5 > gpurun_out/g2_noexact.json 2>&1; python -c "
import json;d=json.load(open('gpurun_out/g2_noexact.json'));print('noexact',d['value'])"
python bench.py --steps 3 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; tail -3 gpurun_out/g2_bench.err; python -c "
import json;d=json.load(open('gpurun_out/g2_bench.json'))
print(json.dumps({k:d[k] for k in ('value','e2e','e2e_plugin','roofline_fp64','cpu_baseline')},indent=0)[:2500])
print(json.dumps(d.get('configs'),indent=0)[:6000])"
