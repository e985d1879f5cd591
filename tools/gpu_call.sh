set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/g3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g3_tests.log
tail -4 gpurun_out/g3_tests.log
for i in 1 2; do python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 5 > gpurun_out/g3_new.json 2>gpurun_out/g3_new.err; python -c "
import json;d=json.load(open('gpurun_out/g3_new.json'));print('new',d['value'],d['roofline']['kernel_ms_per_launch'])"; done
AMICO_B200_LIB=/root/repo/tools/ab/libamx_base.so python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 5 > gpurun_out/g3_base.json 2>&1; python -c "
import json;d=json.load(open('gpurun_out/g3_base.json'));print('base',d['value'],d['roofline']['kernel_ms_per_launch'])"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_noddi -c 12 --csv --log-file gpurun_out/g3_launches.csv python bench.py --no-cpu --no-e2e --no-pipeline --no-configs --steps 1 --warmup 3 > /dev/null 2>&1
grep -E "k_noddi" gpurun_out/g3_launches.csv | tail -4 | cut -d, -f5,15- | cut -c1-200
