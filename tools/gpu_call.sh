set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/g13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g13_tests.log
tail -8 gpurun_out/g13_tests.log
grep -E "atoms:" gpurun_out/g13_tests.log
python -m pytest tests/test_gpu_parity.py -m gpu -q -s -k "custom_grids" 2>&1 | grep -E "atoms|passed|failed"
