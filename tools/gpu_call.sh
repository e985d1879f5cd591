set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/g8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g8_tests.log
tail -8 gpurun_out/g8_tests.log
