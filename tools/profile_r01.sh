set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_noddi_stage -s 9 -c 3 -o gpurun_out/full_stage -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-pipeline > gpurun_out/full_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_preprocess|k_dti|k_scatter_maps" -s 2 -c 5 -o gpurun_out/full_pipe -f python tools/bench_pipeline.py --cfg 2 --steps 1 > gpurun_out/full_pipe.log 2>&1
python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err
tail -2 gpurun_out/bench_r01.err
cat gpurun_out/bench_r01.json
