set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; tail -3 gpurun_out/bench_r02.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_reference.json 2>> gpurun_out/bench_r02.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > gpurun_out/launches_bench.log 2>&1
T=/tmp/ncu; mkdir -p $T
ncu --set full --clock-control none --import-source on -k regex:k_noddi_stage -s 9 -c 3 -o $T/full_stage_r02 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-pipeline --no-configs > gpurun_out/full_stage.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lasso_batched -s 2 -c 1 -o $T/full_lasso_czb_r02 -f python tools/bench_models.py 1048576 CylinderZeppelinBall5 > gpurun_out/full_czb.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lasso_small -s 2 -c 1 -o $T/full_lasso_sandi_r02 -f python tools/bench_models.py 1048576 SANDI4 > gpurun_out/full_sandi.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lasso_batched -s 2 -c 1 -o $T/full_lasso_fw_r02 -f python tools/bench_models.py 1048576 FreeWater1 > gpurun_out/full_fw.log 2>&1
python tools/ncu_to_json.py $T/full_stage_r02.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:k_noddi_stage -s 9 -c 3 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-pipeline --no-configs" 1048576 > gpurun_out/ncu_full_r02_noddi_stage_kernels.json
python tools/ncu_to_json.py $T/full_lasso_czb_r02.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:k_lasso_batched -s 2 -c 1 python tools/bench_models.py 1048576 CylinderZeppelinBall5" 1048576 > gpurun_out/ncu_full_r02_k_lasso_batched_czb.json
python tools/ncu_to_json.py $T/full_lasso_sandi_r02.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:k_lasso_small -s 2 -c 1 python tools/bench_models.py 1048576 SANDI4" 1048576 > gpurun_out/ncu_full_r02_k_lasso_small_sandi.json
python tools/ncu_to_json.py $T/full_lasso_fw_r02.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:k_lasso_batched -s 2 -c 1 python tools/bench_models.py 1048576 FreeWater1" 1048576 > gpurun_out/ncu_full_r02_k_lasso_batched_freewater.json
ncu -i $T/full_stage_r02.ncu-rep --page source --csv --print-source cuda,sass > $T/src.csv 2>/dev/null
python tools/ncu_src_lines.py $T/src.csv 60 > gpurun_out/ncu_r02_noddi_stage_by_line.txt
cp $T/full_stage_r02.ncu-rep gpurun_out/
python tools/bench_models.py 1048576 > gpurun_out/bench_models_r02.log 2>&1; cp gpurun_out/bench_models.json gpurun_out/bench_models_r02.json
python tools/parity_at_scale.py 262144 2 > gpurun_out/parity_r02_cfg2.json 2>/dev/null
python tools/parity_at_scale.py 65536 3 > gpurun_out/parity_r02_cfg3.json 2>/dev/null
du -sh gpurun_out
