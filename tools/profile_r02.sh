# Round-2 measurement pass: tests, bench lines (ours + reference arm), launch list, ncu --set full of the fit kernels, other models, parity at scale.
set -x
mkdir -p gpurun_out
R=gpurun_out
python -m pytest tests -m gpu -q > $R/r2_tests.log 2>&1; echo "tests rc=$?" >> $R/r2_tests.log; tail -4 $R/r2_tests.log
python bench.py > $R/bench_r02.json 2> $R/bench_r02.err; tail -3 $R/bench_r02.err
python bench.py --impl reference --steps 2 --warmup 1 > $R/bench_r02_reference.json 2>> $R/bench_r02.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $R/launches_r02.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-configs > $R/launches_bench.log 2>&1
T=/tmp/ncu; mkdir -p $T
CMD1="ncu --set full --clock-control none --import-source on -k regex:k_noddi_stage -s 12 -c 4 python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-pipeline --no-configs"
ncu --set full --clock-control none --import-source on -k regex:k_noddi_stage -s 12 -c 4 -o $T/full_stage_r02 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-pipeline --no-configs > $R/full_stage.log 2>&1
python tools/ncu_to_json.py $T/full_stage_r02.ncu-rep "$CMD1" 1048576 > $R/ncu_full_r02_noddi_stage_kernels.json
ncu -i $T/full_stage_r02.ncu-rep --page source --csv --print-source cuda,sass > $T/src.csv 2>/dev/null
python tools/ncu_src_lines.py $T/src.csv 40 > $R/ncu_r02_noddi_stage_by_line.txt
for M in CylinderZeppelinBall5:k_lasso_batched:czb SANDI4:k_lasso_small:sandi FreeWater1:k_lasso_batched:freewater; do
  IFS=: read A K N <<< "$M"
  ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 -o $T/full_$N -f python tools/bench_models.py 1048576 $A > $R/full_$N.log 2>&1
  python tools/ncu_to_json.py $T/full_$N.ncu-rep "ncu --set full --clock-control none --import-source on -k regex:$K -s 2 -c 1 python tools/bench_models.py 1048576 $A" 1048576 > $R/ncu_full_r02_${K}_$N.json
done
python tools/bench_models.py 1048576 > $R/bench_models_r02.log 2>&1; cp $R/bench_models.json $R/bench_models_r02.json
python tools/parity_at_scale.py 262144 2 > $R/parity_r02_cfg2.json 2>/dev/null
python tools/parity_at_scale.py 65536 3 > $R/parity_r02_cfg3.json 2>/dev/null
cat $R/parity_r02_cfg2.json $R/parity_r02_cfg3.json
du -sh $R
