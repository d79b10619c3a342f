set -x
export NCU_POINTS=131072 NCU_STEPS=2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python tests/tools/ncu_target.py > gpurun_out/r02_launches.log 2>&1
export NCU_POINTS=65536
for spec in "k_tc2_fwd 7" "k_fused_dx 1" "k_tc2_dw 7"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f -o gpurun_out/r02_$1 python tests/tools/ncu_target.py > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page details > gpurun_out/r02_ncu_$1_details.txt 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>&1
done
timeout 900 python examples/laplace/laplace2d.py --epochs 20000 --output_dir /tmp/out_laplace --result_json gpurun_out/laplace2d_result.json > gpurun_out/laplace2d.log 2>&1
tail -3 gpurun_out/laplace2d.log
timeout 600 python -m pytest tests/test_gpu_train.py -q -m gpu --timeout=600 2>&1 | tail -5
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/bench_r2e_cfg5.json 2> gpurun_out/bench_r2e_cfg5.err; tail -c 600 gpurun_out/bench_r2e_cfg5.json
