set -x
timeout 1500 python -m pytest tests -q -m gpu --timeout=900 2>&1 | tail -8 > gpurun_out/gputests_r2b.txt; tail -4 gpurun_out/gputests_r2b.txt
python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; tail -2 gpurun_out/bench_r2f.err
PPSCI_B200_TC_MASK=191 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r2f_m191.json 2> gpurun_out/bench_r2f_m191.err
timeout 900 python examples/laplace/laplace2d.py --epochs 20000 --output_dir /tmp/out_laplace --result_json gpurun_out/laplace2d_result.json > gpurun_out/laplace2d.log 2>&1
tail -2 gpurun_out/laplace2d.log
export NCU_POINTS=131072 NCU_STEPS=2
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv python tests/tools/ncu_target.py > gpurun_out/r02_launches.log 2>&1
export NCU_POINTS=65536
for spec in "k_fused_fwd16 1" "k_fused_dx 1" "k_tc2_dw 7"; do
  set -- $spec
  ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f -o gpurun_out/r02_$1 python tests/tools/ncu_target.py > gpurun_out/r02_ncu_$1.log 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page details > gpurun_out/r02_ncu_$1_details.txt 2>&1
  ncu -i gpurun_out/r02_$1.ncu-rep --page raw --csv > gpurun_out/r02_ncu_$1_raw.csv 2>&1
  rm -f gpurun_out/r02_$1.ncu-rep   # the report files are 10-25 MB each; gpurun_out/ travels back only below 64 MiB
done
python bench.py --config 5 --steps 5 --warmup 3 > gpurun_out/bench_r2f_cfg5.json 2> gpurun_out/bench_r2f_cfg5.err
python bench.py --config 2 --steps 8 --warmup 3 > gpurun_out/bench_r2f_cfg2.json 2> gpurun_out/bench_r2f_cfg2.err
