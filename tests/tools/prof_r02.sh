# Round-2 evidence run (under gpurun, one B200): GPU tests, benches of all five BASELINE configurations, per-config ncu
# launch lists, full ncu captures of the three dominant kernels with their source pages, timelines of the fused kernels.
# Outputs land in gpurun_out/; the reduced versions committed under profiles/ are listed in profiles/README.md.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout=600 2>&1 | tail -8 > gpurun_out/gputests.txt; tail -4 gpurun_out/gputests.txt
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
for c in 1 2 4 5; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_cfg$c.json 2> gpurun_out/bench_cfg$c.err
done
for c in 1 2 3 4 5; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_cfg$c.csv python tests/tools/ncu_config.py $c > gpurun_out/r02_launches_cfg$c.log 2>&1
done
export NCU_POINTS=65536 NCU_STEPS=2
for spec in "k_fused_fwd16 1" "k_fused_dx 1" "k_tc2_dw 7"; do
  set -- $spec
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 --launch-skip $2 -c 1 -f -o gpurun_out/ncu_$1 python tests/tools/ncu_target.py > gpurun_out/ncu_$1.log 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page details > gpurun_out/ncu_$1_details.txt 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page raw --csv > gpurun_out/ncu_$1_raw.csv 2>&1
  ncu -i gpurun_out/ncu_$1.ncu-rep --page source --csv > gpurun_out/ncu_$1_source.csv 2>&1
  python tests/tools/ncu_source_hotspots.py gpurun_out/ncu_$1_source.csv 0 99999 0.5 > gpurun_out/ncu_$1_stall_hotspots.txt
  rm -f gpurun_out/ncu_$1.ncu-rep   # 10-25 MB each; gpurun_out/ travels back only below 64 MiB
done
timeout 300 python tests/tools/timeline_fused.py > gpurun_out/timeline_fwd16.txt 2>&1
PPSCI_B200_DEBUG_KERNEL=4 timeout 300 python tests/tools/timeline_fused.py > gpurun_out/timeline_fused_dx.txt 2>&1
timeout 600 python examples/laplace/laplace2d.py --epochs 20000 --output_dir /tmp/out_laplace --result_json gpurun_out/laplace2d_result.json > gpurun_out/laplace2d.log 2>&1
