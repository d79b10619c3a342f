timeout 600 python tests/tc_debug_case.py 2>&1 | grep -v "mask=7\|mask=15" | tail -12
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -5
for k in 1 2; do echo "== timeline kernel $k"; PPSCI_B200_DEBUG_KERNEL=$k timeout 120 python scripts_timeline.py 2>&1 | sed -n 2,19p; done
for m in 31; do echo "== bench mask $m"; PPSCI_B200_TC_MASK=$m timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['class_ms'])"; done
