timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
for k in 2 1; do echo "== timeline kernel $k"; PPSCI_B200_DEBUG_KERNEL=$k timeout 120 python scripts_timeline.py 2>&1 | sed -n 2,19p; done
timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['class_ms'])"
