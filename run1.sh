timeout 600 python tests/tc_debug_case.py 2>&1 | grep "ns_f32_tc_256" | tail -4
timeout 900 python -m pytest tests/test_gpu_tc.py -x -q 2>&1 | tail -4
echo "== timeline dW"; PPSCI_B200_DEBUG_KERNEL=0 timeout 120 python scripts_timeline.py 2>&1 | sed -n 2,14p
timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['class_ms'])"
