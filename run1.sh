timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['class_ms'])"
PPSCI_B200_NO_THINV=1 timeout 300 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_err.log | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('no thinv', {k:d[k] for k in ('value','ms_per_step')}); print(d['roofline']['class_ms'])"
