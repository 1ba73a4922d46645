# A/B of the backward kernel: parity suites, then bench timings of the in-tree build and of every variant library given
P=depth-fusion-in-transformer-based-video-object-detection_b200
timeout 900 python -m pytest tests/test_gpu_op_parity.py tests/test_gpu_fused.py tests/test_gpu_tc_forward.py -q -m gpu -x 2>&1 | tail -5
for lib in "" "$@"; do
 for dt in f32 bf16; do for dist in grid init random; do
  MSDA_B200_LIB=$lib timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --dtype $dt --dist $dist 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('LIB=$lib', '$dt', '$dist', 'fwd', round(d['fwd_ms'],4), 'bwd', round(d['bwd_ms'],4), 'value', round(d['value']/1e6,2))
"
 done; done
done
