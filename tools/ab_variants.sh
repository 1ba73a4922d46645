#!/bin/bash
# A/B the kernel build variants under <pkg>/variants/ on the bench workload (fp32 and bf16).
PKG="depth-fusion-in-transformer-based-video-object-detection_b200"
for lib in $PKG/variants/libmsda_*.so; do
  for dt in ${DTYPES:-f32 f32 bf16}; do
    MSDA_B200_LIB=$PWD/$lib python bench.py --steps ${STEPS:-100} --warmup 3 --no-cpu-baseline --no-extras --dtype $dt 2>&1 | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['per_step_ms']; print('$(basename $lib) $dt fwd mean %.4f med %.4f min %.4f max %.4f | bwd mean %.4f med %.4f min %.4f max %.4f | qps=%.3e' % (d['fwd_ms'], p['fwd_median'], p['fwd_min'], p['fwd_max'], d['bwd_ms'], p['bwd_median'], p['bwd_min'], p['bwd_max'], d['value']))"
  done
done
