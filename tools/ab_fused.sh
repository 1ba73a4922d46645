V=depth-fusion-in-transformer-based-video-object-detection_b200/variants/libmsda_b200_fb0.so
timeout 600 python -m pytest tests/test_gpu_fused.py tests/test_gpu_modules.py tests/test_gpu_graphed_step.py -q -m gpu 2>&1 | tail -2
for i in 1 2; do for lib in "" $V; do
MSDA_B200_LIB=$lib timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('LIB=$lib'[-10:], 'train', round(d['train_step']['ms_per_step'],3), 'eager', round(d['train_step']['eager_ms_per_step'],2), 'cf', round(d['encoder_cross_fusion_layer']['coco_4_levels']['fwd_bwd_ms'],3), round(d['encoder_cross_fusion_layer']['shipped_1_level_50x84']['fwd_bwd_ms'],3))
"
done; done
