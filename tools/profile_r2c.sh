set -x
timeout 300 python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > /dev/null 2>&1 || exit 1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_ -s 6 -c 4 -o gpurun_out/prof_r2c_f32 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/prof_r2c_f32.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:msda_ -s 6 -c 4 -o gpurun_out/prof_r2c_bf16 -f python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --dtype bf16 > gpurun_out/prof_r2c_bf16.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/launches_r2c.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:linear_tf32x3 -s 2 -c 1 -o gpurun_out/prof_r2c_tf32x3 -f python tools/run_tf32x3_once.py 256 256 > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:linear_tf32x3 -s 2 -c 1 -o gpurun_out/prof_r2c_tf32x3_k1024 -f python tools/run_tf32x3_once.py 256 1024 > /dev/null 2>&1
ls -la gpurun_out/prof_r2c_*.ncu-rep
