set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_float.py -m gpu -x -q -k "I_small or I_large or L1_128 or wide_head or variants" > gpurun_out/c7_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/c7_tests.log
timeout 400 python bench.py --workload imagenet_large_b4096 --steps 5 --warmup 3 --windows 3 --no-cpu-baseline --no-module-api > gpurun_out/c7_bench_large.json 2> gpurun_out/c7_bench_large.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c7_bench_large.json'));print(d['value'],d['ms_per_step'],d['stages_ms'])"; tail -3 gpurun_out/c7_bench_large.err
timeout 400 python bench.py --workload imagenet_small_b16384 --steps 5 --warmup 3 --windows 3 --no-cpu-baseline --no-module-api > gpurun_out/c7_bench_small.json 2> gpurun_out/c7_bench_small.err; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/c7_bench_small.json'));print(d['value'],d['ms_per_step'],d['stages_ms'])"; tail -3 gpurun_out/c7_bench_small.err
CMD="python tools/profile_step.py --workload imagenet_small_b16384 --batch 4096 --steps 2"
$CMD > gpurun_out/c7_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_bwd_rows -c 1 -o gpurun_out/c7_rows -f $CMD > gpurun_out/c7_ncu.log 2>&1; echo "ncu rc=$?"
