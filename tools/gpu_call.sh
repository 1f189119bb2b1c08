cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_float.py -m gpu -x -q 2>&1 | tail -3
b() { name=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json;d=json.load(open('gpurun_out/$name.json'));print('$name',d['value'],d['ms_per_step'],d['stages_ms'])" || tail -5 gpurun_out/$name.err; }
b c21_small --workload imagenet_small_b16384 --steps 5 --warmup 3 --windows 5 --no-cpu-baseline --no-module-api --no-e2e
b c21_large --workload imagenet_large_b4096 --steps 5 --warmup 3 --windows 5 --no-cpu-baseline --no-module-api --no-e2e
