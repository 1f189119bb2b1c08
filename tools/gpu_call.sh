cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_float.py -m gpu -x -q 2>&1 | tail -2
for U in 1 2 3 4; do timeout 300 python tools/ft_compare.py --workload imagenet_large_b4096 --batch 1 --only fwd --forms gather --reps 15 --opt ft_gather_units=$U 2>/dev/null | python -c "import json,sys;d=json.load(sys.stdin);print('units $U', d['calls']['fwd']['gather'])"; done
b() { name=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json;d=json.load(open('gpurun_out/$name.json'));print('$name',d['value'],d['ms_per_step'],d['stages_ms'])" || tail -5 gpurun_out/$name.err; }
b c16_large --workload imagenet_large_b4096 --steps 5 --warmup 3 --windows 5 --no-cpu-baseline --no-module-api --no-e2e
b c16_small --workload imagenet_small_b16384 --steps 5 --warmup 3 --windows 5 --no-cpu-baseline --no-module-api --no-e2e
b c16_d1k --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e
