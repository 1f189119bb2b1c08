cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests/test_gpu_float.py tests/test_gpu_graph.py -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/f3_tests.log
b() { name=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json;d=json.load(open('gpurun_out/$name.json'));print('$name',d['value'],d['ms_per_step'],d.get('stages_ms'))" || tail -5 gpurun_out/$name.err; }
b f3_l1k --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int
b f3_l1k_noinline --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int --opt gemm_inline_a=0
b f3_small --workload imagenet_small_b16384 --steps 5 --warmup 3 --no-cpu-baseline --no-module-api --no-e2e --no-int
b f3_large --workload imagenet_large_b4096 --steps 5 --warmup 3 --no-cpu-baseline --no-module-api --no-e2e --no-int
