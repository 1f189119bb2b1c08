cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/s3_tests.log
timeout 300 python tools/int_time.py 1 16 256 2048 16384 65536 2>&1 | tee gpurun_out/s3_int_time.log | grep "^B="
b() { name=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json;d=json.load(open('gpurun_out/$name.json'));print('$name',d['value'],d['ms_per_step'],d.get('stages_ms'))" || tail -5 gpurun_out/$name.err; }
b s3_l1k --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int
b s3_l1k_nomid --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int --opt head_mid=0 --opt head_pair_epilogue=0
b s3_d --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int
