cd $GRAFT_REPO_ROOT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29711 tools/exchange_time.py 2>gpurun_out/c13.err | tail -1 | tee gpurun_out/c13_exchange_n2.json; tail -3 gpurun_out/c13.err
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run2() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:3}" 2> gpurun_out/$2.err | tail -1 > gpurun_out/$2.json; echo "bench rc=$?"; python -c "import json;d=json.load(open('gpurun_out/$2.json'));print(d['value'],d['ms_per_step'],d['config']['exchange'],d.get('exchange_check'),d['e2e']['value'],d['windows_ms'])"; tail -3 gpurun_out/$2.err; }
run2 29701 c13_d_n2 --steps 50 --warmup 5 --no-cpu-baseline
NNUE_EXCHANGE_OVERLAP=0 run2 29702 c13_d_n2_nooverlap --steps 50 --warmup 5 --no-cpu-baseline
