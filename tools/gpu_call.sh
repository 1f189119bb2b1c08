cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/exchange_parts.py 2> gpurun_out/s15_parts_n$N.err | tee gpurun_out/s15_parts_n$N.log | tail -1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 --no-cpu-baseline --no-int --no-e2e --no-module-api 2> gpurun_out/s15_d_n8.err | tail -1 > gpurun_out/s15_d_n8.json; python -c "import json;d=json.load(open('gpurun_out/s15_d_n8.json'));print(d['value'],d['ms_per_step'],d.get('exchange_check'))"
