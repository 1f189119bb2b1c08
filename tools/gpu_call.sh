cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/f4_multi_n$N.log
for W in default_cifar_b16384; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 30 --warmup 5 --no-cpu-baseline --no-int --no-e2e --no-module-api 2> gpurun_out/f4_${W}_n$N.err | tail -1 > gpurun_out/f4_${W}_n$N.json; python -c "import json;d=json.load(open('gpurun_out/f4_${W}_n$N.json'));print('$W',d['n_gpus'],d['value'],d['ms_per_step'],d.get('exchange_check',{}).get('identical_on_all_ranks'))" || tail -3 gpurun_out/f4_${W}_n$N.err
done
