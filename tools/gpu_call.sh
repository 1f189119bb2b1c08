cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=${NGPU:-1}
if [ "$N" = "1" ]; then
b() { name=$1; shift; timeout 600 python bench.py "$@" 2> gpurun_out/$name.err | tail -1 > gpurun_out/$name.json; python -c "import json;d=json.load(open('gpurun_out/$name.json'));print('$name',d['value'],d['ms_per_step'])" || tail -5 gpurun_out/$name.err; }
b f2_l1k --workload real_cifar_l1_1024_b16384 --steps 20 --warmup 5 --no-cpu-baseline --no-module-api --no-e2e --no-int
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_bwd|extract_fwd|ft_bitgemm|ft_gbin|head_train|input_bwd_fold|umma_" -s 28 -c 14 -o gpurun_out/f2_d_step -f python tools/profile_step.py --steps 4 > gpurun_out/f2_ncu_full.log 2>&1; tail -2 gpurun_out/f2_ncu_full.log
else
for W in default_cifar_b16384 imagenet_small_b16384; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --workload $W --steps 20 --warmup 5 --no-cpu-baseline --no-int --no-e2e --no-module-api 2> gpurun_out/f2_${W}_n$N.err | tail -1 > gpurun_out/f2_${W}_n$N.json; python -c "import json;d=json.load(open('gpurun_out/f2_${W}_n$N.json'));print('$W',d['n_gpus'],d['value'],d['ms_per_step'],d.get('exchange_check',{}).get('identical_on_all_ranks'))" || tail -3 gpurun_out/f2_${W}_n$N.err
done
fi
