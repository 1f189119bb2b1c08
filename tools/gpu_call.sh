# One validation pass on a GPU box (run through gpurun): the parity suite, the headline bench line, the integer timing table.
cd ${GRAFT_REPO_ROOT:-.}
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) 2>&1 | tee gpurun_out/final_tests.log
timeout 900 python bench.py > gpurun_out/final_bench_d_n1.json 2> gpurun_out/final_bench_d_n1.err
python -c "import json;d=json.loads(open('gpurun_out/final_bench_d_n1.json').read().strip().splitlines()[-1]);print('D',d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline']['value'],d['int_inference']['device_sweep_graph_replay_samples_per_s'])" || tail -5 gpurun_out/final_bench_d_n1.err
