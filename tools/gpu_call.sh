cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_int.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s2_int_tests.log
timeout 300 python tools/int_time.py 2>&1 | tee gpurun_out/s2_int_time.log | grep "^B="
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"q_conv|ft_bitgemm|q_infer" -c 6 -o gpurun_out/s2_int_b16384 -f python tools/int_time.py --profile 16384 > gpurun_out/s2_ncu.log 2>&1; tail -3 gpurun_out/s2_ncu.log
