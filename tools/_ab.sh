timeout 600 python -m pytest tests/test_gpu_kernels.py -q -k "norm_block" > gpurun_out/r2_clt_t.log 2>&1; echo rc=$? >> gpurun_out/r2_clt_t.log
timeout 300 python bench.py --no-decode --no-eager --no-cpu-baseline --steps 15 --warmup 3 > gpurun_out/r2_clt256.json 2> gpurun_out/r2_clt256.err
