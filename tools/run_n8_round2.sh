T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$T --nproc-per-node 8 --master-port 29701 tools/check_dp.py > gpurun_out/r2_check_dp_n8.log 2>&1; echo rc=$? >> gpurun_out/r2_check_dp_n8.log
$T --nproc-per-node 8 --master-port 29702 bench.py --gpus 8 --mode gan --steps 3 --warmup 2 > gpurun_out/r2_gan_n8.json 2> gpurun_out/r2_gan_n8.err; echo rc=$? >> gpurun_out/r2_gan_n8.err
F="--no-decode --no-eager --no-cpu-baseline --no-profile --steps 5 --warmup 3"
$T --nproc-per-node 2 --master-port 29703 bench.py --gpus 2 --batch 2048 --micro-bars 512 $F > gpurun_out/r2_c3_n2.json 2> gpurun_out/r2_c3_n2.err
$T --nproc-per-node 4 --master-port 29704 bench.py --gpus 4 --batch 1024 --micro-bars 512 $F > gpurun_out/r2_c3_n4.json 2> gpurun_out/r2_c3_n4.err
$T --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --batch 512 $F > gpurun_out/r2_c3_n8.json 2> gpurun_out/r2_c3_n8.err
python bench.py --gpus 1 --batch 4096 --micro-bars 512 $F > gpurun_out/r2_c3_n1.json 2> gpurun_out/r2_c3_n1.err
