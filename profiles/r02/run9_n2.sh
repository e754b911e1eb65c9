# 2-GPU session: multi-device tests on real devices, the default bench line at N = 1 and N = 2 (torchrun, NCCL)
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_frontend.py -m gpu -q -x 2>&1 | tail -8
timeout 600 python bench.py --steps 5 --warmup 3 --no-workloads > gpurun_out/b9_n1.json 2> gpurun_out/b9_n1.err; tail -c 300 gpurun_out/b9_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/b9_n2.json 2> gpurun_out/b9_n2.err; tail -c 600 gpurun_out/b9_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload c4_job_loam > gpurun_out/b9_n2_loam.json 2> gpurun_out/b9_n2_loam.err; tail -c 600 gpurun_out/b9_n2_loam.err
ls -la gpurun_out | grep b9_
