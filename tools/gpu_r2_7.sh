set -x
mkdir -p gpurun_out/r2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2/bench7_2gpu.json 2> gpurun_out/r2/bench7_2gpu.err; echo "rc=$?" >> gpurun_out/r2/bench7_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2/bench7_2gpu_ref.json 2> gpurun_out/r2/bench7_2gpu_ref.err; echo "rc=$?" >> gpurun_out/r2/bench7_2gpu_ref.err
