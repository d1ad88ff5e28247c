set -x
O=gpurun_out/r2; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 2 --warmup 3 > $O/bench38_2gpu.json 2> $O/bench38_2gpu.err; echo "rc=$?" >> $O/bench38_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > $O/bench38_2gpu_ref.json 2> $O/bench38_2gpu_ref.err; echo "rc=$?" >> $O/bench38_2gpu_ref.err
timeout 600 python -m pytest tests/test_multi_rank.py -m gpu -x -q > $O/t38_multi.log 2>&1; echo "rc=$?" >> $O/t38_multi.log
