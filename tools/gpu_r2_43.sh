# closing tree on 2 B200: one process per GPU (torchrun), 4 pairs per GPU, no data-path collective
O=gpurun_out/r2k; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "rc=$?" >> $O/bench_2gpu.err
tail -c 600 $O/bench_2gpu.json
