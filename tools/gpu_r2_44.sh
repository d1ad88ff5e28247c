# closing tree on 8 B200: one process per GPU (torchrun), 4 pairs per GPU, no data-path collective; decryption leg off to bound the box time
O=gpurun_out/r2l; mkdir -p $O
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 3 --warmup 3 --no-dec > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "rc=$?" >> $O/bench_8gpu.err
tail -c 300 $O/bench_8gpu.json
