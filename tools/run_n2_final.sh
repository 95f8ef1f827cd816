cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/r02w_bench_n2.err | grep "^{" > gpurun_out/r02w_bench_n2.json
tail -2 gpurun_out/r02w_bench_n2.err
