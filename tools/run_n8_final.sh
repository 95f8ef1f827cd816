cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 2> gpurun_out/r02u_bench_n8.err | grep "^{" > gpurun_out/r02u_bench_n8.json
tail -3 gpurun_out/r02u_bench_n8.err
timeout 300 python -m pytest tests -m gpu -q -s -k "two_contexts" 2>&1 | tail -3 > gpurun_out/r02u_two_contexts.log; cat gpurun_out/r02u_two_contexts.log
