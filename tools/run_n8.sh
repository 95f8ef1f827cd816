cd $GRAFT_REPO_ROOT
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
tail -3 gpurun_out/r02_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -3 gpurun_out/r02_bench_n2.err
python bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/r02_bench_ref_n8.json 2>/dev/null
timeout 300 python -m pytest tests -m gpu -q -s -k "two_contexts" 2>&1 | tail -3 > gpurun_out/r02_two_contexts.log; cat gpurun_out/r02_two_contexts.log
python - <<'PY'
import json
for f in ("gpurun_out/r02_bench_n8.json","gpurun_out/r02_bench_n2.json","gpurun_out/r02_bench_ref_n8.json"):
    try:
        j=json.load(open(f)); print(f, j["value"], j["ms_per_step"], j.get("e2e"), j.get("allreduce_check"), j.get("tail_ms"), {k:(v["value"],v["ms_per_step"]) for k,v in j.get("series",{}).items()})
    except Exception as e: print(f, "ERR", e)
PY
