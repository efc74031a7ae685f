python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 profiles/k1_multi_gpu_probe.py 2>&1 | grep "^rank" | tee gpurun_out/final_k1_order.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-cpu > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo exit=$?
python - <<'PY'
import json
d=json.load(open('gpurun_out/final_bench_n2.json'))
print(round(d['value']), d['ms_per_step'], [round(x,1) for x in d['roofline']['us_per_launch_by_rank']])
PY
