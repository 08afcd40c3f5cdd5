mkdir -p gpurun_out/r2/mg2
timeout 400 python -m pytest tests/test_multi_gpu.py -q --timeout 200 2>&1 | tail -2
timeout 120 python tools/step_cost.py 2 1000000 cfg3 2>&1 | tail -2
timeout 120 python tools/step_cost.py 2 62500000 cfg3 2>&1 | tail -2
timeout 120 python tools/step_cost.py 2 4000000 cfg4 2>&1 | tail -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2/mg2/bench2.json 2> gpurun_out/r2/mg2/bench2.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2/mg2/bench2.json') if x.startswith('{')][-1]
d=json.loads(l)
def show(d,name): print(name, "value %.3e"%d['value'], "ms", round(d.get('ms_per_step'),3), "frac", round(d['roofline']['frac'],3), "e2e %.3e"%d['e2e']['value'], "launches", d.get('gpu_launches'))
show(d,'cfg3')
for k,v in d.get('sub',{}).items(): show(v,k)
PY
