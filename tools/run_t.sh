mkdir -p gpurun_out/r2/mg3
timeout 300 python -m pytest tests/test_multi_gpu.py -q --timeout 200 2>&1 | tail -2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2/mg3/bench2.json 2> gpurun_out/r2/mg3/bench2.err
python - <<'PY'
import json
l=[x for x in open('gpurun_out/r2/mg3/bench2.json') if x.startswith('{')][-1]
d=json.loads(l)
def show(d,name): print(name, "value %.3e"%d['value'], "ms", round(d.get('ms_per_step'),3), "frac", round(d['roofline']['frac'],3), "e2e %.3e"%d['e2e']['value'], "launches", d.get('gpu_launches'), d.get('check',{}).get('count_ok'), d.get('check',{}).get('all_ranks_identical'))
show(d,'cfg3')
for k,v in d.get('sub',{}).items(): show(v,k)
PY
