#!/bin/bash
# one multi-GPU box call: copy ceiling at 1/2/4/8 GPUs, the bench's e2e leg on all GPUs, BASELINE config 5 at
# full length, the multi-device parity tests.   scratch/multi_gpu_round.sh TAG NGPU
TAG=${1:-r02}
N=${2:-8}
O=gpurun_out
nvidia-smi topo -m > $O/${TAG}_topo.txt 2>&1; lscpu | grep -E "Model name|^CPU\(s\)|Socket|NUMA|Hypervisor" >> $O/${TAG}_topo.txt; free -g >> $O/${TAG}_topo.txt
scratch/pcie_probe_multi --gpus 1,2,4,8 > $O/${TAG}_pcie_probe_multi.txt 2>&1
cat $O/${TAG}_pcie_probe_multi.txt | cut -c1-200
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > $O/${TAG}_bench_n$N.json 2> $O/${TAG}_bench_n$N.err
python - <<PY
import json
try:
    l = json.loads(open("$O/${TAG}_bench_n$N.json").read().strip().splitlines()[-1])
    print("bench n=$N value", l["value"], "e2e", {k: l["e2e"][k] for k in ("value", "ceiling", "frac", "ms_per_step")}, "seq", l["e2e"]["sequential"])
except Exception as ex:
    print("bench n=$N failed", ex)
PY
tail -3 $O/${TAG}_bench_n$N.err
DBDE_B200_CHUNK_FRAMES=64 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-cpu-baseline > $O/${TAG}_bench_n${N}_chunk64.json 2> $O/${TAG}_bench_n${N}_chunk64.err
python -c "
import json
l = json.loads(open('$O/${TAG}_bench_n${N}_chunk64.json').read().strip().splitlines()[-1]); print('chunk64 e2e', l['e2e']['value'], l['e2e']['ceiling'])"
timeout 900 python scratch/stream_bench.py 10000 256 micro 4096 4096 > $O/${TAG}_stream_config5_n$N.txt 2>&1; tail -2 $O/${TAG}_stream_config5_n$N.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "shard or multi" > $O/${TAG}_pytest_multi.log 2>&1; tail -3 $O/${TAG}_pytest_multi.log
