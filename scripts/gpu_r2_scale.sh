#!/bin/bash
# round 2: the multi-GPU path on N GPUs of one box (gpurun --gpus N -- 'bash scripts/gpu_r2_scale.sh N [workload] [steps]')
cd $GRAFT_REPO_ROOT
N=${1:-2}; W=${2:-chr1}; K=${3:-5}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_scale_topo_n$N.txt 2>&1
python -c "import bench; bench.make_workload('$W', None)" > /dev/null 2>&1   # build + cache the workload once, outside torchrun
NCCL_DEBUG=WARN timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --workload $W --steps $K --warmup 3 > gpurun_out/r2_scale_${W}_n$N.json 2> gpurun_out/r2_scale_${W}_n$N.err
echo "scale N=$N rc=$?"
tail -c 1500 gpurun_out/r2_scale_${W}_n$N.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_scale_${W}_n$N.json").read().strip().splitlines()[-1])
    print({k: d[k] for k in ("n_gpus", "ms_per_step", "value", "byte_identical_to_oracle")}, d["e2e"]["ms_per_step"], d["phases_ms_rank0"])
except Exception as e:
    print("no bench line:", e)
PY
