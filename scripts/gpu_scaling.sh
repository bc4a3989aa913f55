# usage: bash scripts/gpu_scaling.sh N   (run under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
N=$1
for W in ecoli pangenome; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --workload $W > gpurun_out/scale_${W}_n$N.json 2>/dev/null
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 3 --workload $W 2>/dev/null | grep '^{' > gpurun_out/scale_${W}_n$N.json
  fi
  python -c "
import json
d=json.load(open('gpurun_out/scale_${W}_n$N.json'))
print('$W', d['n_gpus'], round(d['ms_per_step'],3), round(d['value']), round(d['e2e']['value']), d['byte_identical_to_oracle'], d['dijkstra']['kernel_ms_per_step'])"
done
