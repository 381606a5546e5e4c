TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
$TR 29541 tools/h2d_contention.py 2>/dev/null | tail -1
for L in 2 3; do
  $TR 2955$L bench.py --gpus 8 --steps 4 --warmup 3 --lanes $L 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $L', round(d['value'],1), d['e2e']['value'], d['e2e']['ms_per_step'])"
done
PB200_BLOCKING_SYNC=1 $TR 29561 bench.py --gpus 8 --steps 4 --warmup 3 --lanes 4 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes 4 blocking', round(d['value'],1), d['e2e']['value'], d['e2e']['ms_per_step'])"
