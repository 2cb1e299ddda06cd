# 8-GPU evidence run (one box): multi-device tests, in-process deviceCount scaling (row bands per device swept), torchrun bench lines for C3 (weak), C4 / C5 (strong)
O=gpurun_out; T=${1:-r03w}
N=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 200 python -m pytest tests/test_multi_device.py -m gpu -x -q 2>&1 | tail -3 > $O/${T}_multi_pytest.txt
for B in 4 2 1; do echo "== bands per device $B"; SAILOR_PT_BANDS=$B timeout 200 python tools/multi_device_bench.py c3 3; done > $O/${T}_inprocess_c3.txt 2>&1
timeout 300 python bench.py --devices $N --steps 10 --warmup 3 --no-secondary > $O/${T}_bench_c3_devices$N.json 2> $O/${T}_bench_c3_devices$N.err
for W in ${WORKLOADS:-c3 c4 c5}; do
  S=10; WU=3; [ $W = c4 ] && S=3; [ $W = c5 ] && S=2 && WU=1
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $S --warmup $WU --workload $W > $O/${T}_bench_${W}_n$N.json 2> $O/${T}_bench_${W}_n$N.err
done
cat $O/${T}_multi_pytest.txt $O/${T}_inprocess_c3.txt
for f in $O/${T}_bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], 'n_gpus', d['n_gpus'], 'Mrays/s %.0f' % d['value'], 'ms %.1f' % d['ms_per_step'], 'e2e %.0f' % d['e2e']['value'], d['scaling'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
