# weak-scaling sweep on one 8-GPU box (run through gpurun --gpus 8): N = 1, 2, 4, 8
for n in 1 2 4 8; do
  if [ $n = 1 ]; then
    timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager 2>/dev/null > gpurun_out/scale_n$n.json
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline --no-eager 2>/dev/null > gpurun_out/scale_n$n.json
  fi
  python -c "import json; d=json.load(open('gpurun_out/scale_n$n.json')); print('N=$n', round(d['ms_per_step'],2), 'ms', round(d['value'],1), 'images/s')"
done
