# 2-GPU step time under a few environment toggles (run through gpurun --gpus 2)
run() { tag=$1; shift; env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-eager 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', d['ms_per_step'])"; }
run base A=1
run no_early DATR_EARLY_REDUCE=0
run nch4 NCCL_MAX_NCHANNELS=4
run nch4_noearly NCCL_MAX_NCHANNELS=4 DATR_EARLY_REDUCE=0
run nch2 NCCL_MAX_NCHANNELS=2
