#!/bin/bash
# tools/gpu_run.sh -- the ONE script behind every gpurun call of this repository.  Usage (from the repo root):
#     gpurun --timeout 2400 -- 'bash tools/gpu_run.sh <tag> <stage> [<stage> ...]'
# Everything a stage writes goes to gpurun_out/<tag>_*; copy what should be judged into profiles/.
# Stages:
#   tests        python -m pytest tests -m gpu -q                        (the whole GPU suite)
#   tests:<expr> python -m pytest -q <expr>                              (e.g. tests:tests/test_conv_gpu.py)
#   smoke        __graft_entry__.smoke()
#   bench[:workload]   python bench.py --workload <dino|dino5|teacher|msda> (default dino), 10 steps / 3 warm-up
#   bench2       the default bench on 2 GPUs through torch.distributed.run  (call gpurun with --gpus 2)
#   refarm       python bench.py --impl reference --steps 2 --warmup 1   (the CPU arm on the box's host cores)
#   reference[:4scale|5scale[:init|seeded]]   tools/bench_reference_gpu.py: the reference model on this GPU + full-size parity
#   micro_msda | micro_linear | micro_conv    the per-kernel micro-benchmarks against the reference CUDA op / cuBLAS / cuDNN
#   kernels      torch.profiler kernel table + idle-gap report of the graph-replayed step
#   timeline     per-segment device timeline of the graph-replayed step (tools/timeline_report.py)
#   launches     ncu launch list (gpu__time_duration.sum) of the msda workload and of the hand-written kernels of the dino step
#   ncu_msda     ncu --set full of the MSDeformAttn kernels at the config-2 encoder / decoder calls
#   probes       the standalone probes under tools/probes (built here, run there)
#   toggles      default bench with / without the in-graph gradient sinks and the side-branch weight gradients
set -u
TAG=${1:?tag}; shift
mkdir -p gpurun_out
for stage in "$@"; do
  name=${stage%%:*}; arg=""; [[ "$stage" == *:* ]] && arg=${stage#*:}
  echo "== [$TAG] $stage"
  case "$name" in
    tests)
      if [ -z "$arg" ]; then timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_gpu_tests.txt 2>&1
      else timeout 1500 python -m pytest -q $arg > gpurun_out/${TAG}_gpu_tests.txt 2>&1; fi
      echo "rc=$?"; tail -6 gpurun_out/${TAG}_gpu_tests.txt | cut -c1-220 ;;
    smoke) timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.txt 2>&1; echo "rc=$?"; tail -3 gpurun_out/${TAG}_smoke.txt ;;
    bench)
      wl=${arg:-dino}
      timeout 1200 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${TAG}_bench_${wl}.json 2> gpurun_out/${TAG}_bench_${wl}.err
      echo "rc=$?"; cut -c1-240 gpurun_out/${TAG}_bench_${wl}.json ;;
    bench2)
      timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
        > gpurun_out/${TAG}_bench_dino_2gpu.json 2> gpurun_out/${TAG}_bench_dino_2gpu.err
      echo "rc=$?"; cut -c1-240 gpurun_out/${TAG}_bench_dino_2gpu.json ;;
    refarm) timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference_arm.json 2> /dev/null; echo "rc=$?"; cut -c1-300 gpurun_out/${TAG}_bench_reference_arm.json ;;
    reference)
      cfg=${arg%%:*}; cfg=${cfg:-4scale}; w=init; [[ "$arg" == *:* ]] && w=${arg#*:}
      timeout 1500 python tools/bench_reference_gpu.py --config $cfg --weights $w --steps 5 --out gpurun_out/${TAG}_reference_gpu_${cfg}_${w}.json \
        > gpurun_out/${TAG}_reference_gpu_${cfg}_${w}.log 2>&1
      echo "rc=$?"; grep -E "^\[(parity|reference|ours)" gpurun_out/${TAG}_reference_gpu_${cfg}_${w}.log | cut -c1-400 ;;
    micro_msda) timeout 900 python tools/microbench_msda.py --fused > gpurun_out/${TAG}_msda_microbench.txt 2>&1; cat gpurun_out/${TAG}_msda_microbench.txt | cut -c1-160 ;;
    micro_linear) timeout 900 python tools/bench_linear.py > gpurun_out/${TAG}_bench_linear.txt 2>&1; cut -c1-200 gpurun_out/${TAG}_bench_linear.txt ;;
    micro_conv) timeout 900 python tools/bench_conv_backward.py > gpurun_out/${TAG}_bench_conv.txt 2>&1; cut -c1-200 gpurun_out/${TAG}_bench_conv.txt ;;
    kernels)
      timeout 600 python tools/profile_dino.py > /dev/null 2>&1; cp gpurun_out/dino_step_kernels.txt gpurun_out/${TAG}_dino_step_kernels_graphs.txt
      timeout 600 python tools/gap_report.py > /dev/null 2>&1; cp gpurun_out/dino_step_gaps.txt gpurun_out/${TAG}_dino_step_gaps.txt
      head -40 gpurun_out/${TAG}_dino_step_kernels_graphs.txt | cut -c1-180 ;;
    timeline)
      timeout 600 python tools/timeline_report.py > gpurun_out/${TAG}_timeline.log 2>&1; echo "rc=$?"
      cp gpurun_out/dino_step_timeline.txt gpurun_out/${TAG}_dino_step_timeline.txt; head -40 gpurun_out/${TAG}_dino_step_timeline.txt | cut -c1-170 ;;
    launches)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${TAG}_launches_msda.csv \
        python bench.py --workload msda --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu_msda.log 2>&1
      DATR_GRAPHS=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none \
        -k "regex:msda_|linear_tf32|wgrad_tf32|layernorm256|softmax_|conv3x3|colsum|zero_masked|ema_update|attn_fwd|attn_bwd|adamw_step|lsa_kernel|gn_fwd|gn_bwd|sine_embed|pos_embed|bn_relu_maxpool" --launch-skip 1500 -c 2500 --csv \
        --log-file gpurun_out/${TAG}_launches_dino_handwritten.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/${TAG}_bench_under_ncu_dino.log 2>&1
      python tools/launch_list_summary.py gpurun_out/${TAG}_launches_msda.csv gpurun_out/${TAG}_msda_workload_launches.txt "ncu launch list of: python bench.py --workload msda --steps 2 --warmup 3"
      python tools/launch_list_summary.py gpurun_out/${TAG}_launches_dino_handwritten.csv gpurun_out/${TAG}_dino_step_handwritten_launches.txt "ncu launch list (hand-written kernels, eager) of: python bench.py --steps 1 --warmup 3"
      tail -25 gpurun_out/${TAG}_dino_step_handwritten_launches.txt | cut -c1-160 ;;
    ncu_msda)
      timeout 900 ncu --set full --clock-control none --import-source on -k regex:msda_ -s 4 -c 4 -f -o gpurun_out/${TAG}_msda python tools/ncu_target.py > gpurun_out/${TAG}_ncu_msda.log 2>&1
      echo "rc=$?"; tail -2 gpurun_out/${TAG}_ncu_msda.log ;;
    probes)
      for p in gemm_2cta_probe msda_tile_probes tmem_ld_layout_probe; do
        [ -x build/$p ] && { timeout 180 build/$p > gpurun_out/${TAG}_$p.txt 2>&1; echo "$p rc=$?"; tail -4 gpurun_out/${TAG}_$p.txt; }
      done ;;
    toggles)
      # A/B of the graph-side switches on the default workload (short runs, no CPU baseline, no eager leg)
      # TOGGLES="name:ENV=v,ENV2=v name2:..." overrides the list
      for cfg in ${TOGGLES:-new: noside:DATR_WGRAD_SIDE_ROWS=0 base:DATR_WGRAD_SIDE_ROWS=0,DATR_GRAPH_SINKS=0}; do
        nm=${cfg%%:*}; envs=${cfg#*:}; envs=${envs//,/ }
        env $envs timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/${TAG}_toggle_${nm}.json 2> gpurun_out/${TAG}_toggle_${nm}.err
        echo "$nm rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_toggle_${nm}.json')); print(round(d['ms_per_step'],3),'ms', round(d['e2e']['ms_per_step'],3),'ms e2e, loss',d.get('loss'))" 2>&1 | tail -1)"
      done ;;
    *) echo "unknown stage $stage"; exit 2 ;;
  esac
done
