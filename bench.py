#!/usr/bin/env python
"""bench.py -- throughput of the DATR/DINO data-parallel hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload dino|dino5|teacher|msda] [--gpu-baseline]

Metric (BASELINE.json): images/sec at 1333x800, batch_size 2 per GPU (a DA training step consumes
2 source + 2 target images per GPU), plus the MSDeformAttn HBM roofline.  One JSON
line on stdout (rank 0).  Workloads:

  msda  the MultiScaleDeformableAttention calls of one DINO-4scale DA training step at
        BASELINE.json configs[1] shapes: two transformer passes, each 6 encoder calls
        (N=2, Lq=S=22223) and 6 decoder calls (Lq=1100 with denoising queries in the source
        pass, 900 in the target pass), forward AND backward: 24 + 24 launches.
  dino  the full DINO-4scale ResNet-50 DA training step (datr_b200.models), same shapes: BASELINE.json configs[1]/[2]
        (default; the metric's configuration).
  dino5 the DINO-5scale step, batch_size 1/GPU, S = 89023 tokens: configs[3].
  teacher  the teacher-student mutual-learning step (EMA teacher eval pass + pseudo labels + student DA pass with the
        target-domain criterion + teacher EMA update): configs[4].

Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks.
Every call of a step reads its own buffers (the step cycles > 3 GB, far larger than the 126 MB L2).
The oracle (oracle/) is used here only as `cpu_baseline` / `--impl reference`.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

CFG2_LEVELS = [(100, 167), (50, 84), (25, 42), (13, 21)]   # 1333x800 through ResNet-50 C3..C5 + extra
M_HEADS, D_HEAD, N_POINTS = 8, 32, 4
IMAGES_PER_STEP_PER_GPU = 4        # batch_size 2 -> 2 source + 2 target images (util/misc.py:291-300)
MSDA_WORKLOAD = ("MSDeformAttn calls of one DINO-4scale DA training step, 1333x800, batch_size 2/GPU: "
                 "12 encoder (N=2,Lq=S=22223) + 6 decoder Lq=1100 + 6 decoder Lq=900, forward+backward, fp32")
FALLBACK_HBM_GBS = 6650.0          # /opt/skills/guides/B200_PROFILING.md fallback
# dram__bytes_read.sum + dram__bytes_write.sum per launch of the config-2/3 encoder call, from the committed
# `ncu --set full` capture of the fused-prologue kernels the model runs (profiles/r01o_msda_fused_ln_rowmask_ncu_full.txt:
# forward 116.2 + 29.4 MB, backward 238.5 + 98.9 MB; the plain-op kernels of profiles/r01b_msda_ncu_full.txt moved
# 142.6 / 340.9 MB)
TRAFFIC_NCU = {"msda_fwd_f32_d32": 145.6e6, "msda_bwd_f32_d32": 337.4e6}
# algorithmic bytes of the 2-image encoder call those captures profiled: forward 159.29 MB, backward 273.08 MB
TRAFFIC_ALGO = {"msda_fwd_f32_d32": 159.29e6, "msda_bwd_f32_d32": 273.08e6}


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            for k in ("hbm_gbs", "hbm_gb_s", "hbm"):
                if k in d:
                    return float(d[k]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def tensor_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        d = json.load(open(p))
        return float(d["bf16_tflops_sustained"]), "measured bf16 sustained (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, "fallback bf16 sustained (B200_PROFILING.md)"


def make_config(workload_text, images_per_step, world):
    """`config` of the JSON line -- the SAME dict for --impl ours and --impl reference."""
    return {"workload": workload_text, "images_per_step_per_gpu": images_per_step,
            "l2": "every call of a step reads its own buffers; the step cycles >3 GB (L2 is 126 MB)",
            "parallelism": f"dp{world}"}


def gpu_baseline_record():
    """The reference model itself on a B200 of this pool (tools/bench_reference_gpu.py; committed measurement)."""
    out = {}
    for tag in ("4scale", "5scale"):
        # newest committed measurement first (r02bq: timing-only run at the end of round 2, 4 scales)
        cands = [os.path.join(ROOT, "profiles", f"r02bq_reference_gpu_{tag}_timing.json"),
                 os.path.join(ROOT, "profiles", f"r02bb_reference_gpu_{tag}_init.json")]
        p = next((c for c in cands if os.path.exists(c)), cands[-1])
        if os.path.exists(p):
            try:
                d = json.load(open(p))
                out[tag] = {k: {"ms_per_step": v.get("ms_per_step"), "images_per_s": v.get("images_per_s")}
                            for k, v in d.get("reference_step", {}).items() if "ms_per_step" in v}
                out[tag]["ours_same_run"] = d.get("ours_step")
                out[tag]["file"] = os.path.relpath(p, ROOT)
            except Exception:
                pass
    if out:
        out["what"] = ("UNMODIFIED reference DINO (baseline/_ref) + its own MSDeformAttn CUDA extension rebuilt for sm_100a "
                       "(oracle/_ref), eager engine.py step on one B200 of this pool, identical synthetic batch; measured by "
                       "tools/bench_reference_gpu.py in an earlier gpurun call (the `file` of each entry), "
                       "NOT in this run; --gpu-baseline re-measures it live")
    return out or None


class LibraryModuleHooks:
    """Counts nn.Conv2d / nn.GroupNorm / nn.MultiheadAttention modules whose stock forward runs (i.e. cuDNN / ATen does
    the work): the hand-written paths read the module's weights and never call the module."""

    def __init__(self, model):
        import torch.nn as nn
        from datr_b200 import fallbacks
        self.handles = []
        for name, m in model.named_modules():
            if isinstance(m, nn.Conv2d):
                what = (f"nn.Conv2d {m.kernel_size[0]}x{m.kernel_size[1]} s{m.stride[0]} {m.in_channels}->{m.out_channels} "
                        f"(cuDNN fprop" + (" + dgrad/wgrad" if any(p.requires_grad for p in m.parameters()) else "") + ")")
            elif isinstance(m, nn.GroupNorm):
                what = f"nn.GroupNorm({m.num_groups}, {m.num_channels}) (ATen)"
            elif isinstance(m, nn.MultiheadAttention):
                what = "nn.MultiheadAttention (ATen / SDPA)"
            else:
                continue
            self.handles.append(m.register_forward_hook(lambda mod, i, o, w=what: fallbacks.note(w)))

    def remove(self):
        for h in self.handles:
            h.remove()


def msda_algo_bytes(N, S, M, D, L, Lq, P, es=4):
    """compulsory bytes of one call (gather re-reads, zero-fill and atomic RMW not credited)."""
    fwd = es * (N * S * M * D + 3 * N * Lq * M * L * P + N * Lq * M * D)
    bwd = fwd + es * (N * S * M * D + 3 * N * Lq * M * L * P)
    return fwd, bwd


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# workload: the MSDeformAttn calls of one DA training step
# ----------------------------------------------------------------------------------------------
def step_plan():
    """(kind, Lq) for every MSDeformAttn module invocation of one DA training step, forward order.
    Source pass: encoder x6 then decoder x6 with 900+200 denoising queries; target pass: 900."""
    S = sum(h * w for h, w in CFG2_LEVELS)
    plan = []
    for dec_q in (1100, 900):
        plan += [("enc", S)] * 6 + [("dec", dec_q)] * 6
    return S, plan


def synth_call(N, S, Lq, kind, seed, device, pinned=False):
    """Synthetic inputs of one call: value ~ N(0,1); encoder calls sample around each
    token's own reference point (offsets ~ N(0, (2 px)^2)); decoder calls around random box centres;
    attention = softmax(N(0,1)) over L*P."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    L = len(CFG2_LEVELS)
    value = torch.randn((N, S, M_HEADS, D_HEAD), generator=g)
    wh = torch.tensor([[w, h] for h, w in CFG2_LEVELS], dtype=torch.float32)
    if kind == "enc":
        refs = []
        for h, w in CFG2_LEVELS:
            ys, xs = torch.meshgrid((torch.arange(h) + 0.5) / h, (torch.arange(w) + 0.5) / w, indexing="ij")
            refs.append(torch.stack([xs.reshape(-1), ys.reshape(-1)], -1))
        ref = torch.cat(refs)[None, :, None, None, None, :]
        loc = ref + torch.randn((N, Lq, M_HEADS, L, N_POINTS, 2), generator=g) * 2.0 / wh[None, None, None, :, None, :]
    else:
        ctr = torch.rand((N, Lq, 1, 1, 1, 2), generator=g)
        box = torch.rand((N, Lq, 1, 1, 1, 2), generator=g) * 0.3 + 0.02
        loc = ctr + torch.randn((N, Lq, M_HEADS, L, N_POINTS, 2), generator=g) * 0.25 * box
    attn = torch.softmax(torch.randn((N, Lq, M_HEADS, L * N_POINTS), generator=g), -1).view(N, Lq, M_HEADS, L, N_POINTS)
    gout = torch.randn((N, Lq, M_HEADS * D_HEAD), generator=g)
    t = dict(value=value, loc=loc.contiguous(), attn=attn.contiguous(), grad_out=gout)
    if pinned:
        return {k: v.pin_memory() for k, v in t.items()}
    return {k: v.to(device) for k, v in t.items()}


class MsdaStep:
    name = "msda"

    def __init__(self, device, rank):
        from datr_b200 import MultiScaleDeformableAttention as MSDA
        from datr_b200 import native
        self.MSDA, self.native, self.device = MSDA, native, device
        self.S, self.plan = step_plan()
        self.N = 2
        L = len(CFG2_LEVELS)
        self.shapes = torch.tensor(CFG2_LEVELS, dtype=torch.int64, device=device)
        hw = self.shapes[:, 0] * self.shapes[:, 1]
        self.lstart = torch.cat([hw.new_zeros(1), hw.cumsum(0)[:-1]])
        # one private buffer set per call of the step => nothing is L2-resident from a previous call
        self.calls = [synth_call(self.N, self.S, Lq, kind, 1000 * rank + i, device) for i, (kind, Lq) in enumerate(self.plan)]
        self.bytes = [msda_algo_bytes(self.N, self.S, M_HEADS, D_HEAD, L, Lq, N_POINTS) for _, Lq in self.plan]
        # host copies for the end-to-end leg: one encoder set and one set per decoder length
        self.host = {}
        for i, (kind, Lq) in enumerate(self.plan):
            if (kind, Lq) not in self.host:
                self.host[(kind, Lq)] = synth_call(self.N, self.S, Lq, kind, 77 + i, device, pinned=True)
        self.host_out = {k: dict(out=torch.empty_like(v["grad_out"]).pin_memory(), gv=torch.empty_like(v["value"]).pin_memory(),
                                 gl=torch.empty_like(v["loc"]).pin_memory(), ga=torch.empty_like(v["attn"]).pin_memory())
                         for k, v in self.host.items()}
        self.launches_per_step = 2 * len(self.plan)
        self.workload = MSDA_WORKLOAD

    def step(self):
        F, B = self.MSDA.ms_deform_attn_forward, self.MSDA.ms_deform_attn_backward
        for c in self.calls:
            F(c["value"], self.shapes, self.lstart, c["loc"], c["attn"], 64)
        for c in reversed(self.calls):
            B(c["value"], self.shapes, self.lstart, c["loc"], c["attn"], c["grad_out"], 64)

    def e2e_step(self):
        """Same calls through the public op with HOST buffers: H2D of the inputs and D2H of the results
        of every call are inside the timed region."""
        F, B = self.MSDA.ms_deform_attn_forward, self.MSDA.ms_deform_attn_backward
        h2d = d2h = 0
        for kind, Lq in self.plan:
            h, o = self.host[(kind, Lq)], self.host_out[(kind, Lq)]
            v, l, a = (h[k].to(self.device, non_blocking=True) for k in ("value", "loc", "attn"))
            out = F(v, self.shapes, self.lstart, l, a, 64)
            o["out"].copy_(out, non_blocking=True)
            h2d += sum(h[k].numel() * 4 for k in ("value", "loc", "attn")); d2h += out.numel() * 4
        for kind, Lq in reversed(self.plan):
            h, o = self.host[(kind, Lq)], self.host_out[(kind, Lq)]
            v, l, a, g = (h[k].to(self.device, non_blocking=True) for k in ("value", "loc", "attn", "grad_out"))
            gv, gl, ga = B(v, self.shapes, self.lstart, l, a, g, 64)
            o["gv"].copy_(gv, non_blocking=True); o["gl"].copy_(gl, non_blocking=True); o["ga"].copy_(ga, non_blocking=True)
            h2d += sum(h[k].numel() * 4 for k in ("value", "loc", "attn", "grad_out"))
            d2h += (gv.numel() + gl.numel() + ga.numel()) * 4
        return h2d, d2h

    def roofline(self, timers, peak, peak_src):
        """Dominant kernel = the hand-written kernel group with the largest share of the event-timed launches;
        achieved = algorithmic bytes of its launches / their event-timed duration.  MSDeformAttn groups by
        direction and encoder/decoder; the tcgen05 linear kernel groups by (N, K) and also reports TFLOP/s."""
        groups = {}
        for kind, key, e0, e1 in timers:
            dt = e0.elapsed_time(e1) * 1e-3
            if kind in ("linear", "linear_bf16"):
                M, N, K, res = key
                bf = kind == "linear_bf16"
                name = f"linear_{'bf16' if bf else 'tf32'} N={N} K={K}" + ("+res" if res else "")
                g = groups.setdefault(name, [0.0, 0, 0, 0.0])
                # operand bytes: 4 (TF32 kernels read fp32) or 2 (bf16); the output / residual are counted as fp32 (upper bound
                # for the bf16-output launches)
                g[0] += dt; g[1] += (2 if bf else 4) * (M * K + N * K) + 4 * M * N * (2 if res else 1); g[2] += 1
                g[3] += 2.0 * M * N * K
                continue
            N, S, M, D, L, Lq, P, es = key
            name = f"msda_{kind}_f32_d32<{P}> " + ("encoder" if Lq == S else "decoder")
            fb, bb = msda_algo_bytes(N, S, M, D, L, Lq, P, es)
            g = groups.setdefault(name, [0.0, 0, 0, 0.0])
            g[0] += dt; g[1] += fb if kind == "fwd" else bb; g[2] += 1
        total = sum(g[0] for g in groups.values())
        top = max(groups, key=lambda k: groups[k][0])
        t, b, n, _ = groups[top]
        per_kernel = {}
        for k, v in groups.items():
            per_kernel[k] = {"launches": v[2], "avg_us": v[0] / v[2] * 1e6, "gbs": v[1] / v[0] / 1e9,
                             "frac": v[1] / v[0] / 1e9 / peak, "share": v[0] / total}
            if v[3]:
                per_kernel[k]["tflops"] = v[3] / v[0] / 1e12
        tpeak, tsrc = tensor_peak()
        lin = {k: v for k, v in groups.items() if v[3]}
        tensor = None
        if lin:
            fl, tt = sum(v[3] for v in lin.values()), sum(v[0] for v in lin.values())
            ktop = max(lin, key=lambda k: lin[k][0])
            tensor = {"what": "all tcgen05 linear launches of the inspected steps (TF32 and bf16 operands; the TF32 pipe peaks at "
                              "half the bf16 rate the denominator was measured with)",
                      "tflops": fl / tt / 1e12, "peak": tpeak, "peak_source": tsrc, "tensor_pipe_frac": fl / tt / 1e12 / tpeak,
                      "top_kernel": ktop, "top_kernel_tflops": lin[ktop][3] / lin[ktop][0] / 1e12,
                      "top_kernel_frac": lin[ktop][3] / lin[ktop][0] / 1e12 / tpeak, "share_of_timed_kernels": tt / total}
            for k in lin:
                per_kernel[k]["tensor_pipe_frac"] = per_kernel[k]["tflops"] / tpeak
        # ncu's DRAM traffic was captured on a 2-image call; launches of the joint encoder pass carry 4 images: scale to
        # the launch the algorithmic bytes describe (both are linear in the batch)
        traffic = TRAFFIC_NCU.get(top.split("<")[0])
        if traffic is not None and top.endswith("encoder"):
            traffic = traffic * (b / n) / (TRAFFIC_ALGO.get(top.split("<")[0], b / n))
        return {"bound": "hbm", "kernel": top, "achieved": b / t / 1e9, "peak": peak, "peak_source": peak_src, "tensor": tensor,
                "unit": "GB/s", "frac": b / t / 1e9 / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": b / n, "avg_launch_us": t / n * 1e6, "share_of_step": t / total,
                "handwritten_kernel_ms_per_step": None, "per_kernel": per_kernel}


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference's CPU MSDeformAttn path (grid_sample formulation), port in oracle/msda.py
# ----------------------------------------------------------------------------------------------
def cpu_port_images_per_s(threads, repeats=1):
    """Times one encoder, one Lq=1100 and one Lq=900 decoder call (forward + autograd backward) of the
    step on the host and extrapolates to the 12/6/6 calls of the step."""
    from oracle import msda as om
    torch.set_num_threads(threads)
    S, plan = step_plan()
    shapes = np.array(CFG2_LEVELS, dtype=np.int64)
    t_call = {}
    for kind, Lq in (("enc", S), ("dec", 1100), ("dec", 900)):
        c = synth_call(2, S, Lq, kind, 5, "cpu")
        best = None
        for _ in range(repeats + 1):          # first pass = warm-up
            v, l, a = (c[k].clone().requires_grad_(True) for k in ("value", "loc", "attn"))
            t0 = time.perf_counter()
            out = om.core_torch(v, shapes, l, a)
            out.backward(c["grad_out"].view_as(out))
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        t_call[(kind, Lq)] = best
    step_s = sum(t_call[k] for k in plan)
    return IMAGES_PER_STEP_PER_GPU / step_s, step_s, t_call


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("DATR_BENCH_WORKLOAD", "auto"))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-baseline", action="store_true",
                    help="also time the UNMODIFIED reference model on this GPU (tools/bench_reference_gpu.py) and put it in `gpu_baseline`")
    ap.add_argument("--no-eager", action="store_true", help="skip the eager (DATR_GRAPHS=0-equivalent) step time")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner with printf) and anything
    # else that writes to file descriptor 1 is sent to stderr for the whole run; the result goes to the saved descriptor
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    emit = lambda obj: (real_stdout.write(json.dumps(obj) + "\n"), real_stdout.flush())

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    threads = os.cpu_count() or 1

    if args.workload == "auto":
        try:
            import datr_b200.bench_dino  # noqa: F401
            args.workload = "dino"
        except ImportError:
            args.workload = "msda"

    if args.impl == "reference":
        if rank != 0:
            return
        if args.workload in ("dino", "dino5", "teacher"):
            from datr_b200 import bench_dino
            line = bench_dino.reference_arm(args, threads, args.workload)
            text = {"dino": bench_dino.WORKLOAD, "dino5": bench_dino.WORKLOAD_5SCALE, "teacher": bench_dino.WORKLOAD_TEACHER}[args.workload]
            line["config"] = make_config(text, 2 if args.workload == "dino5" else IMAGES_PER_STEP_PER_GPU, max(1, args.gpus))
            line["n_gpus"] = max(1, args.gpus)
            emit(line)
            return
        vals = []
        for _ in range(max(1, min(args.steps, 3))):
            ips, step_s, t_call = cpu_port_images_per_s(threads, repeats=1)
            vals.append((ips, step_s))
        ips, step_s = max(vals)
        sample = ("1 encoder + 1 decoder(1100) + 1 decoder(900) call, forward + autograd backward, of the "
                  "reference's grid_sample CPU path (func.py:41-61, port in oracle/msda.py), extrapolated x12/x6/x6")
        emit({
            "impl": "reference", "metric": "images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": make_config(MSDA_WORKLOAD, IMAGES_PER_STEP_PER_GPU, max(1, args.gpus)),
            "cpu_baseline": {"value": ips, "unit": "images/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line (NCCL prints its version)
        dist.init_process_group("nccl", device_id=device)
        # one process per GPU on one host: share the cores instead of every rank spawning a full intra-op pool
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    if args.workload in ("dino", "dino5", "teacher"):
        from datr_b200 import bench_dino
        cls = {"dino": bench_dino.DinoStep, "dino5": bench_dino.Dino5Step, "teacher": bench_dino.TeacherStep}[args.workload]
        wl = cls(device, rank, world)
    else:
        wl = MsdaStep(device, rank)
    images_per_gpu = getattr(wl, "n_images", IMAGES_PER_STEP_PER_GPU)
    from datr_b200 import MultiScaleDeformableAttention as MSDA
    from datr_b200 import native
    peak, peak_src = hbm_peak()

    for _ in range(args.warmup):
        wl.step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    from datr_b200 import linear as DL
    MSDA._timers = DL._timers = []
    n0 = native.all_launch_count()
    g0 = getattr(getattr(wl, "graphs", None), "replayed_native_launches", 0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        wl.step()
    e1.record()
    barrier()
    launches = native.all_launch_count() - n0
    timers, MSDA._timers, DL._timers = MSDA._timers, None, None
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    timing_note = "per-launch CUDA events on the launching stream inside the timed region"
    sg = getattr(wl, "graphs", None)
    eager = library = None
    if sg is not None:
        # graph replays launch our kernels without passing through the Python shims: count them from the capture
        # bookkeeping, and take the per-launch event timings from eager steps of the same workload right after.  The same
        # eager steps give the step time WITHOUT CUDA graphs (what a caller with varying shapes gets) and the list of
        # library kernels (cuBLAS / cuDNN / ATen modules) still on the path.
        from datr_b200 import fallbacks
        launches += sg.replayed_native_launches - g0
        wl.set_graphs(False)
        wl.step(); wl.step()
        k_eager = max(1, min(args.steps, 3))
        if not args.no_eager:
            barrier()
            g0e, g1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0e.record()
            for _ in range(k_eager):
                wl.step()
            g1e.record()
            barrier()
            eager = {"ms_per_step": g0e.elapsed_time(g1e) / k_eager, "steps": k_eager,
                     "what": "the same step with CUDA-graph replay switched off (every kernel launched from Python); graphs need "
                             "static shapes: per (image size, number of de-noising queries, number of boxes) signature, at "
                             "most 4 signatures per segment are captured, further ones run eagerly like this"}
        hooks = LibraryModuleHooks(wl.model) if hasattr(wl, "model") else None
        fallbacks.reset(); fallbacks.enabled = True
        MSDA._timers = DL._timers = []
        for _ in range(k_eager):
            wl.step()
        torch.cuda.synchronize()
        fallbacks.enabled = False
        if hooks is not None:
            hooks.remove()
        library = {k: v / k_eager for k, v in fallbacks.snapshot().items()}
        timers, MSDA._timers, DL._timers = MSDA._timers, None, None
        wl.set_graphs(True)
        timing_note = (f"per-launch CUDA events in {k_eager} eager steps of the same workload run right after the timed "
                       "region (graph replays cannot carry per-kernel events); value/e2e are measured with graph replay")

    # end-to-end leg: host buffers in, host results out, through the public op / model API
    for _ in range(2):
        wl.e2e_step()
    barrier()
    k2 = max(2, min(args.steps, 5))
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(k2):
        h2d, d2h = wl.e2e_step()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)

    if world > 1:
        t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = t.tolist()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    images = images_per_gpu * world
    value = images * args.steps / (ms * 1e-3)
    e2e = images * k2 / (ms_e2e * 1e-3)
    roof = wl.roofline(timers, peak, peak_src)
    n_timed_steps = args.steps if sg is None else k_eager
    roof["handwritten_kernel_ms_per_step"] = sum(e0.elapsed_time(e1) for _, _, e0, e1 in timers) / n_timed_steps
    roof["timing"] = timing_note
    line = {
        "metric": "images/sec", "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": getattr(wl, "dtype", "f32"), "data": "synthetic",
        "config": make_config(wl.workload, images_per_gpu, world),
        "roofline": roof,
        "e2e": {"value": e2e, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": k2, "ms_per_step": ms_e2e / k2},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if eager is not None:
        eager["images_per_s"] = images / (eager["ms_per_step"] * 1e-3)
        line["eager"] = eager
    if library is not None:
        line["library_calls_per_step"] = library
    gb = gpu_baseline_record()
    if args.gpu_baseline and world == 1:
        # live: the unmodified reference model on THIS GPU, right now (separate process: it monkey-patches nothing here)
        out = os.path.join(tempfile.gettempdir(), "datr_gpu_baseline.json")
        cfg = "5scale" if args.workload == "dino5" else "4scale"
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_reference_gpu.py"), "--config", cfg, "--steps", "5",
                            "--skip-parity", "--reference-only", "--out", out], capture_output=True, text=True)
        try:
            live = json.load(open(out))["reference_step"]
            gb = dict(gb or {}, live_this_run={k: {"ms_per_step": v.get("ms_per_step"), "images_per_s": v.get("images_per_s")}
                                               for k, v in live.items()})
        except Exception as e:  # noqa: BLE001
            gb = dict(gb or {}, live_this_run={"error": repr(e)[:200], "stderr": r.stderr[-300:]})
    if gb:
        line["gpu_baseline"] = gb
    if not args.no_cpu_baseline and world == 1:
        if args.workload in ("dino", "dino5", "teacher"):
            # the reference's own model on the host cores (bounded sample: 1 warm-up + 2 steps of 1 source + 1 target image)
            from datr_b200 import bench_dino
            ref = bench_dino.reference_arm(argparse.Namespace(steps=2, warmup=1, gpus=1), threads, args.workload)
            line["cpu_baseline"] = ref["cpu_baseline"]
        else:
            ips, step_s, t_call = cpu_port_images_per_s(threads, repeats=1)
            line["cpu_baseline"] = {
                "value": ips, "unit": "images/s", "cores": threads, "kind": "port",
                "sample": "1 encoder + 1 decoder(1100) + 1 decoder(900) call fwd+bwd of the grid_sample CPU path "
                          "(oracle/msda.py core_torch), extrapolated x12/x6/x6 to the step; %.2f s per step" % step_s}
    extra = getattr(wl, "extra", None)
    if extra:
        line.update(extra() if callable(extra) else extra)
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
