"""Micro-benchmark of the MSDeformAttn kernels at the BASELINE.json shapes (GPU box only).
Times ours and, when oracle/_ref is built, the reference's own CUDA kernels, with an L2 flush
between iterations.  Prints one line per (shape, direction)."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import msda_cases as mc
from datr_b200 import MultiScaleDeformableAttention as MSDA
from oracle import build_ref

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def algo_bytes(N, S, M, D, L, Lq, P, es=4):
    fwd = es * (N * S * M * D + 3 * N * Lq * M * L * P + N * Lq * M * D)
    return fwd, fwd + es * (N * S * M * D + 3 * N * Lq * M * L * P)


def timeit(fn, iters=20, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(np.min(ts))


def main():
    """Location distributions: `encoder` = pixel centres + N(0, (2 px)^2) per sample (no coherence between neighbouring
    queries); `init` = the offset pattern of MSDeformAttn._reset_parameters (what a
    freshly built model, and bench.py's dino workload, samples); `init+0.3px` = that pattern with per-sample jitter
    (smooth learned offsets); `uniform` = locations anywhere in the map."""
    coherent_inputs = mc.coherent_inputs
    ref = None if "--no-ref" in sys.argv else build_ref.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    cases = [("cfg2_enc", 2, mc.CFG2_LEVELS, -1, "encoder"), ("cfg2_enc_init", 2, mc.CFG2_LEVELS, -1, ("coherent", 0.0)),
             ("cfg2_enc_init+0.3px", 2, mc.CFG2_LEVELS, -1, ("coherent", 0.3)),
             ("cfg2_enc_uniform", 2, mc.CFG2_LEVELS, -1, "uniform"),
             ("cfg2_dec1100", 2, mc.CFG2_LEVELS, 1100, "uniform"), ("cfg1_enc", 1, mc.CFG1_LEVELS, -1, "encoder"),
             ("cfg4_enc", 1, mc.CFG4_LEVELS, -1, "encoder"), ("cfg4_enc_init+0.3px", 1, mc.CFG4_LEVELS, -1, ("coherent", 0.3))]
    only = [a.split("=", 1)[1].split(",") for a in sys.argv if a.startswith("--cases=")]
    fused = "--fused" in sys.argv
    # --bwd-stages=0,1,2: time the backward once per scatter variant (0 = vector reductions, 1 / 2 = TMA reduce) and
    # compare every variant's gradients with variant 0
    stages = [int(v) for a in sys.argv if a.startswith("--bwd-stages=") for v in a.split("=", 1)[1].split(",")]
    from datr_b200 import native
    for name, N, levels, Lq, mode in cases:
        if only and name not in only[0]:
            continue
        if isinstance(mode, tuple):
            inp = coherent_inputs(N, 8, 4, levels, mode[1], 1)
        else:
            inp = mc.make_inputs(N, 8, 32, Lq, 4, levels, mode, 1, np.float32)
        d = {k: torch.from_numpy(v).cuda() for k, v in inp.items()}
        args = (d["value"], d["shapes"], d["level_start"], d["loc"], d["attn"])
        S = d["value"].shape[1]; LQ = d["loc"].shape[1]
        fb, bb = algo_bytes(N, S, 8, 32, len(levels), LQ, 4)
        if fused and LQ == S:
            # the module-level entry points the model calls: raw offsets + logits + 2-d reference points
            refp = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(
                mc.encoder_reference_points(levels)[None, :, None, :], (N, S, len(levels), 2)), dtype=np.float32)).cuda()
            wh = d["shapes"].flip(-1).float()
            off = ((d["loc"] - refp[:, :, None, :, None, :]) * wh[None, None, None, :, None, :]).contiguous()
            lg = d["attn"].flatten(3).log().contiguous()
            fa = (d["value"], d["shapes"], d["level_start"], off, lg, refp)
            tf, _ = timeit(lambda: MSDA.ms_deform_attn_fused_forward(*fa), flush=flush)
            tb, _ = timeit(lambda: MSDA.ms_deform_attn_fused_backward(*fa, d["grad_out"]), flush=flush)
            print(f"{name:18s} fused cold fwd {tf*1e3:8.1f} us ({fb/tf/1e6:7.1f} GB/s, {fb/tf/1e6/PEAK:5.3f}) "
                  f"bwd {tb*1e3:8.1f} us ({bb/tb/1e6:7.1f} GB/s, {bb/tb/1e6/PEAK:5.3f})", flush=True)
            if "--pairs" in sys.argv:
                want = MSDA.ms_deform_attn_fused_forward(*fa)
                for dt in (torch.bfloat16, torch.float16):
                    tp, _ = timeit(lambda: MSDA.pack_value_pairs(d["value"], d["shapes"], d["level_start"], dt), flush=flush)
                    pairs = MSDA.pack_value_pairs(d["value"], d["shapes"], d["level_start"], dt)
                    tf2, _ = timeit(lambda: MSDA.ms_deform_attn_fused_forward(*fa, pairs=pairs), flush=flush)
                    got = MSDA.ms_deform_attn_fused_forward(*fa, pairs=pairs)
                    rounded = MSDA.ms_deform_attn_fused_forward(d["value"].to(dt).float(), *fa[1:])
                    fb2 = fb - 2 * N * S * 8 * 32       # the value map is read as 2 bytes per element
                    print(f"{name:18s} fused fwd on {str(dt)[6:]} pair rows {tf2*1e3:8.1f} us ({fb2/tf2/1e6:7.1f} GB/s, {fb2/tf2/1e6/PEAK:5.3f}) "
                          f"+ pack {tp*1e3:6.1f} us; max rel diff vs fp32 rows {float((got - want).abs().max() / want.abs().max()):.2e}, "
                          f"vs fp32 rows of rounded values {float((got - rounded).abs().max() / want.abs().max()):.2e}", flush=True)
        base = None
        for st in stages:
            native.lib().datr_msda_set_backward_stages(st)
            tb, tbm = timeit(lambda: MSDA.ms_deform_attn_backward(*args, d["grad_out"], 64), flush=flush)
            got = MSDA.ms_deform_attn_backward(*args, d["grad_out"], 64)
            base = got if base is None else base
            err = [float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(got, base)]
            print(f"{name:18s} bwd scatter variant {st} (active {native.lib().datr_msda_get_backward_stages()}): {tb*1e3:8.1f} us "
                  f"(min {tbm*1e3:8.1f}; {bb/tb/1e6:7.1f} GB/s, {bb/tb/1e6/PEAK:5.3f})  max rel diff vs variant {stages[0]}: "
                  f"value {err[0]:.2e} loc {err[1]:.2e} attn {err[2]:.2e}", flush=True)
        if stages:
            native.lib().datr_msda_set_backward_stages(-1)
            continue
        for impl, mod in (("ours", MSDA), ("ref", ref)):
            if mod is None:
                continue
            for flushed, fl in (("cold", flush),):
                tf, tfm = timeit(lambda: mod.ms_deform_attn_forward(*args, 64), flush=fl)
                tb, tbm = timeit(lambda: mod.ms_deform_attn_backward(*args, d["grad_out"], 64), flush=fl)
                print(f"{name:18s} {impl:5s} {flushed} fwd {tf*1e3:8.1f} us ({fb/tf/1e6:7.1f} GB/s, {fb/tf/1e6/PEAK:5.3f}) "
                      f"bwd {tb*1e3:8.1f} us ({bb/tb/1e6:7.1f} GB/s, {bb/tb/1e6/PEAK:5.3f})", flush=True)


if __name__ == "__main__":
    main()
