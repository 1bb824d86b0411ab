#!/usr/bin/env python
"""Launch-parameter sweep for the tiled MSDeformAttn kernels (run on the GPU box).

    python tools/sweep.py [--frames 1] [--kind encoder] [--dtype f32] [--out gpurun_out/sweep.json]

Times every (variant, mode, tile, ctas_per_sm) combination L2-cold (rotating buffer sets whose total
exceeds 4x L2), prints the best ones, and -- when oracle/_ref/libmsda_refcuda.so exists -- the unmodified
reference kernel under the same protocol (the "kernel to beat").  Results never differ between
configurations (tests/test_msda_gpu.py); only time does.
"""
import argparse
import ctypes
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gomatching_b200 as g  # noqa: E402
from gomatching_b200 import _native  # noqa: E402
import bench  # noqa: E402


def time_launches(fn, sets, iters):
    for i in range(3):
        fn(sets[i % len(sets)])
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--kind", default="encoder")
    ap.add_argument("--dtype", default="f32")
    ap.add_argument("--dist", default="local")
    ap.add_argument("--iters", type=int, default=24)
    ap.add_argument("--fused", type=int, default=0)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--force-v1", type=int, default=0)
    ap.add_argument("--walk", type=int, default=0)
    ap.add_argument("--variants", default="", help="comma list; default all")
    ap.add_argument("--cps", default="", help="comma list of CTAs per SM; default 2,3,4 (quick) or 1..8")
    ap.add_argument("--top", type=int, default=12)
    ap.add_argument("--staged-diag", action="store_true",
                    help="also time the staged kernel's diagnostic builds (variants 1-4: WRONG results, time attribution only)")
    ap.add_argument("--height", type=int, default=720)
    ap.add_argument("--width", type=int, default=1280)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    F = args.frames
    bench.HEIGHT, bench.WIDTH = args.height, args.width
    from gomatching_b200 import synthetic as syn
    S_all = sum(h * w for h, w in syn.level_shapes(args.height, args.width, 4))
    one = bench.algorithmic_bytes(F, S_all, S_all if args.kind == "encoder" else 2500)
    nsets = max(2, int(4 * 126e6 / one) + 1)
    sets = [bench.device_workload(args.kind, F, 300 + i, args.dist, dev) for i in range(nsets)]
    for w in sets:
        w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)
        if args.dtype == "bf16":
            w["value"] = w["value"].to(torch.bfloat16)
    b_alg = one if args.dtype == "f32" else one - F * S_all * 256 * 2 - F * sets[0]["Lq"] * 256 * 2

    def runner(tn):
        if args.fused:
            return lambda w: g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"],
                                                            w["logits"], tuning=tn)
        return lambda w: g.ms_deform_attn_forward(w["value"], w["shapes"], w["lsi"], w["loc"], w["attn"], 64, tuning=tn)

    results = []
    variants = range(_native.lib().msda_b200_variant_count())
    if args.variants:
        variants = [int(v) for v in args.variants.split(",")]
    cps_list = [1, 2, 3, 4, 6, 8] if not args.quick else [2, 3, 4]
    if args.cps:
        cps_list = [int(v) for v in args.cps.split(",")]
    tiles = []
    if args.kind == "encoder":
        tiles += [dict(mode=2, tile_h=h, tile_w=w) for h, w in
                  ([(8, 16), (16, 8), (16, 16), (8, 32), (16, 32)] if args.quick else
                   [(4, 8), (4, 16), (8, 8), (8, 16), (16, 8), (16, 16), (8, 32), (16, 32), (4, 32), (2, 32), (32, 32)])]
    tiles += [dict(mode=1, tile_q=q) for q in ((64, 128, 256) if args.quick else (16, 32, 64, 128, 256))]
    staged = []
    if args.kind == "encoder" and args.dtype == "f32":   # shared-memory-window kernel: tile_h = query levels staged
        staged = [dict(mode=4, tile_h=n, variant=v, ctas_per_sm=1, force_v1=0, walk=args.walk) for n in (1, 2, 3) for v in ((0, 1, 2, 3, 4) if n == 1 and not args.fused and args.staged_diag else (0,))]
        staged.append(dict(mode=5, variant=0, ctas_per_sm=1, force_v1=0, walk=0))   # producer / consumer window kernel   # variant 1: diagnostic, no fallback path (wrong results)
    combos = [dict(t, variant=v, ctas_per_sm=cps, force_v1=args.force_v1, walk=args.walk)
              for v, t, cps in itertools.product(variants, tiles, cps_list)] + staged
    for tn in combos:
        try:
            us = time_launches(runner(tn), sets, args.iters)
        except Exception as e:  # noqa: BLE001
            print("skip", tn, e)
            continue
        results.append({"tuning": tn, "us": us, "gbs": b_alg / us / 1e3})
    results.sort(key=lambda r: r["us"])
    base = time_launches(runner(None), sets, args.iters)
    print("== %s %dx%d frames=%d dtype=%s fused=%d dist=%s  B_alg=%.2f MB  sets=%d" % (args.kind, args.height, args.width, F, args.dtype, args.fused,
                                                                               args.dist, b_alg / 1e6, nsets))
    print("default heuristics: %.2f us  %.0f GB/s" % (base, b_alg / base / 1e3))
    for r in results[:args.top]:
        print("%8.2f us %7.0f GB/s  %s" % (r["us"], r["gbs"], r["tuning"]))
    print("worst: %.2f us %s" % (results[-1]["us"], results[-1]["tuning"]))
    for v in variants:
        best = next((r for r in results if r["tuning"]["variant"] == v), None)
        if best:
            print("best of variant %d: %8.2f us %7.0f GB/s  %s" % (v, best["us"], best["gbs"], best["tuning"]))
    generic = time_launches(runner(dict(mode=3)), sets, 6) if not args.fused else None
    if generic:
        print("generic kernel: %.2f us" % generic)
    ref_us = None
    refp = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")
    if os.path.exists(refp) and args.dtype == "f32":
        lib = ctypes.CDLL(refp)
        for name in ("refcuda_msda_forward_f32", "refcuda_msda_forward_f32_nomemset"):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 2
        outs = [torch.empty(F, w["Lq"], 256, device=dev) for w in sets[:2]]

        def ref_runner(name):
            fn = getattr(lib, name)

            def run(w):
                o = outs[0]
                rc = fn(w["value"].data_ptr(), w["shapes"].data_ptr(), w["lsi"].data_ptr(), w["loc"].data_ptr(),
                        w["attn"].data_ptr(), F, w["S"], 8, 32, 4, w["Lq"], 4, o.data_ptr(),
                        torch.cuda.current_stream().cuda_stream)
                assert rc == 0
            return run
        ref_us = {n: time_launches(ref_runner(n), sets, args.iters) for n in
                  ("refcuda_msda_forward_f32", "refcuda_msda_forward_f32_nomemset")}
        print("reference CUDA kernel (sm_100a rebuild): with memset %.2f us, kernel only %.2f us  (%.0f GB/s)" % (
            ref_us["refcuda_msda_forward_f32"], ref_us["refcuda_msda_forward_f32_nomemset"],
            b_alg / ref_us["refcuda_msda_forward_f32_nomemset"] / 1e3))
    if args.out:
        os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
        json.dump({"args": vars(args), "b_alg": b_alg, "default_us": base, "generic_us": generic, "reference_us": ref_us,
                   "results": results}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
