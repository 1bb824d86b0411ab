#!/usr/bin/env python
"""bench.py -- GoMatching's video hot path on B200: DeepSolo + LST-Matcher frames/s, MSDeformAttn HBM GB/s.  One JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload clip|op]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload "clip" (default; BASELINE.json metric "frames/sec (DeepSolo+LST, 1280x720)", configs[3]): the reference's OWN
GoMatching model -- ResNet-50 + 6+6-layer DeepSolo spotter + LSTMatcher, default-initialised and seeded, imported
unmodified from baseline/_ref (tools/refhost) -- with the B200 operator stack installed
(``gomatching_b200.install_into_adet``), driven by ``gomatching_b200.video.ClipTracker``: frames sharded over the
ranks, one record gather per round to the clip's tracker rank, the reference's sequential tracker (unchanged) INSIDE the
timed loop, overlapped with the next round's spotting.  At N GPUs the job is N concurrent clips (clip c tracked on rank
c, every clip's frames sharded over all N ranks): the tracker is sequential per clip, so clips are the axis along which
its work stays fixed per GPU; ``single_clip`` reports ONE clip over all N GPUs (bounded by one tracker).
One STEP = one round of every clip = ``--frames`` 1280x720 frames per rank per clip.
  value      frames/s, uint8 frames resident in HBM when the timed region starts; CUDA events; max over ranks
  e2e        same through the public API with HOST frames: pinned uint8 frame H2D per frame and the frame's track
             ids D2H inside the timed region
  spotting_only   one clip's loop without the association (the part that shards frame-wise)
Workload "op" (round 1's line; also measured in every clip run as ``msda``): the MSDeformAttn calls of F = 8 frames per
GPU, 6 encoder (Lq = S = 19160) + 6 decoder (Lq = 2500) launches per step on buffers larger than L2, fp32.
  roofline   dominant kernel of the path = the encoder-shape sampler launch; achieved = SURVEY.md s8(d) algorithmic
             bytes per launch / its mean CUDA-event duration inside that timed region; peak = MEASURED_PEAKS.json.
             ``roofline.in_pipeline`` = the same kernel timed live inside the clip's model forward (N = 1, value just
             written by value_proj, i.e. L2-warm).
  cpu_baseline / --impl reference   the reference's CPU implementation on the box's host cores, bounded sample:
             clip = the unmodified reference model end to end on one frame per step (MSDeformAttn via
             ms_deform_attn_core_pytorch, BASELINE.json configs[0]; kind "reference");
             op = 6 + 6 core_pytorch calls of one frame (kind "port", oracle.core_gridsample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

HEIGHT, WIDTH = 720, 1280
FRAMES_PER_STEP = 8
ENC_LAYERS, DEC_LAYERS = 6, 6
M, D, L, P = 8, 32, 4, 4


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "clip", "op"])
    ap.add_argument("--frames", type=int, default=0, help="frames per step per GPU (clip: 4, op: 8)")
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--config", default="GoMatching_ICDAR15",
                    help="clip: the reference config file (configs/<name>.yaml), e.g. GoMatching_PP_DSText (BASELINE.json configs[4])")
    ap.add_argument("--verbatim-tracker", action="store_true",
                    help="clip: run GoMatching.run_short_term_match / run_long_term_match verbatim instead of "
                         "video/association.py (identical IDs; the A/B for the tracker term)")
    ap.add_argument("--level", default="heads", choices=["op", "module", "layers", "transformer", "heads"],
                    help="install_into_adet level (clip)")
    ap.add_argument("--no-graph", action="store_true", help="clip: eager spotter instead of the CUDA-graph replay")
    ap.add_argument("--detections", type=int, default=40,
                    help="clip: calibrate the score threshold on the first frame so that this many of the 100 queries pass "
                         "(0: keep the config's 0.3, at which a default-initialised model passes all 100)")
    ap.add_argument("--tracker-frames", type=int, default=-1,
                    help="frames per step the tracker rank spots itself (clip, N > 1); -1 = same as the others")
    ap.add_argument("--dist", default="local", choices=["local", "uniform", "oor", "center"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-multi-clip", action="store_true",
                    help="clip, N > 1: time ONE clip over all GPUs as the headline instead of N concurrent clips")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sublines", action="store_true", help="skip the F=1 / 1080p / reference-kernel sub-lines")
    ap.add_argument("--unfused", action="store_true", help="op: time the core operator (loc/attn precomputed)")
    ap.add_argument("--tuning", default="", help="k=v,... msda_b200_tuning_t overrides for the encoder launch")
    return ap.parse_args()


def clip_available():
    try:
        from tools.refhost import loader
        return loader.reference_root() is not None
    except Exception:
        return False


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 100 ms while the timed regions run (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# configs (identical for both arms: built from the arguments only)
# ---------------------------------------------------------------------------------------------------
def clip_config(args):
    return {
        "workload": "clip%dp: synthetic %dx%d uint8 clip through the reference's GoMatching built from configs/%s.yaml "
                    "(ResNet-50 + DeepSolo 6+6 layers, point-query proposals x 25 points, d=256, 8 heads, 4 levels, 4 points + "
                    "the config's LST-Matcher head in the loop), default-initialised seeded weights, 1 frame per forward"
                    % (args.height, args.width, args.height, args.config),
        "frame": "%dx%d" % (args.width, args.height),
        "clips": "N concurrent clips at N GPUs (weak scaling in clips: one LST-Matcher rank per clip, each clip's frames sharded "
                 "over all N ranks); `single_clip` reports ONE clip over all N GPUs",
        "detections": ("score threshold calibrated on the first frame so that %d of the queries pass (SURVEY s8d: 20-60 at 100 "
                       "queries; default-initialised scores are flat)" % args.detections) if args.detections > 0 else "config threshold",
        "l2": "every frame's forward streams > L2 of activations (the encoder feed-forward intermediate alone is 78 MB "
              "per layer); the op-level lines rotate buffer sets larger than L2",
    }


def op_config(args):
    return {
        "workload": "MSDeformAttn hot path per 1280x720 DeepSolo-R50 frame: 6 encoder self-attn (Lq=S=19160, 4 levels "
                    "90x160..12x20) + 6 point-query decoder cross-attn (Lq=100x25) forwards, d=256, 8 heads, 4 points",
        "sampling_distribution": args.dist,
        "l2": "inputs larger than L2 (549 MB per encoder launch, 208 MB per decoder launch at 8 frames; buffer sets rotate)",
    }


# ---------------------------------------------------------------------------------------------------
# CPU reference paths -- baselines only
# ---------------------------------------------------------------------------------------------------
def cpu_op_frame_seconds(repeats: int, dist: str):
    """The reference's CPU path of the operator (ms_deform_attn_core_pytorch -> F.grid_sample, restated in
    oracle/msda_oracle.py) on ALL 6 + 6 calls of one 1280x720 frame; returns (median seconds per frame, threads)."""
    import torch
    from gomatching_b200 import synthetic as syn
    from oracle import msda_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    threads = torch.get_num_threads()
    ws = {kind: syn.make_workload(kind, HEIGHT, WIDTH, n=1, seed=0, dist=dist) for kind in ("encoder", "decoder")}

    def call(w):
        wh = torch.stack([w.shapes[:, 1], w.shapes[:, 0]], -1)
        attn = torch.softmax(w.logits, -1).view(w.attn.shape)                          # ms_deform_attn.py:139
        loc = w.ref[:, :, None, :, None, :] + w.offsets / wh[None, None, None, :, None, :]   # :143-144
        return O.core_gridsample(w.value, w.shapes.tolist(), loc, attn)                # :40-60

    def frame():
        for _ in range(ENC_LAYERS):
            call(ws["encoder"])
        for _ in range(DEC_LAYERS):
            call(ws["decoder"])

    ts = []
    for _ in range(repeats):
        t0 = time.perf_counter()
        frame()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), threads


class CpuClip:
    """The unmodified reference model on the host cores, through its own API only: one step = a fresh 2-frame mini clip
    through ``GoMatching.batch_inference`` (spotting with MSDeformAttn via ms_deform_attn_core_pytorch -- BASELINE.json
    configs[0] -- plus first-frame ID assignment and one short-term match)."""
    FRAMES = 2

    def __init__(self, args):
        import torch
        from tools.refhost import loader as L
        torch.set_num_threads(os.cpu_count() or 1)
        self.threads = torch.get_num_threads()
        self.L = L
        self.model = L.build_gomatching(L.build_cfg(config=args.config, device="cpu"), seed=0)
        self.inputs = L.frames_to_inputs(L.synthetic_clip(4, args.height, args.width, seed=1))
        if args.detections > 0:
            L.calibrate_detections(self.model, self.inputs[0], args.detections)
        self.k = 0

    def step(self):
        """returns seconds for the step's FRAMES frames"""
        import torch
        inp = [self.inputs[(self.k + i) % len(self.inputs)] for i in range(self.FRAMES)]
        self.k += 1
        t0 = time.perf_counter()
        with torch.no_grad():
            self.model.batch_inference(inp, 0, 0, [], self.L.new_time_cost())
        return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workload = resolve_workload(args)
    if workload == "clip":
        cpu = CpuClip(args)
        for _ in range(args.warmup):
            cpu.step()
        secs = [cpu.step() for _ in range(args.steps)]
        sec = sum(secs) / len(secs) / cpu.FRAMES
        threads, kind, fps_step = cpu.threads, "reference", cpu.FRAMES
        sample = ("per step: a fresh %d-frame %dx%d mini clip through the unmodified reference GoMatching.batch_inference on "
                  "the host (spotting with MSDeformAttn via ms_deform_attn_core_pytorch + ID assignment / short-term "
                  "match), %d threads" % (cpu.FRAMES, args.width, args.height, threads))
        cfg, metric = clip_config(args), "frames/sec (DeepSolo+LST, %dx%d)" % (args.width, args.height)
    else:
        secs = []
        for i in range(args.warmup + args.steps):
            s, threads = cpu_op_frame_seconds(1, args.dist)
            if i >= args.warmup:
                secs.append(s)
        sec = sum(secs) / len(secs)
        kind, fps_step = "port", 1
        sample = ("per step: all 6 encoder (Lq=S=19160) + 6 decoder (Lq=2500) MSDeformAttn calls of one 1280x720 frame via "
                  "F.grid_sample (oracle.core_gridsample = ms_deform_attn_core_pytorch restated), %d threads" % threads)
        cfg, metric = op_config(args), "frames/sec (MSDeformAttn hot path)"
    val = 1.0 / sec
    line = {
        "impl": "reference", "metric": metric, "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * fps_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "frames_per_step": fps_step,
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def resolve_workload(args):
    if args.workload == "auto":
        return "clip" if clip_available() else "op"
    if args.workload == "clip" and not clip_available():
        raise SystemExit("bench.py: --workload clip needs the reference's Python (run __graft_entry__.build() in the "
                         "build container to stage baseline/_ref)")
    return args.workload


# ---------------------------------------------------------------------------------------------------
# op workload
# ---------------------------------------------------------------------------------------------------
def device_workload(kind, frames, seed, dist, device, height=None, width=None, proposals=100):
    """One buffer set for a batch of `frames` frames, generated on the device with a seeded generator
    (reference points come from the CPU generator of gomatching_b200.synthetic)."""
    import torch
    from gomatching_b200 import synthetic as syn
    g = torch.Generator(device=device).manual_seed(seed)
    shapes_l = syn.level_shapes(height or HEIGHT, width or WIDTH, L)
    shapes = torch.as_tensor(shapes_l, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    if kind == "encoder":
        ref = syn.encoder_reference_points(shapes_l, 1).expand(frames, -1, -1, -1).contiguous()
    else:
        ref = syn.decoder_reference_points(torch.Generator().manual_seed(seed), frames, proposals, 25, L)
    Lq = ref.shape[1]
    ref = ref.to(device)
    wh = torch.stack([shapes[:, 1], shapes[:, 0]], -1).float().to(device)
    value = torch.randn(frames, S, M, D, generator=g, device=device)
    logits = torch.randn(frames, Lq, M, L * P, generator=g, device=device)
    if dist == "local":
        offsets = torch.randn(frames, Lq, M, L, P, 2, generator=g, device=device) * float(os.environ.get("MSDA_SIGMA_PX", "2.0"))
        offsets += syn.compass_offsets(M, L, P).to(device)[None, None]
    else:
        target = torch.rand(frames, Lq, M, L, P, 2, generator=g, device=device) * 1.2 - 0.1
        if dist == "oor":          # diagnostic: every sample out of range -> no gather at all
            target = target + 5.0
        elif dist == "center":     # diagnostic: every sample at the map centre -> one hot cache line per level
            target = target * 0.0 + 0.5
        offsets = (target - ref[:, :, None, :, None, :]) * wh[None, None, None, :, None, :]
    return {"kind": kind, "value": value, "ref": ref, "offsets": offsets.contiguous(), "logits": logits,
            "shapes": shapes.to(device), "lsi": syn.level_start_index(shapes_l).to(device), "Lq": Lq, "S": S,
            "shapes_list": shapes_l}


def algorithmic_bytes(frames, S, Lq):
    v = min(frames * S * M * D, 4 * frames * Lq * M * L * P * D) * 4
    return v + 12 * frames * Lq * M * L * P + 4 * frames * Lq * M * D


def time_launches(fn, sets, reps, warm=2):
    """Mean CUDA-event microseconds of fn(w) over `reps` passes of the rotating buffer sets (after `warm` passes)."""
    import torch
    for _ in range(warm):
        for w in sets:
            fn(w)
    evs = []
    for _ in range(reps):
        for w in sets:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(w)
            b.record()
            evs.append((a, b))
    torch.cuda.synchronize()
    return 1e3 * sum(a.elapsed_time(b) for a, b in evs) / len(evs)


def run_op_workload(args, world, rank, device, steps, warmup, barrier, with_e2e):
    import torch
    import torch.distributed as dist
    import gomatching_b200 as g
    from gomatching_b200 import _native

    F = args.frames or FRAMES_PER_STEP
    tuning = None
    if args.tuning:
        tuning = {k: int(v) for k, v in (kv.split("=") for kv in args.tuning.split(","))}
    enc = [device_workload("encoder", F, 100 + 7 * rank + i, args.dist, device) for i in range(ENC_LAYERS)]
    dec = [device_workload("decoder", F, 200 + 7 * rank + i, args.dist, device) for i in range(DEC_LAYERS)]
    if args.unfused:
        for w in enc + dec:
            w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)

    def launch(w, tn=None):
        if args.unfused:
            return g.ms_deform_attn_forward(w["value"], w["shapes"], w["lsi"], w["loc"], w["attn"], 64, tuning=tn)
        return g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"],
                                              tuning=tn, spatial_shapes_list=w["shapes_list"])

    rec_block = None
    if world > 1:
        from gomatching_b200 import video as V
        schema = V.RecordSchema(max_instances=100)
        rec_block = torch.randint(0, 255, (F, schema.stride), dtype=torch.uint8, device=device)

    enc_events = []

    def step(record=False):
        for w in enc:
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            launch(w, tuning)
            if record:
                b.record()
                enc_events.append((a, b))
        for w in dec:
            launch(w)
        if rec_block is not None:
            V.gather_records(rec_block, F * world, dst=0)

    for _ in range(max(warmup, 3)):
        step()
    barrier()
    calls0 = _native.calls
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step(record=True)
    t1.record()
    barrier()
    launches = _native.calls - calls0
    ms_total = t0.elapsed_time(t1)
    enc_ms = [a.elapsed_time(b) for a, b in enc_events]
    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / steps
    res = {"value": world * F / (ms_per_step * 1e-3), "ms_per_step": ms_per_step, "frames_per_step_per_gpu": F,
           "steps": steps, "launches": launches * world, "enc_mean_ms": sum(enc_ms) / len(enc_ms), "enc_launches": len(enc_ms),
           "b_alg": algorithmic_bytes(F, enc[0]["S"], enc[0]["Lq"]), "S": enc[0]["S"], "Lq": enc[0]["Lq"], "F": F,
           "fused_glue": not args.unfused, "e2e": None, "sublines": None}

    if with_e2e:
        keys = ("value", "ref", "offsets", "logits")
        host_in = [{k: torch.empty(w[k].shape, dtype=w[k].dtype, pin_memory=True).copy_(w[k]) for k in keys}
                   for w in enc + dec]
        host_out = [torch.empty((F, w["Lq"], M * D), dtype=torch.float32, pin_memory=True) for w in enc + dec]
        h2d = sum(t.numel() * t.element_size() for h in host_in for t in h.values())
        d2h = sum(t.numel() * t.element_size() for t in host_out)
        copy_stream = torch.cuda.Stream()

        def e2e_step():
            cur = torch.cuda.current_stream()
            staged = []
            for w, h in zip(enc + dec, host_in):
                with torch.cuda.stream(copy_stream):
                    d = {k: h[k].to(device, non_blocking=True) for k in keys}
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                staged.append((w, d, ev))
            for (w, d, ev), ho in zip(staged, host_out):
                cur.wait_event(ev)
                o = g.ms_deform_attn_forward_fused(d["value"], w["shapes"], w["lsi"], d["ref"], d["offsets"], d["logits"],
                                                   spatial_shapes_list=w["shapes_list"])
                for t in d.values():
                    t.record_stream(cur)
                ho.copy_(o, non_blocking=True)
            cur.synchronize()

        e2e_steps = max(3, min(steps, 10))
        for _ in range(2):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        e2e_s = (time.perf_counter() - w0) / e2e_steps
        if world > 1:
            t = torch.tensor([e2e_s], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        res["e2e"] = {"value": world * F / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                      "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                      "api": "gomatching_b200.ms_deform_attn_forward_fused on pinned host tensors (H2D on a copy stream, "
                             "D2H of every output): intermediate activations cross PCIe, which the clip workload avoids"}
        del host_in, host_out

    # ---- sub-lines (rank 0, N = 1 only): F = 1, bf16 storage, the 1080p shape (config 5), the reference kernel ---------
    if rank == 0 and world == 1 and not args.no_sublines:
        sub = {}
        peak = _peak()[0]

        def line(us, frames, S, Lq, esize=4):
            v = min(frames * S * M * D, 4 * frames * Lq * M * L * P * D) * esize
            b = v + 12 * frames * Lq * M * L * P + esize * frames * Lq * M * D
            return {"us_per_launch": us, "frames_per_launch": frames, "GBps": b / us / 1e3, "frac": b / us / 1e3 / peak}

        dec_us = time_launches(lambda w: launch(w), dec, 3)
        sub["decoder_f32_F%d" % F] = line(dec_us, F, dec[0]["S"], dec[0]["Lq"])
        del enc[2:], dec[2:]
        torch.cuda.empty_cache()
        e1 = [device_workload("encoder", 1, 300 + i, args.dist, device) for i in range(8)]
        sub["encoder_f32_F1"] = line(time_launches(lambda w: launch(w), e1, 5), 1, e1[0]["S"], e1[0]["Lq"])
        d1 = [device_workload("decoder", 1, 320 + i, args.dist, device) for i in range(8)]
        sub["decoder_f32_F1"] = line(time_launches(lambda w: launch(w), d1, 5), 1, d1[0]["S"], d1[0]["Lq"])
        del e1, d1
        for w in enc + dec:
            w["value"] = w["value"].bfloat16()
        sub["encoder_bf16_F%d" % F] = line(time_launches(lambda w: launch(w), enc, 5), F, enc[0]["S"], enc[0]["Lq"], 2)
        sub["decoder_bf16_F%d" % F] = line(time_launches(lambda w: launch(w), dec, 5), F, dec[0]["S"], dec[0]["Lq"], 2)
        # the bf16 operator MODE: neighbour-paired value layout (built outside the timed launch, like value itself)
        for w in enc + dec:
            w["paired"] = g.pair_value_bf16(w["value"], w["shapes"], w["lsi"])
            w["value"] = None

        def launch_paired(w):
            return g.ms_deform_attn_forward_fused_paired(w["paired"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"])
        sub["encoder_bf16_paired_F%d" % F] = line(time_launches(launch_paired, enc, 5), F, enc[0]["S"], enc[0]["Lq"], 2)
        sub["decoder_bf16_paired_F%d" % F] = line(time_launches(launch_paired, dec, 5), F, dec[0]["S"], dec[0]["Lq"], 2)
        pv = torch.randn(F, enc[0]["S"], M, D, device=device)
        t_pair = time_launches(lambda w: g.pair_value_bf16(pv, w["shapes"], w["lsi"]), enc[:1], 5)
        sub["pair_value_bf16_F%d" % F] = {"us_per_launch": t_pair, "GBps": (pv.numel() * 4 + pv.numel() * 4) / t_pair / 1e3,
                                          "note": "layout conversion fp32 (N,S,M,D) -> paired bf16 (N,M,S,2,D): read + write bytes"}
        del pv
        del enc[:], dec[:]
        torch.cuda.empty_cache()
        big = [device_workload("encoder", 4, 400 + i, args.dist, device, 1080, 1920) for i in range(3)]
        sub["encoder_f32_1080p_F4"] = line(time_launches(lambda w: launch(w), big, 3), 4, big[0]["S"], big[0]["Lq"])
        bigd = [device_workload("decoder", 4, 420 + i, args.dist, device, 1080, 1920, proposals=300) for i in range(3)]
        sub["decoder_f32_1080p_F4_300q"] = line(time_launches(lambda w: launch(w), bigd, 3), 4, bigd[0]["S"], bigd[0]["Lq"])
        del big, bigd
        torch.cuda.empty_cache()
        sub["gemm_3xtf32"] = gemm_sublines(device)
        res["sublines"] = sub
    return res


def gemm_sublines(device):
    """The projection GEMM (csrc/proj_gemm.cu) at the five encoder-layer shapes of ONE 1280x720 frame (M = 19 160 tokens) and
    at 8 frames: CUDA-event time of 20 back-to-back calls replayed from a CUDA graph (4 rotating input / output sets),
    against the tensor-pipe bound of its three TF32 passes -- peak TF32 = half the measured dense bf16 rate."""
    import torch
    from gomatching_b200.projections import linear_3xtf32
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = float(peaks.get("bf16_tflops", 2250.0)) / 2.0
    out = {"tensor_peak_tflops_tf32": tf32_peak,
           "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2" if "bf16_tflops" in peaks else "nominal 2250 / 2"}
    for M_rows in (19160, 153280):
        for name, N, K, relu in (("value_proj", 256, 256, False), ("offsets_attn", 384, 256, False), ("ffn1", 1024, 256, True),
                                 ("ffn2", 256, 1024, False)):
            x = [torch.randn(M_rows, K, device=device) for _ in range(4)]
            w = torch.randn(N, K, device=device) * 0.05
            b = torch.randn(N, device=device)
            y = [torch.empty(M_rows, N, device=device) for _ in range(4)]
            linear_3xtf32(x[0], w, b, out=y[0], relu=relu)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                with torch.cuda.graph(gr, stream=st):
                    for i in range(20):
                        linear_3xtf32(x[i % 4], w, b, out=y[i % 4], relu=relu)
            gr.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gr.replay()
            gr.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 40 * 1e3
            mma_tflops = 3 * 2.0 * M_rows * N * K / us / 1e6
            out["%s_M%d" % (name, M_rows)] = {"N": N, "K": K, "us_per_call": us, "tf32_mma_tflops": mma_tflops,
                                              "frac_of_tensor_peak": mma_tflops / tf32_peak,
                                              "GBps": (M_rows * K + M_rows * N + 2 * N * K) * 4 / us / 1e3}
            del x, y, gr
            torch.cuda.empty_cache()
    return out


def reference_loop_on_gpu(args, device, frames=16):
    """Context for the GPU number, baseline leg only: the reference's OWN serial loop (GoMatching.batch_inference, its
    eager modules, its host-side float conversion) on the same B200 with the UNMODIFIED reference CUDA kernel rebuilt for
    sm_100a behind adet._C -- what the reference delivers on this GPU without this library.  None if the kernel library
    was not built."""
    import torch
    from tools.refhost import loader as Lr
    model = Lr.build_gomatching(Lr.build_cfg(config=args.config, device=str(device)), seed=0)   # reference classes (un-patched)
    if not Lr.use_reference_cuda_kernel():
        Lr.restore_reference_classes()
        return None
    clip = Lr.synthetic_clip(8, args.height, args.width, seed=11)
    inputs = Lr.frames_to_inputs([clip[i % 8] for i in range(frames + 4)])
    if args.detections > 0:
        Lr.calibrate_detections(model, Lr.frames_to_inputs(Lr.synthetic_clip(1, args.height, args.width, seed=1))[0], args.detections)
    with torch.no_grad():
        model.batch_inference(inputs[:4], 0, 0, [], Lr.new_time_cost())               # warm-up clip
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        model.batch_inference(inputs[4:], 0, 0, [], Lr.new_time_cost())
        torch.cuda.synchronize(device)
        sec = (time.perf_counter() - t0) / frames
    Lr.restore_reference_classes()
    del model
    torch.cuda.empty_cache()
    return {"value": 1.0 / sec, "unit": "frames/s", "ms_per_frame": sec * 1e3, "frames": frames,
            "note": "the reference's serial loop (GoMatching.batch_inference, eager modules, host float conversion) on this "
                    "B200 with the unmodified reference CUDA kernel rebuilt for sm_100a; wall clock around one %d-frame clip" % frames}


def reference_cuda_kernel_us(device, dist_name, F):
    """The UNMODIFIED reference CUDA kernel rebuilt for sm_100a (oracle/_ref/libmsda_refcuda.so), same encoder-shape
    launch, same protocol -- the "kernel to beat".  Baseline leg only; None when the library was not built."""
    import ctypes
    import torch
    import gomatching_b200 as g
    path = os.path.join(ROOT, "oracle", "_ref", "libmsda_refcuda.so")
    if not os.path.exists(path):
        return None
    lib = ctypes.CDLL(path)
    lib.refcuda_msda_forward_f32.restype = ctypes.c_int
    lib.refcuda_msda_forward_f32.argtypes = [ctypes.c_void_p] * 5 + [ctypes.c_int] * 7 + [ctypes.c_void_p] * 2
    sets = [device_workload("encoder", F, 500 + i, dist_name, device) for i in range(3)]
    for w in sets:
        w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)
        w["out"] = torch.empty(F, w["Lq"], M * D, device=device)

    def fn(w):
        rc = lib.refcuda_msda_forward_f32(w["value"].data_ptr(), w["shapes"].data_ptr(), w["lsi"].data_ptr(),
                                          w["loc"].data_ptr(), w["attn"].data_ptr(), F, w["S"], M, D, L, w["Lq"], P,
                                          w["out"].data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    return time_launches(fn, sets, 3, warm=1)


# ---------------------------------------------------------------------------------------------------
# clip workload
# ---------------------------------------------------------------------------------------------------
def run_clip_workload(args, world, rank, device, steps, warmup, barrier):
    import torch
    import torch.distributed as dist
    from gomatching_b200 import _native
    from gomatching_b200.video.tracking import ClipTracker, round_plan
    from tools.refhost import loader as Lr

    F = args.frames or 4
    cfg = Lr.build_cfg(config=args.config, device=str(device))
    model = Lr.build_gomatching(cfg, seed=0, b200=args.level)
    pool_n = 8
    threshold = None
    if args.detections > 0:
        threshold = Lr.calibrate_detections(model, Lr.frames_to_inputs(Lr.synthetic_clip(1, args.height, args.width, seed=1))[0],
                                            args.detections)
    clip = Lr.synthetic_clip(pool_n, args.height, args.width, seed=11 + rank)
    host_pool = [torch.from_numpy(f).pin_memory() for f in clip]
    dev_pool = [f.to(device) for f in host_pool]

    def reduce_max(x):
        if world > 1:
            t = torch.tensor([x], device=device, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def run(pool, n_steps, clips=1, associate=True, host_results=False, graph=None, verbatim=None):
        """`clips` concurrent clips; clip c's tracker lives on rank c; every clip's round shards F frames to every rank
        (the tracker rank of a SINGLE clip spots --tracker-frames instead).  One step = one round of every clip.  Returns
        (ms per step by CUDA events incl. the association tail -- max over ranks, info, sampler event log)."""
        w0 = F if (args.tracker_frames < 0 or world == 1 or clips > 1) else args.tracker_frames
        weights = [w0] + [F] * (world - 1)
        per_round = sum(weights)
        mine = [(t, s) for t, r, s in round_plan(per_round, weights)[0] if r == rank]
        cts = [ClipTracker(model, weights=weights, tracker_rank=c, overlap=True, associate=associate,
                           host_results=host_results, graph=False if args.no_graph else graph,
                           fast_association=not (args.verbatim_tracker if verbatim is None else verbatim)) for c in range(clips)]
        my = cts[rank] if rank < clips else None           # the clip this rank tracks
        sg = cts[0].spotter_graph
        k = [0]

        def one_step():
            for c, ct in enumerate(cts):                       # same order on every rank: the gathers are collectives
                frames = [None] * per_round
                for t, s in mine:
                    frames[t] = pool[(k[0] * F + s + c) % pool_n]
                ct.feed(frames)
            k[0] += 1

        def flush():
            if my is not None:
                my.flush()

        for _ in range(max(warmup, 3)):
            one_step()
        flush()
        barrier()
        log = []
        _native.event_log = log
        calls0 = _native.calls + (sg.replayed_calls if sg else 0)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0 = my.association_seconds() if my else 0.0
        t0.record()
        for _ in range(n_steps):
            one_step()
        flush()                         # the last rounds' association is inside the timed region
        t1.record()
        barrier()
        _native.event_log = None
        ms = reduce_max(t0.elapsed_time(t1) / n_steps)
        frames_tracked = n_steps * per_round
        info = {"calls": _native.calls + (sg.replayed_calls if sg else 0) - calls0, "per_step": per_round * clips, "weights": weights,
                "assoc_ms_per_frame": ((my.association_seconds() - a0) * 1e3 / frames_tracked) if my else None,
                "d2h": (sum(t.numel() * t.element_size() for t in my.host_ids[-frames_tracked:]) / n_steps * clips)
                if (host_results and my is not None) else 0}
        for ct in cts:
            ct.drain()
        info["detections_per_frame"] = (sum(len(x) for x in my.instances[-frames_tracked:]) / frames_tracked
                                        if my is not None and associate and my.instances else None)
        info["graph"] = None if sg is None else {"replays": sg.replays, "failed": sg.failed}
        cts[0].close()
        return ms, info, log

    clips = world if (world > 1 and not args.no_multi_clip) else 1
    ms, info, log = run(dev_pool, steps, clips=clips)
    res = {"value": info["per_step"] / (ms * 1e-3), "ms_per_step": ms, "frames_per_step": info["per_step"], "weights": info["weights"],
           "clips": clips, "launches": info["calls"] * world, "assoc_ms_per_frame": info["assoc_ms_per_frame"], "level": args.level,
           "detections_per_frame": info["detections_per_frame"], "graph": info["graph"], "score_threshold": threshold}
    if not log:         # graph replay: the sampler launches are inside the graph; time them in a short eager pass of the same forward
        _, _, log = run(dev_pool, 3, associate=False, graph=False)
    enc_us = [a.elapsed_time(b) * 1e3 for tag, a, b in log if tag[3] == tag[2]]       # Lq == S: encoder self-attention
    dec_us = [a.elapsed_time(b) * 1e3 for tag, a, b in log if tag[3] != tag[2]]
    if enc_us:
        S = log[0][0][2]
        res["in_pipeline"] = {"encoder_us": statistics.mean(enc_us), "decoder_us": statistics.mean(dec_us) if dec_us else None,
                              "encoder_launches": len(enc_us), "S": S, "b_alg": algorithmic_bytes(1, S, S)}
    if not args.no_e2e:
        ms_e, info_e, _ = run(host_pool, steps, clips=clips, host_results=True)
        frame_bytes = host_pool[0].numel()
        res["e2e"] = {"value": info_e["per_step"] / (ms_e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": frame_bytes * info_e["per_step"],
                      "d2h_bytes_per_step": int(info_e["d2h"]), "ms_per_step": ms_e, "steps": steps,
                      "api": "gomatching_b200.video.ClipTracker.feed(pinned uint8 HWC frames) -> per-frame track ids on the "
                             "host (frame H2D, frame-batcher kernel, spotter, record gather, reference tracker, ids D2H)"}
    if world == 1 and not args.no_sublines and not args.verbatim_tracker:
        # the same loop with GoMatching.run_short_term_match / run_long_term_match called verbatim (identical track ids)
        ms_v, info_v, _ = run(dev_pool, max(3, steps // 2), clips=1, verbatim=True)
        res["verbatim_tracker"] = {"value": info_v["per_step"] / (ms_v * 1e-3), "unit": "frames/s", "ms_per_step": ms_v,
                                   "tracker_ms_per_frame": info_v["assoc_ms_per_frame"],
                                   "note": "ClipTracker(fast_association=False): the reference's matchers verbatim; same ids"}
    if not args.no_e2e and world == 1 and not args.no_sublines:
        # the same end-to-end loop fed with the frames as JPEG FILES (quality 90, 4:2:0): host Huffman stage in the
        # ClipTracker's decode-ahead threads, IDCT / upsampling / colour conversion on the device (bit-identical to Pillow)
        import io
        from PIL import Image
        jpeg_pool = []
        for f in clip:
            buf = io.BytesIO()
            Image.fromarray(f[:, :, ::-1].copy()).save(buf, "JPEG", quality=90)
            jpeg_pool.append(buf.getvalue())
        ms_j, info_j, _ = run(jpeg_pool, steps, clips=1, host_results=True)
        res["e2e_jpeg"] = {"value": info_j["per_step"] / (ms_j * 1e-3), "unit": "frames/s", "ms_per_step": ms_j,
                           "jpeg_bytes_per_frame": int(sum(len(b) for b in jpeg_pool) / len(jpeg_pool)),
                           "api": "ClipTracker.feed(JPEG files as bytes) -> per-frame track ids on the host"}
    if clips > 1:
        # ONE clip over all N GPUs (BASELINE.json configs[3] read literally): one tracker for the whole job -- the Amdahl term
        ms_1, info_1, _ = run(dev_pool, max(3, steps // 2), clips=1)
        res["single_clip"] = {"value": info_1["per_step"] / (ms_1 * 1e-3), "unit": "frames/s", "ms_per_step": ms_1,
                              "frames_per_step": info_1["per_step"], "round_weights": info_1["weights"],
                              "tracker_ms_per_frame": info_1["assoc_ms_per_frame"],
                              "note": "one clip sharded over all N GPUs, one tracker rank: bounded by the reference's sequential "
                                      "tracker (1000 / tracker_ms_per_frame frames/s) however many GPUs spot"}
    ms_s, info_s, _ = run(dev_pool, max(3, steps // 2), clips=1, associate=False)
    res["spotting_only"] = {"value": info_s["per_step"] / (ms_s * 1e-3), "unit": "frames/s", "ms_per_step": ms_s,
                            "note": "same loop, records gathered, association skipped: the part that shards"}
    return res


def _peak():
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    if "hbm_gbs" in peaks:
        return float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    return 6650.0, "6650 GB/s (of fallback, B200_PROFILING.md)"


_JSON_OUT = sys.stdout


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from gomatching_b200 import _native

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    _native.lib()
    workload = resolve_workload(args)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    clip = None
    if workload == "clip":
        clip = run_clip_workload(args, world, rank, device, args.steps, args.warmup, barrier)
        op_steps = max(3, min(args.steps, 5))
        op = run_op_workload(args, world, rank, device, op_steps, 3, barrier, with_e2e=False)
    else:
        op = run_op_workload(args, world, rank, device, args.steps, args.warmup, barrier, with_e2e=not args.no_e2e)
    clocks = sampler.stop() if rank == 0 else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = _peak()
    achieved = op["b_alg"] / (op["enc_mean_ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "target_frac": 0.60,
                "kernel": "encoder-shape sampler launch (N=%d, Lq=S=%d), fused glue" % (op["F"], op["S"]),
                "algorithmic_bytes_per_launch": op["b_alg"], "mean_launch_us": op["enc_mean_ms"] * 1e3,
                "launches_timed": op["enc_launches"], "peak_source": peak_src,
                "gather_bytes_per_launch": 4 * op["F"] * op["Lq"] * M * L * P * D * 4}
    if clocks and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        rows = 4 * op["F"] * op["Lq"] * M * L * P
        roofline["l1_pipe"] = {"bound": "l1 data pipe (1 row/clk/SM)", "unit": "rows/clk/SM", "peak": 1.0,
                               "achieved": rows / (op["enc_mean_ms"] * 1e-3 * clocks["sm_mhz"] * 1e6 * sms)}
    if clip and clip.get("in_pipeline"):
        ip = clip["in_pipeline"]
        ip["GBps"] = ip["b_alg"] / ip["encoder_us"] / 1e3
        ip["frac"] = ip["GBps"] / peak
        ip["note"] = "same kernel timed live inside the clip's model forward: N=1, value just written by value_proj (L2-warm)"
        roofline["in_pipeline"] = ip
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get("encoder_dram_bytes_per_launch_at_frames", {}).get(str(op["F"]))
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    cpu = None
    ref_kernel = None
    ref_gpu = None
    if not args.no_cpu_baseline:
        if workload == "clip":
            c = CpuClip(args)
            c.step()
            secs = [c.step() for _ in range(2)]
            sec = statistics.median(secs) / c.FRAMES
            cpu = {"value": 1.0 / sec, "unit": "frames/s", "cores": c.threads, "kind": "reference",
                   "sample": "2 steps after 1 warm-up, each a fresh %d-frame %dx%d mini clip through the unmodified reference "
                             "GoMatching.batch_inference on the host (MSDeformAttn via ms_deform_attn_core_pytorch)"
                             % (c.FRAMES, args.width, args.height)}
            del c
        else:
            sec, threads = cpu_op_frame_seconds(3, args.dist)
            cpu = {"value": 1.0 / sec, "unit": "frames/s", "cores": threads, "kind": "port",
                   "sample": "median of 3: all 6 encoder + 6 decoder calls of one frame via F.grid_sample "
                             "(oracle.core_gridsample = ms_deform_attn_core_pytorch restated)"}
        if world == 1 and workload == "clip":
            ref_gpu = reference_loop_on_gpu(args, device)
        if world == 1 and not args.no_sublines:
            us = reference_cuda_kernel_us(device, args.dist, op["F"])
            if us is not None:
                ref_kernel = {"ref_cuda_kernel_us": us, "b200_kernel_us": op["enc_mean_ms"] * 1e3,
                              "speedup": us / (op["enc_mean_ms"] * 1e3),
                              "note": "unmodified reference kernel rebuilt for sm_100a (core op; loc/attn precomputed), same "
                                      "encoder-shape launch and protocol; the B200 time includes the fused glue"}

    msda = {"metric": "frames/sec (MSDeformAttn hot path only)", "value": op["value"], "unit": "frames/s",
            "ms_per_step": op["ms_per_step"], "frames_per_step_per_gpu": op["frames_per_step_per_gpu"], "steps": op["steps"],
            "msda_hbm_gbs": achieved, "fused_glue": op["fused_glue"], "sublines": op["sublines"], "reference_kernel": ref_kernel,
            "config": op_config(args)}
    if workload == "clip":
        line = {
            "metric": "frames/sec (DeepSolo+LST, %dx%d)" % (args.width, args.height), "value": clip["value"], "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": clip["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded uint8 frames; reference initialisers, seeded)", "config": clip_config(args),
            "frames_per_step": clip["frames_per_step"], "round_weights": clip["weights"], "install_level": clip["level"],
            "clips": clip["clips"], "single_clip": clip.get("single_clip"),
            "parallelism": "dp%d: %d concurrent clip(s), clip c tracked on rank c; every clip's frames sharded over all ranks "
                           "(N=1 per forward), one NCCL gather of each round's records to the clip's tracker rank, the "
                           "reference's tracker in a worker thread there" % (world, clip["clips"]),
            "tracker": ("GoMatching.run_short_term_match / run_long_term_match verbatim" if args.verbatim_tracker else
                        "video/association.py: the reference's association modules and Hungarian step, ID bookkeeping on "
                        "the host (track ids bit-identical to the verbatim matchers)"),
            "tracker_ms_per_frame": clip["assoc_ms_per_frame"], "detections_per_frame": clip["detections_per_frame"],
            "score_threshold": clip["score_threshold"], "cuda_graph": clip["graph"], "spotting_only": clip["spotting_only"],
            "clocks": clocks, "e2e": clip.get("e2e"), "e2e_jpeg": clip.get("e2e_jpeg"), "verbatim_tracker": clip.get("verbatim_tracker"), "gpu_launches": clip["launches"],
            "gpu_launches_note": "kernel-launching C-ABI calls of libmsda_b200.so in the timed region, all ranks (each "
                                 "enqueues >= 1 kernel); cuDNN / cuBLAS kernels of the reference's eager code not counted",
            "roofline": roofline, "cpu_baseline": cpu, "reference_on_gpu": ref_gpu, "msda": msda, "msda_hbm_gbs": achieved,
        }
    else:
        line = {
            "metric": "frames/sec (MSDeformAttn hot path)", "value": op["value"], "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": op["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded, generated on device)",
            "config": op_config(args), "frames_per_step": op["frames_per_step_per_gpu"] * world, "clocks": clocks,
            "e2e": op["e2e"], "gpu_launches": op["launches"], "roofline": roofline, "cpu_baseline": cpu, "msda": msda,
            "msda_hbm_gbs": achieved,
        }
    print(json.dumps(line), file=_JSON_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _only_json_on_stdout():
    """Libraries (NCCL prints its version banner on the first communicator) must not add lines to stdout: the driver
    reads ONE JSON line there.  File descriptor 1 is pointed at stderr for the whole run; the JSON lines go to the
    saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, "w")


if __name__ == "__main__":
    _JSON_OUT = _only_json_on_stdout()
    main()
