#!/usr/bin/env python
"""bench.py -- the MSDeformAttn hot path of GoMatching/DeepSolo on B200, one JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], with configs[2]'s decoder shape riding along): one STEP is the
MSDeformAttn hot path of a batch of F = 8 synthetic 1280x720 DeepSolo-R50 frames per GPU -- 6 encoder
self-attention calls (Lq = S = 19160 tokens) + 6 point-query decoder cross-attention calls
(Lq = 100 x 25 = 2500), d_model 256, 8 heads, 4 levels, 4 points, fp32 -- each call doing the four things
north_star names: offset->location math, softmax over levels x points, multi-level bilinear gather,
weighted reduction.  metric = frames/s (frames whose hot path completed per second, whole job).

  value      inputs resident in HBM; CUDA events; max over ranks
  e2e        same work through the public API with HOST (pinned) buffers: H2D of every input and D2H of
             every output inside the timed region
  roofline   dominant kernel = the encoder-shape launch; achieved = SURVEY.md s8(d) algorithmic bytes per
             launch / its mean CUDA-event duration inside the timed region; peak = MEASURED_PEAKS.json
  cpu_baseline  the reference's CPU path (ms_deform_attn_core_pytorch -> F.grid_sample, restated in
             oracle/msda_oracle.py because the reference's Python cannot travel to the GPU box) on the
             box's host cores, bounded sample
  --impl reference   that CPU path as its own arm (rank 0 only under torchrun)

L2 hygiene: every launch works on buffers larger than L2 (one encoder call at F=8 touches 549 MB, one
decoder call 208 MB, L2 is 126 MB) and consecutive launches use different buffer sets.
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
    ap.add_argument("--frames", type=int, default=FRAMES_PER_STEP, help="frames per step per GPU")
    ap.add_argument("--dist", default="local", choices=["local", "uniform", "oor", "center"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="time the core operator (loc/attn precomputed)")
    ap.add_argument("--tuning", default="", help="k=v,... msda_b200_tuning_t overrides for the encoder launch")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md recipe)."""
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
                                          "--format=csv,noheader,nounits", "-lms", "200"],
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
        time.sleep(0.25)
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
# CPU reference path (oracle port of ms_deform_attn_core_pytorch + the eager glue) -- checker / baseline only
# ---------------------------------------------------------------------------------------------------
def cpu_reference_frame_seconds(repeats: int, dist: str):
    """Times 1 encoder + 1 decoder MSDeformAttn call of ONE frame on the host; returns
    (seconds per frame = 6*t_enc + 6*t_dec, t_enc, t_dec, threads)."""
    import torch
    from gomatching_b200 import synthetic as syn
    from oracle import msda_oracle as O

    torch.set_num_threads(os.cpu_count() or 1)
    threads = torch.get_num_threads()
    out = {}
    for kind in ("encoder", "decoder"):
        w = syn.make_workload(kind, HEIGHT, WIDTH, n=1, seed=0, dist=dist)
        wh = torch.stack([w.shapes[:, 1], w.shapes[:, 0]], -1)

        def call():
            attn = torch.softmax(w.logits, -1).view(w.attn.shape)                          # ms_deform_attn.py:139
            loc = w.ref[:, :, None, :, None, :] + w.offsets / wh[None, None, None, :, None, :]   # :143-144
            return O.core_gridsample(w.value, w.shapes.tolist(), loc, attn)                # :40-60
        call()
        ts = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            call()
            ts.append(time.perf_counter() - t0)
        out[kind] = statistics.median(ts)
    return ENC_LAYERS * out["encoder"] + DEC_LAYERS * out["decoder"], out["encoder"], out["decoder"], threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each "step" = one frame's 1 enc + 1 dec call on the host, scaled to the frame's 6 + 6 calls
    import torch  # noqa: F401
    per_frame = []
    threads = 1
    for i in range(args.warmup + args.steps):
        s, te, td, threads = cpu_reference_frame_seconds(1, args.dist)
        if i >= args.warmup:
            per_frame.append(s)
    sec = sum(per_frame) / len(per_frame)
    val = 1.0 / sec
    sample = ("per step: 1 encoder (Lq=S=19160) + 1 decoder (Lq=2500) MSDeformAttn call of one 1280x720 frame via "
              "F.grid_sample, scaled x6 each")
    line = {
        "impl": "reference", "metric": "frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": val, "unit": "frames/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def workload_config(args, frames):
    return {
        "workload": "MSDeformAttn hot path per 1280x720 DeepSolo-R50 frame: 6 encoder self-attn (Lq=S=19160, 4 levels "
                    "90x160..12x20) + 6 point-query decoder cross-attn (Lq=100x25) forwards, d=256, 8 heads, 4 points",
        "frames_per_step_per_gpu": frames, "sampling_distribution": args.dist,
        "fused_glue": not args.unfused,
        "l2": "inputs larger than L2 (549 MB per encoder launch, 208 MB per decoder launch; buffer sets rotate)",
        "parallelism": "frames sharded across GPUs (dp%d); no collective inside the op, one NCCL gather of the "
                       "per-frame records to the tracker rank per step when N > 1" % args.gpus,
    }


# ---------------------------------------------------------------------------------------------------
# device workload
# ---------------------------------------------------------------------------------------------------
def device_workload(kind, frames, seed, dist, device):
    """One buffer set for a batch of `frames` frames, generated on the device with a seeded generator
    (reference points come from the CPU generator of gomatching_b200.synthetic)."""
    import torch
    from gomatching_b200 import synthetic as syn
    g = torch.Generator(device=device).manual_seed(seed)
    shapes_l = syn.level_shapes(HEIGHT, WIDTH, L)
    shapes = torch.as_tensor(shapes_l, dtype=torch.long)
    S = int(shapes.prod(1).sum())
    if kind == "encoder":
        ref = syn.encoder_reference_points(shapes_l, 1).expand(frames, -1, -1, -1).contiguous()
    else:
        ref = syn.decoder_reference_points(torch.Generator().manual_seed(seed), frames, 100, 25, L)
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
            "shapes": shapes.to(device), "lsi": syn.level_start_index(shapes_l).to(device), "Lq": Lq, "S": S}


def algorithmic_bytes(frames, S, Lq):
    v = min(frames * S * M * D, 4 * frames * Lq * M * L * P * D) * 4
    return v + 12 * frames * Lq * M * L * P + 4 * frames * Lq * M * D


_JSON_OUT = sys.stdout


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    import gomatching_b200 as g
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

    F = args.frames
    tuning = None
    if args.tuning:
        tuning = {k: int(v) for k, v in (kv.split("=") for kv in args.tuning.split(","))}

    # one buffer set per layer so consecutive launches never touch the same bytes
    enc = [device_workload("encoder", F, 100 + 7 * rank + i, args.dist, device) for i in range(ENC_LAYERS)]
    dec = [device_workload("decoder", F, 200 + 7 * rank + i, args.dist, device) for i in range(DEC_LAYERS)]
    if args.unfused:
        for w in enc + dec:
            w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)

    def launch(w, tn=None):
        if args.unfused:
            return g.ms_deform_attn_forward(w["value"], w["shapes"], w["lsi"], w["loc"], w["attn"], 64, tuning=tn)
        return g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"],
                                              tuning=tn)

    # N > 1: the path's one exchange step -- per-step gather of this rank's frame records (query embeddings and
    # rescored detections, 0.49 MB per frame at 100 queries) to the tracker rank over NCCL / NVLink
    rec_block = None
    if world > 1:
        from gomatching_b200 import video as V
        schema = V.RecordSchema(max_instances=100)
        rec_block = torch.randint(0, 255, (F, schema.stride), dtype=torch.uint8, device=device)

    enc_events = []

    def step(record=False):
        outs = []
        for w in enc:
            if record:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            outs.append(launch(w, tuning))
            if record:
                b.record()
                enc_events.append((a, b))
        for w in dec:
            outs.append(launch(w))
        if rec_block is not None:
            V.gather_records(rec_block, F * world, dst=0)
        return outs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        step(record=True)
    t1.record()
    barrier()
    ms_total = t0.elapsed_time(t1)
    enc_ms = [a.elapsed_time(b) for a, b in enc_events]
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([ms_total], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * F / (ms_per_step * 1e-3)

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region -------------------------
    e2e = None
    if not args.no_e2e:
        keys = ("value", "ref", "offsets", "logits")
        host_in = [{k: torch.empty(w[k].shape, dtype=w[k].dtype, pin_memory=True).copy_(w[k]) for k in keys}
                   for w in enc + dec]
        host_out = [torch.empty((F, w["Lq"], M * D), dtype=torch.float32, pin_memory=True) for w in enc + dec]
        h2d = sum(t.numel() * t.element_size() for h in host_in for t in h.values())
        d2h = sum(t.numel() * t.element_size() for t in host_out)
        copy_stream = torch.cuda.Stream()

        def e2e_step():
            # copies for launch i+1 overlap the kernel of launch i (copy stream + events); results go back on the
            # compute stream's tail so every output byte reaches the host inside the step
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
                o = g.ms_deform_attn_forward_fused(d["value"], w["shapes"], w["lsi"], d["ref"], d["offsets"], d["logits"])
                for t in d.values():
                    t.record_stream(cur)
                ho.copy_(o, non_blocking=True)
            cur.synchronize()

        e2e_steps = max(3, min(args.steps, 10))
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
        e2e = {"value": world * F / e2e_s, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
               "api": "gomatching_b200.ms_deform_attn_forward_fused on pinned host tensors (H2D on a copy stream, D2H of every output)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
    b_alg = algorithmic_bytes(F, enc[0]["S"], enc[0]["Lq"])
    enc_mean_ms = sum(enc_ms) / len(enc_ms)
    achieved = b_alg / (enc_mean_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "msda_fwd_fast_kernel<float,32,...,FUSED> encoder launch (N=%d, Lq=S=%d)" % (F, enc[0]["S"]),
                "algorithmic_bytes_per_launch": b_alg, "mean_launch_us": enc_mean_ms * 1e3,
                "launches_timed": len(enc_ms), "peak_source": peak_src,
                "gather_bytes_per_launch": 4 * F * enc[0]["Lq"] * M * L * P * D * 4}
    # secondary bound: the SM's L1 data pipe returns at most one 128-byte row per clock per SM
    if clocks and clocks.get("sm_mhz"):
        sms = torch.cuda.get_device_properties(device).multi_processor_count
        rows = 4 * F * enc[0]["Lq"] * M * L * P
        roofline["l1_pipe"] = {"bound": "l1 data pipe (1 row/clk/SM)", "unit": "rows/clk/SM", "peak": 1.0,
                               "achieved": rows / (enc_mean_ms * 1e-3 * clocks["sm_mhz"] * 1e6 * sms)}
    traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(traffic_file):
        try:
            tr = json.load(open(traffic_file))
            roofline["traffic"] = tr.get("encoder_dram_bytes_per_launch_at_frames", {}).get(str(F))
            roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    cpu = None
    if not args.no_cpu_baseline:
        sec, te, td, threads = cpu_reference_frame_seconds(5, args.dist)
        cpu = {"value": 1.0 / sec, "unit": "frames/s", "cores": threads, "kind": "port",
               "sample": "median of 5: 1 encoder (%.3f s) + 1 decoder (%.3f s) call of one frame via F.grid_sample "
                         "(oracle.core_gridsample = ms_deform_attn_core_pytorch restated), scaled x6 each" % (te, td)}

    line = {
        "metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded, generated on device)",
        "config": workload_config(args, F), "clocks": clocks, "e2e": e2e,
        "gpu_launches": world * args.steps * (ENC_LAYERS + DEC_LAYERS), "roofline": roofline, "cpu_baseline": cpu,
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
