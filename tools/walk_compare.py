"""A/B timing of the encoder-shape launch: TMA window kernel with the frame-group walk, with round 1's single group
(tuning walk=1), and the register-gather kernel.  Per-launch CUDA events, rotating L2-cold buffer sets, 3 rounds each.
    python tools/walk_compare.py"""
import sys, statistics, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import gomatching_b200 as g, bench
dev = torch.device("cuda", 0)
CFG = (("window, frame groups", dict(mode=5)), ("window, one group (r1)", dict(mode=5, walk=1)), ("auto (None)", None),
       ("register gather", dict(mode=1, variant=3, tile_q=64, ctas_per_sm=4)))
for F, (H, W) in ((8, (720, 1280)), (4, (1080, 1920))):
    sets = [bench.device_workload("encoder", F, 300 + i, "local", dev, H, W) for i in range(6)]
    for w in sets:
        w["loc"], w["attn"] = g.locations_softmax(w["shapes"], w["ref"], w["offsets"], w["logits"], 8)
    for fused in (1, 0):
        for name, tn in CFG:
            if fused:
                fn = lambda w: g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"], tuning=tn)
            else:
                fn = lambda w: g.ms_deform_attn_forward(w["value"], w["shapes"], w["lsi"], w["loc"], w["attn"], 64, tuning=tn)
            rounds = []
            for r in range(3):
                rounds.append(bench.time_launches(fn, sets, 3, warm=1))
            print("F=%d %dx%d fused=%d  %-24s %s us" % (F, H, W, fused, name, "  ".join("%.1f" % x for x in rounds)))
