"""Tilings of the coarse-level (query levels 1..3) part of the default encoder launch.  python tools/rest_compare.py"""
import sys, torch
sys.path.insert(0, __file__.rsplit("/", 2)[0])
import gomatching_b200 as g, bench
dev = torch.device("cuda", 0)
sets = [bench.device_workload("encoder", 8, 300 + i, "local", dev) for i in range(6)]
for name, tn in (("linear 64 (default)", dict(mode=5)), ("linear 32", dict(mode=5, tile_q=32)), ("linear 128", dict(mode=5, tile_q=128)),
                 ("pyramid 8x8", dict(mode=5, tile_h=8, tile_w=8)), ("pyramid 4x16", dict(mode=5, tile_h=4, tile_w=16)),
                 ("pyramid 8x16", dict(mode=5, tile_h=8, tile_w=16)), ("pyramid 4x8", dict(mode=5, tile_h=4, tile_w=8)),
                 ("pyramid 16x8", dict(mode=5, tile_h=16, tile_w=8)), ("register gather all", dict(mode=1, variant=3, tile_q=64, ctas_per_sm=4))):
    fn = lambda w: g.ms_deform_attn_forward_fused(w["value"], w["shapes"], w["lsi"], w["ref"], w["offsets"], w["logits"], tuning=tn)
    print("%-24s %s us" % (name, "  ".join("%.1f" % bench.time_launches(fn, sets, 3, warm=1) for _ in range(2))))
