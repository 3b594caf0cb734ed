#!/usr/bin/env python
"""In-pipeline kernel timeline of the bench step (CUPTI activity records through torch.profiler): per-kernel
durations as they run back to back (warm L2, PDL overlap), and the idle gaps between consecutive kernels.
Diagnostic only — never a bench number.

  python tools/timeline.py [--steps 4] [--streams 1] [--out gpurun_out/timeline.json]
"""
import argparse
import collections
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--out", default=None)
    ap.add_argument("--fgm", type=float, default=0.0, help="with --train: feature_grad_mult (> 0 = whole-model step)")
    ap.add_argument("--train", action="store_true", help="profile the fine-tuning step of BASELINE config 5 (frozen extractors) instead")
    args = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    from multimodalvc_b200 import build as avh_build
    avh_build.build()
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    model = AVHubertModel(AVHubertConfig.named("large")).to(dev, torch.bfloat16).eval()
    model.remove_pretraining_modules()
    B, T = 16, 150
    vids = [torch.randn(B, 1, T, 88, 88, device=dev).bfloat16() for _ in range(4)]
    auds = [torch.randn(B, 104, T, device=dev).bfloat16() for _ in range(4)]
    streams = [torch.cuda.Stream(dev) for _ in range(args.streams)]

    def step(i):
        with torch.cuda.stream(streams[i % len(streams)]):
            model.extract_finetune({"audio": auds[i % 4], "video": vids[i % 4]}, None)

    if args.train:
        del model
        m5 = AVHubertModel(AVHubertConfig.named("large", feature_grad_mult=args.fgm, trainable=True, dropout=0.0, attention_dropout=0.0,
                                                activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0))
        m5.remove_pretraining_modules()
        m5 = m5.to(dev, torch.bfloat16).train()
        v5, a5 = vids[0][:8].contiguous(), auds[0][:8].contiguous()

        def step(i):
            y5, _ = m5.extract_finetune({"audio": a5, "video": v5}, None)
            y5.float().pow(2).mean().backward()
            for p5 in m5.parameters():
                p5.grad = None

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            step(i)
        torch.cuda.synchronize()
    evs = []
    for e in prof.events():
        if e.device_type is not None and str(e.device_type).endswith("CUDA") and e.time_range is not None:
            evs.append((e.time_range.start, e.time_range.end, e.name))
    evs.sort()
    if not evs:
        print("no CUDA activity records (CUPTI unavailable?)")
        return
    short = lambda n: re.sub(r"^void |avh::<unnamed>::|avh::\(anonymous namespace\)::|\(.*$", "", n)[:70]
    agg = collections.OrderedDict()
    gaps = collections.OrderedDict()
    busy_end = evs[0][0]
    total_gap = 0.0
    for i, (s, t, n) in enumerate(evs):
        a = agg.setdefault(short(n), [0, 0.0, 0.0])
        a[0] += 1
        a[1] += t - s
        a[2] += max(0.0, t - max(s, busy_end)) if i > 0 else t - s      # exclusive: after every earlier kernel ended
        if i > 0:
            gap = max(0.0, s - busy_end)
            total_gap += gap
            g = gaps.setdefault(short(n), [0, 0.0])
            g[0] += 1
            g[1] += gap
        busy_end = max(busy_end, t)
    span = evs[-1][1] - evs[0][0]
    ksum = sum(v[1] for v in agg.values())
    print(f"steps {args.steps} streams {args.streams}: span {span / args.steps:.1f} us/step, kernel sum "
          f"{ksum / args.steps:.1f} us/step, idle gaps {total_gap / args.steps:.1f} us/step, launches {len(evs) / args.steps:.0f}/step")
    print(f"{'kernel':72s} {'n/step':>7s} {'us/step':>9s} {'avg us':>8s} {'excl us/step':>13s} {'excl avg':>9s} {'gap avg us':>11s}")
    for k, (c, v, x) in sorted(agg.items(), key=lambda kv: -kv[1][2]):
        g = gaps.get(k, [1, 0.0])
        print(f"{k:72s} {c / args.steps:7.1f} {v / args.steps:9.1f} {v / c:8.2f} {x / args.steps:13.1f} {x / c:9.2f} {g[1] / max(1, g[0]):11.2f}")
    if args.out:
        with open(args.out, "w") as f:
            json.dump({"steps": args.steps, "streams": args.streams, "span_us_per_step": span / args.steps,
                       "kernel_sum_us_per_step": ksum / args.steps, "idle_us_per_step": total_gap / args.steps,
                       "kernels": {k: {"per_step": c / args.steps, "us_per_step": v / args.steps, "avg_us": v / c, "exclusive_us_per_step": x / args.steps,
                                       "gap_before_avg_us": gaps.get(k, [1, 0.0])[1] / max(1, gaps.get(k, [1, 0.0])[0])}
                                   for k, (c, v, x) in agg.items()}}, f, indent=1)


if __name__ == "__main__":
    main()
