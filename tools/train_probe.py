#!/usr/bin/env python
"""Workload for ncu captures of the fine-tuning step's kernels (BASELINE config 5): N steps, no profiler.
  python tools/train_probe.py [steps] [feature_grad_mult]      (0 = frozen extractors, > 0 = whole model)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    fgm = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    m = AVHubertModel(AVHubertConfig.named("large", feature_grad_mult=fgm, trainable=True, dropout=0.0, attention_dropout=0.0,
                                           activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0))
    m.remove_pretraining_modules()
    m = m.to(dev, torch.bfloat16).train()
    v = torch.randn(8, 1, 150, 88, 88, device=dev).bfloat16()
    a = torch.randn(8, 104, 150, device=dev).bfloat16()
    for _ in range(steps):
        y, _ = m.extract_finetune({"audio": a, "video": v}, None)
        y.float().pow(2).mean().backward()
        for p in m.parameters():
            p.grad = None
    torch.cuda.synchronize()
    print("ok", float(y.float().abs().mean()))


if __name__ == "__main__":
    main()
