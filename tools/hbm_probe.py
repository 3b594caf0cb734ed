"""Workload for the ncu capture of the HBM-bound kernels north_star names (profiles/r2_ncu_hbm_kernels.txt):
log-fbank + stack + LN + collate (fbank_kernel), noise mixing (noise_*), and the Large encoder's LayerNorms
(layernorm_f32_vec_kernel) inside one config-2 forward.  Inputs larger than L2 for the audio kernels
(256 clips x 6 s = 49 MB of int16 samples).

  python tools/hbm_probe.py            # plain run (must exit 0 before the ncu run of the same command)
  ncu --set full --clock-control none -k regex:'fbank_kernel|noise_|layernorm' -c 12 -o gpurun_out/r2_hbm python tools/hbm_probe.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodalvc_b200 import AVHubertConfig, AVHubertModel, audio  # noqa: E402

dev = torch.device("cuda", 0)
n_clips, n_samp, T = 256, 96000, 150
g = torch.Generator().manual_seed(5)
flat = (torch.randn(n_clips * n_samp, generator=g) * 3000).clamp(-32768, 32767).to(torch.int16).to(dev)
off = (torch.arange(n_clips + 1, dtype=torch.int64) * n_samp).to(dev)
vlen = torch.full((n_clips,), T, dtype=torch.int32, device=dev)
noise = torch.randn(100000, generator=g).mul(2000).to(dev)
for _ in range(2):
    mixed = audio.add_noise_packed(flat, off, noise, 0.0)
    a, pm = audio.logfbank_stack_collate_packed(mixed, off, T, vlen)
torch.cuda.synchronize()
if os.environ.get("AVH_PROBE_ENCODER", "1") != "0":
    torch.manual_seed(1234)
    m = AVHubertModel(AVHubertConfig.named("large"))
    m.remove_pretraining_modules()
    m = m.to(dev, torch.bfloat16).eval()
    v = torch.randn(16, 1, T, 88, 88, generator=g).to(dev, torch.bfloat16)
    os.environ.setdefault("AVH_GRAPHS", "0")
    y, _ = m.extract_finetune({"audio": a[:16].to(torch.bfloat16), "video": v}, None)
    torch.cuda.synchronize()
    print("probe ok", float(y.float().abs().mean()))
else:
    print("probe ok", float(a.abs().mean()))
