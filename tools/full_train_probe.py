"""Debug probe: backward intermediates of the whole-model training step vs float64 autograd on the oracle (fp32 mode).
This is the tool that attributed the ResNet gradient differences to PReLU kink flips: the gradient entering
layer4.1.bn1 matches to 1e-5, the one leaving it differs in exactly the channels that hold a unit with |v| < 1e-5.
  python tools/full_train_probe.py        (needs a B200)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalvc_b200 import AVHubertConfig, AVHubertModel
from oracle import avhubert_oracle as ao
sys.path.insert(0, os.path.join(ROOT, "tests"))
o = ao.build_oracle("tiny", seed=1234).train().double()
o32 = ao.build_oracle("tiny", seed=1234)
B, T = 2, 16
src, pm = ao.synthetic_inputs(B, T, lengths=[16, 11], seed=19)
cfg = AVHubertConfig.named("tiny", feature_grad_mult=1.0, trainable=True, dropout=0.0, attention_dropout=0.0,
                           activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0)
m = AVHubertModel(cfg); m.remove_pretraining_modules(); m.load_state_dict(o32.state_dict(), strict=False); m = m.cuda().train()
g = torch.Generator().manual_seed(6)
w = torch.randn(B, T, 128, generator=g)
def loss(y, w, pm):
    return ((y * w) * (~pm).unsqueeze(-1)).sum()
blk = o.feature_extractor_video.resnet.trunk.layer4[1]
got = {}
def hook(name):
    def f(mod, gin, gout):
        got[name + "_in"] = gin[0]; got[name + "_out"] = gout[0]
    return f
blk.conv2.register_full_backward_hook(hook("conv2"))
blk.conv1.register_full_backward_hook(hook("conv1"))
fv = o.feature_extractor_video(src["video"].double())
fa = o.feature_extractor_audio(src["audio"].double())
feats = o.layer_norm(torch.cat([fa, fv], dim=1).transpose(1, 2))
if o.post_extract_proj is not None:
    feats = o.post_extract_proj(feats)
y_ref = o.encoder(feats, pm)
loss(y_ref, w.double(), pm).backward()
y, _ = m.extract_finetune({k: v.cuda() for k, v in src.items()}, pm.cuda())
loss(y, w.cuda(), pm.cuda()).backward()
N = B * T
def dev(name, C=512):
    return m.read_stage(name, N * 9 * C).view(N, 3, 3, C).permute(0, 3, 1, 2).cpu().double()
def rel(a, b):
    return ((a - b).abs().max() / b.abs().max()).item()
print("d_a1    (conv2 grad_in)", rel(dev("grad_layer4_1_conv2_in"), got["conv2_in"]))
print("d_raw1  (conv1 grad_out)", rel(dev("grad_layer4_1_conv1_out"), got["conv1_out"]))
d = dev("grad_layer4_1_conv2_in") - got["conv2_in"]
print("d_a1 err by pixel", d.abs().amax(dim=(0, 1)) / got["conv2_in"].abs().max())
print("d_a1 err by frame", d.abs().amax(dim=(1, 2, 3)) / got["conv2_in"].abs().max())
dr = dev("grad_layer4_1_conv1_out"); rr = got["conv1_out"]
d = dr - rr
print("d_raw1: per-channel mean of diff (abs max)", d.mean(dim=(0, 2, 3)).abs().max().item(), "per-channel std of diff max",
      d.std(dim=(0, 2, 3)).max().item(), "ref absmax", rr.abs().max().item())
print("sum over rows dev", dr.sum(dim=(0, 2, 3)).abs().max().item(), "ref", rr.sum(dim=(0, 2, 3)).abs().max().item())
print("err by frame", d.abs().amax(dim=(1, 2, 3)) / rr.abs().max())
print("err by pixel", d.abs().amax(dim=(0, 1)) / rr.abs().max())
ch = d.abs().amax(dim=(0, 2, 3)) / rr.abs().max()
print("bad channels", (ch > 1e-3).sum().item(), "of 512; worst", ch.argmax().item())
pn = dict(m.named_parameters()); po = dict(o.named_parameters())
for n in ["bn1.weight", "bn1.bias", "relu1.weight", "conv1.weight"]:
    k = "feature_extractor_video.resnet.trunk.layer4.1." + n
    a = pn[k].grad.cpu().double(); r = po[k].grad
    print(n, rel(a, r), "n bad", ((a - r).abs() > 1e-3 * r.abs().max()).sum().item(), "of", r.numel())
