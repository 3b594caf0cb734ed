"""How long does a fine-tuning step take INCLUDING the weight refresh after an optimizer step (Large, 8 x 150, bf16)?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from multimodalvc_b200 import AVHubertConfig, AVHubertModel
dev = torch.device("cuda", 0)
fgm = float(sys.argv[1]) if len(sys.argv) > 1 else 0.1
m = AVHubertModel(AVHubertConfig.named("large", feature_grad_mult=fgm, trainable=True, dropout=0.0, attention_dropout=0.0,
                                       activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0))
m.remove_pretraining_modules()
m = m.to(dev, torch.bfloat16).train()
v = torch.randn(8, 1, 150, 88, 88, device=dev).bfloat16()
a = torch.randn(8, 104, 150, device=dev).bfloat16()
params = m.full_parameters(True, True)[0] if fgm > 0 else m.tail_parameters()
opt = torch.optim.SGD(params, lr=1e-3)
for i in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    y, _ = m.extract_finetune({"audio": a, "video": v}, None)
    loss = y.float().pow(2).mean()
    loss.backward()
    torch.cuda.synchronize(); t1 = time.perf_counter()
    opt.step(); opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"step {i}: forward(+refresh)+backward {1e3 * (t1 - t0):.1f} ms, optimizer {1e3 * (t2 - t1):.1f} ms, loss {loss.item():.5f}")
