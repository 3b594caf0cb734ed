import torch, sys
sys.path.insert(0, "/root/repo")
from oracle import avhubert_oracle as ao
from multimodalvc_b200 import AVHubertConfig, AVHubertModel, video, audio
from oracle import fbank_oracle as fo, video_oracle as vo
oracle = ao.build_oracle("tiny", seed=1234)
for dtype in (torch.bfloat16, torch.float32):
    m = AVHubertModel(AVHubertConfig.named("tiny")); m.remove_pretraining_modules()
    m.load_state_dict(oracle.state_dict(), strict=False); m = m.to("cuda", dtype).eval()
    for B, T, lens in [(2, 7, [7, 3]), (3, 33, None), (1, 1, None)]:
        src, pm = ao.synthetic_inputs(B, T, lengths=lens, seed=T)
        s = {k: v.cuda().to(dtype) for k, v in src.items()}
        for _ in range(3):
            y, _ = m.extract_finetune(s, pm.cuda() if pm is not None else None)
        torch.cuda.synchronize()
        print(dtype, B, T, float(y.float().abs().mean()))
    raw = torch.from_numpy(vo.synthetic_frames(2 * 5, 96, 96)).view(2, 1, 5, 96, 96).cuda()
    y, _ = m.extract_finetune({"audio": torch.randn(2, 104, 5, device="cuda", dtype=dtype), "video": raw}, None)
    torch.cuda.synchronize()
a, pmk = audio.logfbank_stack_collate([torch.from_numpy(fo.synthetic_wave(12800, 1))], video_lens=[20])
torch.cuda.synchronize()
print("sanitizer workload done")
