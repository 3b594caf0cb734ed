"""Worker of test_overlapped_gradient_allreduce (2 ranks, NCCL): the gradient all-reduce issued bucket by bucket from inside
the device backward (GradientAllReducer.attach) gives the gradients the post-backward reduction gives, for the frozen-
extractor and the whole-model fine-tuning step.  Launched with torch.distributed.run; prints OK on rank 0."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    from multimodalvc_b200.distributed import GradientAllReducer
    from oracle import avhubert_oracle as ao
    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl")
    dev = torch.device("cuda", local)
    o = ao.build_oracle("tiny", seed=1234)
    B, T = 2, 24
    src, pm = ao.synthetic_inputs(B, T, lengths=[24, 17], seed=100 + rank)
    w = torch.randn(B, T, 128, generator=torch.Generator().manual_seed(5 + rank)).to(dev)
    dsrc = {k: v.to(dev, torch.bfloat16) for k, v in src.items()}
    for fgm in (0.0, 0.5):
        results = []
        for overlap in (False, True):
            cfg = AVHubertConfig.named("tiny", feature_grad_mult=fgm, trainable=True, dropout=0.0, attention_dropout=0.0,
                                       activation_dropout=0.0, encoder_layerdrop=0.0, dropout_input=0.0)
            m = AVHubertModel(cfg)
            m.remove_pretraining_modules()
            m.load_state_dict(o.state_dict(), strict=False)
            m = m.to(dev, torch.bfloat16).train()
            params = m.full_parameters(True, True)[0] if fgm > 0 else m.tail_parameters()
            red = GradientAllReducer(params)
            if overlap:
                red.attach(m)
            for step in range(2):                      # second step: graphs replayed, buckets re-used
                for p in params:
                    p.grad = None
                y, _ = m.extract_finetune(dsrc, pm.to(dev))
                ((y.float() * w) * (~pm.to(dev)).unsqueeze(-1)).sum().backward()
                if overlap:
                    assert len(red._presynced) == len(params), (len(red._presynced), len(params))
                issued = red.all_reduce_grads()
                assert (issued == 0) == overlap, (issued, overlap)
            torch.cuda.synchronize()
            results.append([p.grad.detach().float().clone() for p in params])
        for i, (a, b) in enumerate(zip(*results)):
            assert torch.equal(a, b), (fgm, i, (a - b).abs().max().item())
        # and they are averages: identical on both ranks
        probe = torch.stack([g.flatten()[:8] for g in results[1][:4]]).contiguous()
        both = [torch.empty_like(probe) for _ in range(2)]
        dist.all_gather(both, probe)
        assert torch.equal(both[0], both[1])
    dist.barrier()
    if rank == 0:
        print("OK")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
