"""the consumers of the encoder output between AV-HuBERT and the Q-Former — audio feature
conv, fusion, query-length arithmetic, per-sample linear resize — CUDA path vs oracle/fusion_oracle.py (torch's own
Conv1d / F.interpolate called as src/model.py calls them)."""
import pytest
import torch
import torch.nn as nn

from oracle import fusion_oracle as fz


def test_query_lengths_host_arithmetic():
    from multimodalvc_b200 import fusion
    rates = [0.3, 1.0, 1.37, 2.0, 5.5]
    lens = [25, 150, 97, 600, 40]
    assert fusion.query_lengths(rates, lens, 4) == fz.query_lengths(rates, lens, 4)
    assert fusion.query_lengths(rates, lens, 3) == fz.query_lengths(rates, lens, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 2e-2)])
def test_audio_feature_conv_is_the_reference_conv1d(dtype, tol):
    from multimodalvc_b200 import fusion
    torch.manual_seed(0)
    for k in (2, 4):
        conv = nn.Conv1d(256, 256, kernel_size=k, stride=k, padding=0)
        m = fusion.AudioFeatureConv(256, k)
        m.load_state_dict(conv.state_dict(), strict=True)
        m = m.to("cuda", dtype)
        x = torch.randn(3, 51, 256)
        with torch.no_grad():
            y_ref = fz.afeat_conv(conv, x)
        y = m(x.to("cuda", dtype)).float().cpu()
        assert y.shape == y_ref.shape
        assert (y - y_ref).abs().max() < tol * y_ref.abs().max()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_resize_av_features_matches_per_sample_interpolate(dtype):
    from multimodalvc_b200 import fusion
    g = torch.Generator().manual_seed(3)
    av = torch.randn(5, 60, 96, generator=g).to(dtype)
    len_feat = [60, 41, 1, 17, 60]
    rates = [1.0, 1.37, 2.0, 1.93, 1.5]
    _, resized = fz.query_lengths(rates, len_feat, 4)
    ref, ref_mask = fz.resize(av.float(), len_feat, resized)
    out, mask = fusion.resize_av_features(av.cuda(), len_feat, resized)
    assert out.shape == ref.shape and mask.dtype == torch.int64
    assert torch.equal(mask.cpu(), ref_mask)
    tol = 1e-5 if dtype == torch.float32 else 1.6e-2
    assert (out.float().cpu() - ref).abs().max() < tol * ref.abs().max()
    for b, n in enumerate(resized):
        assert not out[b, int(n):].any()
    # fused features: slice + concat / add
    w = torch.randn(5, 75, 32, generator=g)
    assert torch.equal(fusion.fuse_av(w, av.float(), "concat"), fz.fuse(w, av.float(), "concat"))
    assert torch.equal(fusion.fuse_av(w[:, :, :32].repeat(1, 1, 3), av.float(), "add"), fz.fuse(w.repeat(1, 1, 3), av.float(), "add"))
