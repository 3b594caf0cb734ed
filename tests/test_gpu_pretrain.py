"""Pretraining-mode extras on the device: span-mask substitution (bit-exact copies), compute_logits
and the eval-mode masked-prediction forward against goldens of the REAL AVHubertModel.forward(mask=True)
(oracle/make_golden_pretrain.py) and against oracle/pretrain_oracle.py."""
import numpy as np
import pytest
import torch

from helpers import cosine, load_pretrain_case, make_pretrain_device_model, rel_err
from oracle import make_golden_pretrain as mg
from oracle import pretrain_oracle as po

pytestmark = pytest.mark.gpu


def _seed():
    np.random.seed(7)
    torch.manual_seed(7)


@pytest.mark.parametrize("name", ["input_other", "input_same", "input_b1"])
def test_apply_input_mask_is_bit_exact(name):
    c = load_pretrain_case(name)
    z = c["z"]
    m = make_pretrain_device_model(c, torch.float32)
    video, audio, pm = c["src"]["video"].cuda(), c["src"]["audio"].cuda(), c["pm"].cuda()
    keep_v, keep_a = video.clone(), audio.clone()
    _seed()
    v_m, mi_v = m.apply_input_mask(video, pm, None)
    a_m, mi_a = m.apply_input_mask(audio, pm, None)
    assert np.array_equal(mi_v.cpu().numpy(), z["mask_video"]) and np.array_equal(mi_a.cpu().numpy(), z["mask_audio"])
    assert np.array_equal(a_m.cpu().numpy(), z["audio_masked"])                      # mask_emb rows, everything else a copy
    assert np.array_equal(v_m.double().sum(dim=(-1, -2)).cpu().numpy(), z["video_masked_framesum"])
    assert np.array_equal(v_m[..., 0, :4].cpu().numpy(), z["video_masked_first_px"])
    assert torch.equal(video, keep_v) and torch.equal(audio, keep_a)                 # the caller's tensors are untouched
    assert v_m.is_contiguous() and a_m.is_contiguous() and v_m.shape == video.shape and a_m.shape == audio.shape
    # against the CPU oracle in full, incl. the collater's transposed audio view and half precision
    from multimodalvc_b200 import masking
    _seed()
    v_o, _ = po.apply_input_mask(c["head"], c["src"]["video"].clone(), c["pm"], masking.compute_mask_indices)
    a_o, _ = po.apply_input_mask(c["head"], c["src"]["audio"].clone(), c["pm"], masking.compute_mask_indices)
    assert torch.equal(v_m.cpu(), v_o) and torch.equal(a_m.cpu(), a_o)
    a_view = c["src"]["audio"].transpose(1, 2).contiguous().cuda().transpose(1, 2)     # [B,F,T] view of [B,T,F] storage
    _seed()
    m.apply_input_mask(video, pm, None)
    a_m2, _ = m.apply_input_mask(a_view, pm, None)
    assert torch.equal(a_m2, a_m)
    mh = make_pretrain_device_model(c, torch.bfloat16)
    _seed()
    v_h, _ = mh.apply_input_mask(video.bfloat16(), pm, None)
    a_h, _ = mh.apply_input_mask(audio.bfloat16(), pm, None)
    assert torch.equal(v_h.cpu(), v_o.bfloat16())
    if c["src"]["audio"].size(0) > 1:
        assert torch.equal(a_h.cpu(), torch.where(torch.from_numpy(z["mask_audio"])[:, None, :],
                                                  c["head"].mask_emb.bfloat16()[None, :, None],
                                                  c["src"]["audio"].bfloat16()))


def test_compute_logits_kernel_matches_oracle():
    from multimodalvc_b200 import AVHubertConfig, AVHubertModel
    m = AVHubertModel(AVHubertConfig.named("tiny")).cuda().eval()
    g = torch.Generator().manual_seed(3)
    for (B, T, F, V) in [(3, 37, 32, 23), (2, 150, 256, 504), (1, 1, 7, 1), (4, 64, 64, 128)]:
        f = torch.randn(B, T, F, generator=g)
        e = torch.randn(V, F, generator=g)
        f[0, 0] = 0                                            # zero-norm row: the 1e-6 clamp decides
        for sim in ("cosine", "dot"):
            ref = po.compute_logits(f, e, sim, 0.1)
            got = m.compute_logits(f.cuda(), e.cuda(), sim_type=sim, logit_temp=0.1).cpu()
            assert got.shape == ref.shape and got.dtype == torch.float32
            assert (got - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
        bias = torch.randn(V, generator=g)
        lin = m.compute_logits(f.cuda(), e.cuda(), bias=bias.cuda(), sim_type="dot", logit_temp=1.0).cpu()
        assert (lin - (f @ e.t() + bias)).abs().max().item() < 1e-4
        fb = f.bfloat16()
        got = m.compute_logits(fb.cuda(), e.cuda(), sim_type="cosine").cpu()
        assert (got - po.compute_logits(fb.float(), e, "cosine", 0.1)).abs().max().item() < 1e-4
    with pytest.raises(RuntimeError):
        m.compute_logits(torch.randn(2, 3, 8), torch.randn(4, 8))          # CPU tensors: no CPU path
    # the C entry point refuses an in-place substitution (sources must be read from the un-substituted tensor)
    import ctypes
    from multimodalvc_b200 import _lib
    x = torch.zeros(1, 4, 8, device="cuda")
    code = torch.full((4,), -1, dtype=torch.int32, device="cuda")
    vp = ctypes.c_void_p
    rc = _lib.load().avh_mask_substitute(vp(x.data_ptr()), _lib.AVH_F32, 0, None, 1, 4, 8, vp(code.data_ptr()), None, 0, None,
                                         vp(x.data_ptr()), _lib.AVH_F32, None)
    assert rc != 0 and b"out of place" in _lib.load().avh_last_error()


@pytest.mark.parametrize("name", list(mg.CASES))
def test_pretraining_forward_matches_real_reference_fp32(name):
    c = load_pretrain_case(name)
    z = c["z"]
    m = make_pretrain_device_model(c, torch.float32)
    src = {k: v.cuda() for k, v in c["src"].items()}
    targets = [t.cuda() for t in c["targets"]]
    _seed()
    res = m(src, target_list=targets, padding_mask=c["pm"].cuda(), mask=True, features_only=False)
    for i in range(c["n_dicts"]):
        assert np.array_equal(res["target_m_list"][i].cpu().numpy(), z[f"target_m{i}"])
        assert np.array_equal(res["target_u_list"][i].cpu().numpy(), z[f"target_u{i}"])
        for key, lg in (("logit_m", res["logit_m_list"][i]), ("logit_u", res["logit_u_list"][i])):
            ref = torch.from_numpy(z[f"{key}{i}"])
            assert lg.shape == ref.shape
            assert rel_err(lg.cpu(), ref) < 2e-3, (key, rel_err(lg.cpu(), ref))
    assert abs(res["features_pen"].item() - float(z["features_pen"])) < 2e-3 * float(z["features_pen"])
    assert torch.equal(res["padding_mask"].cpu(), c["pm"])
    _seed()
    fo = m(src, target_list=None, padding_mask=c["pm"].cuda(), mask=True, features_only=True, output_layer=1)
    assert rel_err(fo["x"].cpu(), torch.from_numpy(z["fo_x"])) < 2e-3
    assert rel_err(fo["features"].cpu(), torch.from_numpy(z["fo_features"])) < 2e-3
    # extract_finetune(mask=True): input masking only (hubert.py:696-699, feature masking is not applied there)
    from multimodalvc_b200 import masking
    _seed()
    y, _ = m.extract_finetune(src, c["pm"].cuda(), mask=True)
    head = c["head"]
    _seed()
    if head.masking_type == "input":
        v_o, _ = po.apply_input_mask(head, c["src"]["video"].clone(), c["pm"], masking.compute_mask_indices)
        a_o, _ = po.apply_input_mask(head, c["src"]["audio"].clone(), c["pm"], masking.compute_mask_indices)
    else:
        v_o, a_o = c["src"]["video"], c["src"]["audio"]
    y_o, _ = c["oracle"].extract_finetune({"video": v_o, "audio": a_o}, c["pm"])
    assert rel_err(y.cpu(), y_o) < 2e-3


def test_pretraining_forward_bf16_and_errors():
    c = load_pretrain_case("input_other")
    z = c["z"]
    m = make_pretrain_device_model(c, torch.bfloat16)
    src = {k: v.cuda().bfloat16() for k, v in c["src"].items()}
    targets = [t.cuda() for t in c["targets"]]
    _seed()
    res = m(src, target_list=targets, padding_mask=c["pm"].cuda(), mask=True)
    assert np.array_equal(res["target_m_list"][0].cpu().numpy(), z["target_m0"])
    assert cosine(res["logit_m_list"][0].cpu(), torch.from_numpy(z["logit_m0"])) > 0.995
    assert cosine(res["logit_u_list"][0].cpu(), torch.from_numpy(z["logit_u0"])) > 0.995
    with pytest.raises(TypeError):
        m(src, target_list=targets, padding_mask=None, mask=True)
    with pytest.raises(ValueError):
        m({"audio": None, "video": src["video"]}, target_list=targets, padding_mask=c["pm"].cuda())
    m.train()
    with pytest.raises(NotImplementedError):
        m(src, target_list=targets, padding_mask=c["pm"].cuda())
    m.eval()
    m.remove_pretraining_modules()
    _seed()
    with pytest.raises(RuntimeError):
        m(src, target_list=targets, padding_mask=c["pm"].cuda())
