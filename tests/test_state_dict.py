"""Drop-in surface: state-dict key names, module attributes and wrapper output dict."""
import torch

from multimodalvc_b200 import AVHubertConfig, AVHubertModel, HubertEncoderWrapper
from oracle import avhubert_oracle as ao


def test_state_dict_keys_match_reference_names():
    m = AVHubertModel(AVHubertConfig.named("tiny"))
    keys = set(m.state_dict().keys())
    o = ao.build_oracle("tiny")
    assert set(o.state_dict().keys()) <= keys           # oracle keys == reference keys (pinned elsewhere)
    for k in ["feature_extractor_video.resnet.frontend3D.0.weight",
              "feature_extractor_video.resnet.frontend3D.1.running_var",
              "feature_extractor_video.resnet.frontend3D.2.weight",
              "feature_extractor_video.resnet.trunk.layer2.0.downsample.0.weight",
              "feature_extractor_video.resnet.trunk.layer4.1.relu2.weight",
              "feature_extractor_audio.proj.bias", "layer_norm.weight", "post_extract_proj.weight",
              "encoder.pos_conv.0.weight_g", "encoder.pos_conv.0.weight_v", "encoder.pos_conv.0.bias",
              "encoder.layers.1.self_attn.q_proj.weight", "encoder.layers.0.final_layer_norm.bias",
              "encoder.layer_norm.weight", "mask_emb", "final_proj.weight"]:
        assert k in keys, k


def test_reference_style_load_and_freeze():
    m = AVHubertModel.build_model(AVHubertConfig.named("tiny"))
    state = {"model": ao.build_oracle("tiny").state_dict()}
    res = m.load_state_dict(state["model"], strict=False)       # src/model.py:224
    assert not res.unexpected_keys
    m.remove_pretraining_modules()                              # src/model.py:226
    assert m.final_proj is None and "final_proj.weight" not in m.state_dict()
    for p in m.parameters():                                    # src/model.py:96-97
        p.requires_grad = False
    assert m.encoder.embedding_dim == 128                       # hubert_asr.py:309
    assert m.encoder.pos_conv[0].weight_g.shape == (1, 1, 128)


def test_param_count_matches_survey():
    base = AVHubertModel(AVHubertConfig.named("base"))
    base.remove_pretraining_modules()
    n = sum(p.numel() for k, p in base.named_parameters() if k != "mask_emb")
    assert abs(n - 102.6e6) < 0.3e6            # Base 102.6 M


def test_forward_padding_mask_semantics():
    pm = torch.tensor([[False, False, True, True, True], [False] * 5])
    assert torch.equal(AVHubertModel.forward_padding_mask(5, pm), pm)
    pm2 = torch.tensor([[False, False, False, True, True, True, True]])     # 7 % 3 = 1 extra -> trimmed
    out = AVHubertModel.forward_padding_mask(3, pm2)
    assert out.tolist() == [[False, False, True]]


def test_wrapper_contract():
    class Fake(torch.nn.Module):
        def extract_finetune(self, source, padding_mask):
            return torch.zeros(2, 5, 8), padding_mask
    w = HubertEncoderWrapper(Fake())
    pm = torch.zeros(2, 5, dtype=torch.bool)
    out = w(source={"audio": None, "video": None}, padding_mask=pm)
    assert set(out) == {"encoder_out", "encoder_padding_mask", "padding_mask"}
    assert out["encoder_out"].shape == (5, 2, 8)
    out = w.reorder_encoder_out(out, torch.tensor([1, 0]))
    assert out["encoder_out"].shape == (5, 2, 8)


def test_reorder_encoder_out_variants_follow_the_reference():
    """avhubert/hubert_asr.py:396-409 (wrapper: all three entries reordered) vs :356-365 (HubertEncoder:
    'padding_mask' left alone)."""
    from multimodalvc_b200 import HubertEncoder
    pm = torch.tensor([[True, False], [False, False]])
    x = torch.arange(2 * 2 * 3, dtype=torch.float32).view(2, 2, 3)           # T x B x C
    order = torch.tensor([1, 0])
    d = {"encoder_out": x, "encoder_padding_mask": pm, "padding_mask": pm}
    out = HubertEncoderWrapper.reorder_encoder_out(None, dict(d), order)
    assert torch.equal(out["encoder_out"], x[:, [1, 0]])
    assert torch.equal(out["encoder_padding_mask"], pm[[1, 0]]) and torch.equal(out["padding_mask"], pm[[1, 0]])
    out = HubertEncoder.reorder_encoder_out(None, dict(d), order)
    assert torch.equal(out["encoder_padding_mask"], pm[[1, 0]]) and torch.equal(out["padding_mask"], pm)
