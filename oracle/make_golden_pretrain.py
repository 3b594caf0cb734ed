"""TEST INFRASTRUCTURE ONLY — golden vectors for the pretraining-mode extras from the REAL reference.

Runs in the build container only (needs /root/reference):  python -m oracle.make_golden_pretrain
  * tests/golden/mask_indices.npz   — outputs of the real ``compute_mask_indices`` (avhubert/utils.py:142-270) over a
    sweep of numpy seeds, shapes, padding masks and span-length distributions;
  * tests/golden/pretrain_<case>.npz — the real ``AVHubertModel.forward`` (avhubert/hubert.py:591-674) in eval mode with
    mask=True: input masking ('same_other_seq', 'same_seq', B = 1), feature masking with channel masks, tied / untied
    heads, cosine / dot logits.  Encoder weights = the seeded oracle's (rebuilt by the tests), head weights stored.
"""
import importlib.util
import os
import types
import warnings

import numpy as np
import torch

from oracle import avhubert_oracle as ao
from oracle import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def real_utils():
    spec = importlib.util.spec_from_file_location("avh_ref_utils", os.path.join(ref_import.REF, "avhubert", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


MASK_SWEEP = [  # (seed, B, T, ragged, prob, length, kind, other, min_masks)
    (s, 4, 50 + 7 * s, s % 2 == 0, (0.65, 0.3, 0.8)[s % 3], (10, 5, 3)[s % 3], kind, other, 2)
    for s in range(6) for kind, other in (("static", 0), ("uniform", 2), ("normal", 2.0), ("poisson", 0))
]


def sweep_padding(seed, B, T):
    lens = np.random.RandomState(1000 + seed).randint(T // 3, T + 1, size=B)
    lens[0] = T
    return torch.arange(T)[None, :] >= torch.from_numpy(lens)[:, None]


def make_mask_golden():
    ru = real_utils()
    out = {}
    for i, (seed, B, T, ragged, prob, length, kind, other, mm) in enumerate(MASK_SWEEP):
        pm = sweep_padding(seed, B, T) if ragged else None
        np.random.seed(seed)
        m, s, e, b = ru.compute_mask_indices((B, T), pm, prob, length, kind, other, min_masks=mm)
        out[f"m{i}"], out[f"s{i}"], out[f"e{i}"], out[f"b{i}"] = m, s, e, b
        out[f"next{i}"] = np.array(np.random.rand())          # the generator's position afterwards
    np.savez_compressed(os.path.join(OUT, "mask_indices.npz"), **out)
    print("mask_indices.npz:", len(MASK_SWEEP), "cases")


CASES = {
    # name: (B, T, lengths, cfg overrides, n_dicts)
    "input_other": (3, 40, [40, 33, 25], dict(mask_prob_image=0.5, mask_length_image=5, mask_prob_audio=0.5,
                                             mask_length_audio=5), 1),
    "input_same": (3, 36, [36, 30, 36], dict(selection_type="same_seq", mask_prob_image=0.4, mask_length_image=4,
                                            mask_prob_audio=0.6, mask_length_audio=6), 1),
    "input_b1": (1, 30, [30], dict(mask_prob_image=0.5, mask_length_image=5, mask_prob_audio=0.5, mask_length_audio=5,
                                   sim_type="dot"), 1),
    "feature": (3, 40, [40, 28, 35], dict(masking_type="feature", mask_prob_image=0.5, mask_length_image=5,
                                         mask_prob_audio=0.5, mask_length_audio=5, mask_channel_prob=0.3,
                                         mask_channel_length=8), 1),
    "untied": (2, 32, [32, 27], dict(untie_final_proj=True, mask_prob_image=0.5, mask_length_image=5,
                                    mask_prob_audio=0.5, mask_length_audio=5), 2),
}
NUM_CLASSES = [23, 17]
FINAL_DIM = 32


def build_real(over, n_dicts):
    hub, _ = ref_import.install()
    hub.compute_mask_indices = real_utils().compute_mask_indices
    cfg = hub.AVHubertConfig()
    cfg.encoder_layers, cfg.encoder_embed_dim, cfg.encoder_ffn_embed_dim, cfg.encoder_attention_heads = 2, 128, 256, 2
    cfg.audio_feat_dim, cfg.modality_fuse, cfg.layer_norm_first, cfg.label_rate = 104, "concat", True, 25
    cfg.final_dim = FINAL_DIM
    for k in ["dropout", "attention_dropout", "activation_dropout", "encoder_layerdrop", "dropout_input",
              "dropout_features"]:
        setattr(cfg, k, 0.0)
    for k, v in over.items():
        setattr(cfg, k, v)
    torch.manual_seed(99)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = hub.AVHubertModel(cfg, types.SimpleNamespace(sample_rate=25), [list(range(n)) for n in NUM_CLASSES[:n_dicts]])
    return model.eval()


def case_inputs(name, B, T, lengths, n_dicts):
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=21)
    g = torch.Generator().manual_seed(5)
    targets = [torch.randint(0, n, (B, T), generator=g) for n in NUM_CLASSES[:n_dicts]]
    return src, pm, targets


def make_forward_case(name):
    B, T, lengths, over, n_dicts = CASES[name]
    enc_over = {k: v for k, v in over.items() if k in ("modality_fuse",)}
    oracle = ao.build_oracle("tiny", seed=1234, **enc_over)
    ref = build_real(over, n_dicts)
    missing = ref.load_state_dict(oracle.state_dict(), strict=False)
    assert not missing.unexpected_keys, missing
    src, pm, targets = case_inputs(name, B, T, lengths, n_dicts)
    rec = {}
    with torch.no_grad():
        if over.get("masking_type", "input") == "input":
            # the tensor side alone, same seeds as the forward below
            np.random.seed(7)
            torch.manual_seed(7)
            # (clones: the reference's transpose(1,2).contiguous() of a [B,1,T,H,W] tensor is the SAME storage, so its
            # in-place assignment writes through to the caller's video tensor)
            v_m, mi_v = ref.apply_input_mask(src["video"].clone(), pm, None)
            a_m, mi_a = ref.apply_input_mask(src["audio"].clone(), pm, None)
            rec["video_masked_framesum"] = v_m.double().sum(dim=(-1, -2)).numpy()
            rec["video_masked_first_px"] = v_m[..., 0, :4].numpy()
            rec["audio_masked"] = a_m.numpy()
            rec["mask_video"], rec["mask_audio"] = mi_v.numpy(), mi_a.numpy()
        np.random.seed(7)
        torch.manual_seed(7)
        res = ref({k: t.clone() for k, t in src.items()}, target_list=targets, padding_mask=pm, mask=True, features_only=False)
        np.random.seed(7)
        torch.manual_seed(7)
        fo = ref({k: t.clone() for k, t in src.items()}, target_list=None, padding_mask=pm, mask=True, features_only=True,
                 output_layer=1)
    for i in range(n_dicts):
        rec[f"logit_m{i}"], rec[f"logit_u{i}"] = res["logit_m_list"][i].numpy(), res["logit_u_list"][i].numpy()
        rec[f"target_m{i}"], rec[f"target_u{i}"] = res["target_m_list"][i].numpy(), res["target_u_list"][i].numpy()
    rec["features_pen"] = np.array(res["features_pen"].item())
    rec["fo_x"], rec["fo_features"] = fo["x"].numpy(), fo["features"].numpy()
    rec["mask_emb"] = ref.mask_emb.detach().numpy()
    rec["final_proj_w"], rec["final_proj_b"] = ref.final_proj.weight.detach().numpy(), ref.final_proj.bias.detach().numpy()
    rec["label_embs"] = ref.label_embs_concat.detach().numpy()
    rec["checksum"] = np.array(ao_checksum(oracle))
    np.savez_compressed(os.path.join(OUT, f"pretrain_{name}.npz"), **rec)
    print(f"pretrain_{name}.npz: masked rows {[int(x.shape[0]) for x in res['logit_m_list']]}, "
          f"features_pen {res['features_pen'].item():.5f}")


def ao_checksum(oracle):
    from oracle.make_golden import state_checksum
    return state_checksum(oracle.state_dict())


def main():
    os.makedirs(OUT, exist_ok=True)
    make_mask_golden()
    for name in CASES:
        make_forward_case(name)


if __name__ == "__main__":
    main()
