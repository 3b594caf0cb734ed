"""TEST INFRASTRUCTURE ONLY — CPU restatement of AV-HuBERT's pretraining-mode extras on top of
``OracleAVHubert``: ``apply_input_mask`` (avhubert/hubert.py:442-494), ``apply_feature_mask`` (:496-536),
``compute_logits`` (:576-589) and the eval-mode masked-prediction ``forward`` (:591-674).

Pinned by tests/golden/pretrain_*.npz = outputs of the REAL reference (oracle/make_golden_pretrain.py).  The span
positions come from a ``draw`` callable with the signature of ``compute_mask_indices`` (avhubert/utils.py:142-270);
the tests pass the product's host function, which is itself pinned bit for bit (incl. the generator position) against
tests/golden/mask_indices.npz from the real function.  Never imported by the product.
"""
import numpy as np
import torch


class Head:
    """The parameters forward() uses besides the encoder's: mask_emb, final_proj, label embeddings, and the config."""

    def __init__(self, mask_emb, final_proj_w, final_proj_b, label_embs, num_classes, **cfg):
        self.mask_emb = torch.as_tensor(mask_emb).float()
        self.w, self.b = torch.as_tensor(final_proj_w).float(), torch.as_tensor(final_proj_b).float()
        self.label_embs = torch.as_tensor(label_embs).float()
        self.num_classes = list(num_classes)
        d = dict(mask_prob_audio=0.65, mask_length_audio=10, mask_prob_image=0.65, mask_length_image=10,
                 mask_selection="static", mask_other=0, no_mask_overlap=False, mask_min_space=1, mask_channel_prob=0.0,
                 mask_channel_length=10, mask_channel_selection="static", mask_channel_other=0,
                 no_mask_channel_overlap=False, mask_channel_min_space=1, logit_temp=0.1, sim_type="cosine",
                 selection_type="same_other_seq", masking_type="input", untie_final_proj=False)
        d.update(cfg)
        self.__dict__.update(d)


def apply_input_mask(head, x, padding_mask, draw):
    B, C, T = x.shape[:3]
    audio = x.dim() == 3
    prob, length = (head.mask_prob_audio, head.mask_length_audio) if audio else (head.mask_prob_image, head.mask_length_image)
    if not prob > 0:
        return x, None
    m, starts, ends, owners = draw((B, T), padding_mask, prob, length, head.mask_selection, head.mask_other, min_masks=2,
                                   no_overlap=head.no_mask_overlap, min_space=head.mask_min_space)
    mt = torch.from_numpy(m)
    y = x.transpose(1, 2).contiguous().clone()             # [B,T,C(,H,W)]
    if B == 1:
        y[mt] = 0
    elif audio:
        y[mt] = head.mask_emb.to(y.dtype)
    elif head.selection_type == "same_other_seq":
        shift = torch.randint(low=1, high=B, size=(1,))
        other = y[(torch.arange(B) + shift) % B]
        y[mt] = other[mt]
    elif head.selection_type == "same_seq":
        rows, cols = [], []
        for b, s, e in zip(owners, starts, ends):
            n = int(e - s)
            free = np.setdiff1d(np.arange(T), np.arange(max(0, s - n), e))
            first = int(np.random.choice(free, size=1)[0]) if len(free) else 0
            cols.append(np.arange(first, first + n).clip(max=T - 1))
            rows.append(np.full(n, b, dtype=np.int64))
        y[mt] = y[np.concatenate(rows), np.concatenate(cols)]
    else:
        raise ValueError(head.selection_type)
    return y.transpose(1, 2).contiguous(), mt


def apply_feature_mask(head, x, padding_mask, draw):
    B, T, C = x.shape
    x = x.clone()
    mt = None
    if head.mask_prob_audio > 0:
        m, _, _, _ = draw((B, T), padding_mask, head.mask_prob_audio, head.mask_length_image, head.mask_selection,
                          head.mask_other, min_masks=2, no_overlap=head.no_mask_overlap, min_space=head.mask_min_space)
        mt = torch.from_numpy(m)
        x[mt] = head.mask_emb.to(x.dtype)
    if head.mask_channel_prob > 0:
        mc, _, _, _ = draw((B, C), None, head.mask_channel_prob, head.mask_channel_length, head.mask_channel_selection,
                           head.mask_channel_other, no_overlap=head.no_mask_channel_overlap,
                           min_space=head.mask_channel_min_space)
        x[torch.from_numpy(mc).unsqueeze(1).expand(-1, T, -1)] = 0
    return x, mt


def compute_logits(feats, emb, sim_type="cosine", temp=0.1):
    feats, emb = feats.float(), emb.float()
    if sim_type == "dot":
        logits = feats @ emb.t()
    elif sim_type == "cosine":
        lead = feats.shape[:-1]
        f = feats.reshape(-1, feats.size(-1))
        num = f @ emb.t()
        den = f.pow(2).sum(-1).sqrt()[:, None] * emb.pow(2).sum(-1).sqrt()[None, :]
        logits = (num / den.clamp(min=1e-6)).view(*lead, -1)
    else:
        raise NotImplementedError
    return logits / temp


@torch.no_grad()
def forward(oracle, head, source, target_list, padding_mask, draw, mask=True, features_only=False, output_layer=None):
    """Eval-mode AVHubertModel.forward on the oracle encoder.  Label rate == frame rate (no target trimming)."""
    a, v = source["audio"], source["video"]
    mi = None
    if mask and head.masking_type == "input":
        v, mv = apply_input_mask(head, v, padding_mask, draw)
        a, ma = apply_input_mask(head, a, padding_mask, draw)
        mi = torch.logical_or(ma, mv)
    np.random.random(), np.random.random()                 # modality-dropout coins (hubert.py:611)
    fv = oracle.feature_extractor_video(v)
    fa = oracle.feature_extractor_audio(a)
    feats = torch.cat([fa, fv], dim=1) if oracle.cfg.modality_fuse == "concat" else fa + fv
    pen = feats.float().pow(2).mean()
    feats = oracle.layer_norm(feats.transpose(1, 2))
    pm = oracle.forward_padding_mask(feats, padding_mask) if padding_mask is not None else None
    if oracle.post_extract_proj is not None:
        feats = oracle.post_extract_proj(feats)
    if mask and head.masking_type == "feature":
        feats, mi = apply_feature_mask(head, feats, pm, draw)
    x = oracle.encoder(feats, pm, None if output_layer is None else output_layer - 1)
    if pm is not None:
        feats = feats.masked_fill(pm.unsqueeze(-1), 0.0)   # the encoder's in-place index_put (wav2vec2.py:869-870)
    if features_only:
        return {"x": x, "padding_mask": pm, "features": feats}
    proj = x @ head.w.t() + head.b
    projs = proj.chunk(len(head.num_classes), dim=-1) if head.untie_final_proj else [proj] * len(head.num_classes)
    embs = head.label_embs.split(head.num_classes, 0)
    logits = [compute_logits(p, e, head.sim_type, head.logit_temp).view(-1, n) for p, e, n in zip(projs, embs, head.num_classes)]
    sel_m, sel_u = (mi & ~pm).view(-1), (~mi & ~pm).view(-1)
    return {"logit_m_list": [lg[sel_m] for lg in logits], "logit_u_list": [lg[sel_u] for lg in logits],
            "target_m_list": [t.reshape(-1)[sel_m].long() for t in target_list],
            "target_u_list": [t.reshape(-1)[sel_u].long() for t in target_list],
            "padding_mask": pm, "features_pen": pen}
