"""Mirrors of the reference's encoder wrappers: ``HubertEncoderWrapper`` (avhubert/hubert_asr.py:375-409), the
object MMS-LLaMA owns as ``self.avhubert`` (src/model.py:96-97,220) and calls as
``self.avhubert(source={...}, padding_mask=...)``, and ``HubertEncoder`` (avhubert/hubert_asr.py:251-372), the
encoder of the CTC / seq2seq fine-tuning models."""
import ctypes

import torch
import torch.nn as nn

from . import _lib


class HubertEncoderWrapper(nn.Module):
    def __init__(self, w2v_model):
        super().__init__()
        self.w2v_model = w2v_model

    def forward(self, source, padding_mask, **kwargs):
        """avhubert/hubert_asr.py:380-394: B x T x C -> T x B x C view + the three-key dict."""
        x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask)
        x = x.transpose(0, 1)
        return {
            "encoder_out": x,                        # T x B x C
            "encoder_padding_mask": padding_mask,    # B x T
            "padding_mask": padding_mask,
        }

    def reorder_encoder_out(self, encoder_out, new_order):
        """avhubert/hubert_asr.py:396-409 (the wrapper reorders all three entries)"""
        if encoder_out["encoder_out"] is not None:
            encoder_out["encoder_out"] = encoder_out["encoder_out"].index_select(1, new_order)
        if encoder_out["encoder_padding_mask"] is not None:
            encoder_out["encoder_padding_mask"] = encoder_out["encoder_padding_mask"].index_select(0, new_order)
        if encoder_out["padding_mask"] is not None:
            encoder_out["padding_mask"] = encoder_out["padding_mask"].index_select(0, new_order)
        return encoder_out


class HubertEncoder(nn.Module):
    """avhubert/hubert_asr.py:251-372 around an already-built ``AVHubertModel`` (the reference builds it from the
    checkpoint's config through the fairseq task, :299-305; here the caller passes it in).  ``forward`` =
    extract_finetune (no grad: the device path is inference-only) -> T x B x C -> final_dropout (identity in eval)
    -> optional ``proj`` (CTC vocabulary head ``Linear(d, len(tgt_dict))`` or ``Linear(d, decoder_embed_dim)``).
    The projection runs on the library's tcgen05 GEMM (bf16 operands, fp32 accumulation)."""

    def __init__(self, w2v_model, tgt_dict_size=None, decoder_embed_dim=None, final_dropout=0.0, apply_mask=False,
                 freeze_finetune_updates=0):
        super().__init__()
        w2v_model.remove_pretraining_modules()
        d = w2v_model.encoder.embedding_dim
        self.w2v_model = w2v_model
        self.apply_mask = apply_mask
        self.final_dropout = nn.Dropout(final_dropout)
        self.freeze_finetune_updates = freeze_finetune_updates
        self.num_updates = 0
        if tgt_dict_size is not None:
            self.proj = _linear(d, tgt_dict_size)
        elif decoder_embed_dim is not None and decoder_embed_dim != d:
            self.proj = _linear(d, decoder_embed_dim)
        else:
            self.proj = None

    def set_num_updates(self, num_updates):
        self.num_updates = num_updates

    @torch.no_grad()
    def forward(self, source, padding_mask, tbc=True, **kwargs):
        if self.training:
            raise RuntimeError("HubertEncoder on the device path is inference-only: call .eval()")
        x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask, mask=False)
        if self.proj is not None:
            x = _project(x, self.proj)
        if tbc:
            x = x.transpose(0, 1)                 # B x T x C -> T x B x C
        return {
            "encoder_out": x,
            "encoder_padding_mask": padding_mask,
            "padding_mask": padding_mask,
        }

    def reorder_encoder_out(self, encoder_out, new_order):
        """avhubert/hubert_asr.py:356-365: unlike the wrapper's, "padding_mask" is left as it is"""
        if encoder_out["encoder_out"] is not None:
            encoder_out["encoder_out"] = encoder_out["encoder_out"].index_select(1, new_order)
        if encoder_out["encoder_padding_mask"] is not None:
            encoder_out["encoder_padding_mask"] = encoder_out["encoder_padding_mask"].index_select(0, new_order)
        return encoder_out

    def max_positions(self):
        return None


def _linear(in_features, out_features):
    """fairseq-style Linear (hubert_asr.py:644-649): xavier_uniform weight, zero bias."""
    m = nn.Linear(in_features, out_features)
    nn.init.xavier_uniform_(m.weight)
    nn.init.constant_(m.bias, 0.0)
    return m


def _packed_head(lin, device):
    """bf16 weight planes [Np, D] (hi, mid = bf16 of the rounding residual) / fp32 bias [Np] of a projection head,
    columns padded to a multiple of 32, cached on the module until its parameters change (version counters) or move."""
    key = (lin.weight._version, lin.bias._version, lin.weight.data_ptr(), str(device))
    cache = getattr(lin, "_avh_packed", None)
    if cache is None or cache[0] != key:
        N, D = lin.out_features, lin.in_features
        Np = (N + 31) // 32 * 32
        wf = torch.zeros(Np, D, device=device, dtype=torch.float32)
        wf[:N] = lin.weight.detach().to(device=device, dtype=torch.float32)
        hi = wf.to(torch.bfloat16)
        mid = (wf - hi.float()).to(torch.bfloat16)
        bias = torch.zeros(Np, device=device, dtype=torch.float32)
        bias[:N] = lin.bias.detach().to(device=device, dtype=torch.float32)
        cache = (key, hi, mid, bias)
        lin._avh_packed = cache
    return cache[1], cache[2], cache[3]


def _project(x, lin):
    """[B,T,D] @ W^T + b on the library's tcgen05 GEMM (avh_gemm_bf16).  Half / bf16 modules: one bf16 product.  fp32
    modules: the split-precision form of the library's fp32 mode — operands as bf16 hi + mid planes, three products
    (mid*hi, hi*mid, hi*hi) chained through the fp32 residual input, ~2^-16 relative error per product."""
    B, T, D = x.shape
    N = lin.out_features
    w_hi, w_mid, bias = _packed_head(lin, x.device)
    Np = w_hi.size(0)
    xf = x.reshape(B * T, D)
    a_hi = xf.to(torch.bfloat16).contiguous()
    out = torch.empty(B * T, Np, device=x.device, dtype=torch.float32)
    vp = ctypes.c_void_p
    lib = _lib.load()
    with torch.cuda.device(x.device):
        stream = vp(torch.cuda.current_stream(x.device).cuda_stream)

        def gemm(a, w, b, res):
            _lib.check(lib.avh_gemm_bf16(vp(a.data_ptr()), vp(w.data_ptr()), B * T, Np, D, vp(b.data_ptr()) if b is not None else None,
                                         0, vp(out.data_ptr()) if res else None, 1, vp(out.data_ptr()), 1, 0, 0, 0, stream))
        if x.dtype == torch.float32:
            a_mid = (xf.float() - a_hi.float()).to(torch.bfloat16).contiguous()
            gemm(a_mid, w_hi, None, False)
            gemm(a_hi, w_mid, None, True)
            gemm(a_hi, w_hi, bias, True)
        else:
            gemm(a_hi, w_hi, bias, False)
    return out[:, :N].to(x.dtype).view(B, T, N)
