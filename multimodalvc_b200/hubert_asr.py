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


class _LinearFn(torch.autograd.Function):
    """y = x W^T + b on the library's GEMM, differentiable: dX = dY W, dW = dY^T X (both through the same GEMM entry
    point; the operand transposes are copies), db = column sums."""

    @staticmethod
    def forward(ctx, x, lin, weight, bias):
        ctx.save_for_backward(x, weight)
        ctx.lin = lin
        return _project(x, lin)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, T, D = x.shape
        N = weight.size(0)
        dy2 = dy.reshape(B * T, N)
        dx = _matmul_nt(dy2, weight.detach().t().contiguous()).view(B, T, D).to(x.dtype)        # [M,N] x [D,N]^T
        dw = _matmul_nt(dy2.t().contiguous(), x.reshape(B * T, D).t().contiguous()).to(weight.dtype)   # [N,M] x [D,M]^T
        db = dy2.float().sum(0).to(weight.dtype)
        return dx, None, dw, db


def _matmul_nt(a, b):
    """a [M,K] @ b[N,K]^T -> fp32 [M,N] on avh_gemm_bf16; fp32 inputs run in split precision (hi + mid planes, three
    products), K padded to 8 and N to 32 with zeros."""
    M, K = a.shape
    N = b.size(0)
    Kp, Np = (K + 7) // 8 * 8, (N + 31) // 32 * 32
    dev = a.device
    af = torch.zeros(M, Kp, device=dev, dtype=torch.float32)
    af[:, :K] = a.float()
    bf = torch.zeros(Np, Kp, device=dev, dtype=torch.float32)
    bf[:N, :K] = b.float()
    split = a.dtype == torch.float32
    a_hi, b_hi = af.to(torch.bfloat16), bf.to(torch.bfloat16)
    out = torch.empty(M, Np, device=dev, dtype=torch.float32)
    vp = ctypes.c_void_p
    lib = _lib.load()
    with torch.cuda.device(dev):
        stream = vp(torch.cuda.current_stream(dev).cuda_stream)

        def gemm(aa, bb, res):
            _lib.check(lib.avh_gemm_bf16(vp(aa.data_ptr()), vp(bb.data_ptr()), M, Np, Kp, None, 0,
                                         vp(out.data_ptr()) if res else None, 1, vp(out.data_ptr()), 1, 0, 0, 0, stream))
        if split:
            a_mid = (af - a_hi.float()).to(torch.bfloat16)
            b_mid = (bf - b_hi.float()).to(torch.bfloat16)
            gemm(a_mid, b_hi, False)
            gemm(a_hi, b_mid, True)
            gemm(a_hi, b_hi, True)
        else:
            gemm(a_hi, b_hi, False)
    return out[:, :N]


class _DropoutFn(torch.autograd.Function):
    """nn.Dropout on the library's Philox stream (avh_dropout): the mask is a pure function of (seed, site, element), so
    the backward applies the same call to the incoming gradient."""

    @staticmethod
    def forward(ctx, x, p, seed):
        ctx.p, ctx.seed = p, seed
        return _DropoutFn._apply(x, p, seed)

    @staticmethod
    def _apply(x, p, seed):
        y = x.contiguous().clone()
        from .hubert import _DTYPES
        with torch.cuda.device(y.device):
            stream = torch.cuda.current_stream(y.device).cuda_stream
            _lib.check(_lib.load().avh_dropout(ctypes.c_void_p(y.data_ptr()), _DTYPES[y.dtype], y.numel(), float(p),
                                               int(seed), 7, ctypes.c_void_p(stream)))
        return y

    @staticmethod
    def backward(ctx, dy):
        return _DropoutFn._apply(dy, ctx.p, ctx.seed), None, None


class HubertEncoder(nn.Module):
    """avhubert/hubert_asr.py:251-372 around an already-built ``AVHubertModel`` (the reference builds it from the
    checkpoint's config through the fairseq task, :299-305; here the caller passes it in).  ``forward`` =
    extract_finetune -> T x B x C -> final_dropout -> optional ``proj`` (CTC vocabulary head ``Linear(d,
    len(tgt_dict))`` or ``Linear(d, decoder_embed_dim)``) on the library's tcgen05 GEMM.  Training (:329-354): the
    backbone runs under no_grad until ``num_updates`` reaches ``freeze_finetune_updates``, afterwards with gradients
    (needs the backbone built with ``cfg.trainable`` and frozen feature extractors: SURVEY row A18); ``mask`` =
    ``apply_mask and self.training``; the head is always differentiable."""

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

    def forward(self, source, padding_mask, tbc=True, **kwargs):
        if not (self.training and torch.is_grad_enabled()):
            with torch.no_grad():
                x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask,
                                                                  mask=self.apply_mask and self.training)
                if self.training and self.final_dropout.p > 0:
                    x = _DropoutFn._apply(x, self.final_dropout.p, int(torch.randint(0, 2 ** 62, (1,)).item()))
                if self.proj is not None:
                    x = _project(x, self.proj)
        else:
            ft = self.freeze_finetune_updates <= self.num_updates
            if ft and not self.w2v_model.cfg.trainable and any(p.requires_grad for p in self.w2v_model.parameters()):
                raise RuntimeError("fine-tuning the backbone needs AVHubertConfig(trainable=True, feature_grad_mult=0): "
                                   "build it that way, or freeze its parameters (requires_grad_(False))")
            if ft and self.w2v_model.cfg.trainable:
                x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask,
                                                                  mask=self.apply_mask)
            else:
                with torch.no_grad():
                    x, padding_mask = self.w2v_model.extract_finetune(source=source, padding_mask=padding_mask,
                                                                      mask=self.apply_mask)
            if self.final_dropout.p > 0:
                x = _DropoutFn.apply(x, self.final_dropout.p, int(torch.randint(0, 2 ** 62, (1,)).item()))
            if self.proj is not None:
                x = _LinearFn.apply(x, self.proj, self.proj.weight, self.proj.bias)
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
