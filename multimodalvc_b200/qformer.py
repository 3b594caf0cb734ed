"""The Q-Former stage of MMS-LLaMA: ``compression_using_qformer`` (src/model.py:584-619) — resize
the fused AV features per clip, then ``Qformer.bert(query_embeds=query_tokens, attention_mask, encoder_hidden_states,
encoder_attention_mask)`` (src/sub_model/Qformer.py:805-968) — on ``libavh_b200.so``.

``QFormerCompressor`` carries the reference's attribute / state-dict layout for this stage (``Qformer.bert.*``,
``query_tokens``: the keys of the MMS-LLaMA checkpoint, src/model.py:121-132); its torch modules are parameter
containers, every Linear runs on the tcgen05 GEMM and the attention on the library's kernels through
``avh_qformer_forward``.  Unused members of the reference's ``BertLMHeadModel`` (word / position embeddings, the text
branch ``intermediate`` / ``output``, the ``cls`` head) are not allocated; ``load_state_dict(strict=False)`` skips them.
"""
import ctypes
from dataclasses import dataclass
from typing import Sequence

import torch
import torch.nn as nn

from . import _lib
from .fusion import resize_av_features
from .hubert import _DTYPES


@dataclass
class QFormerConfig:
    """bert-large-uncased fields the query path uses, as src/model.py:121-127 sets them."""
    hidden_size: int = 1024             # cfg.qformer_dim
    num_hidden_layers: int = 2          # cfg.qformer_layers
    num_attention_heads: int = 16
    intermediate_size: int = 4096
    encoder_width: int = 2048           # fused AV feature dim (concat of Whisper 1024 + AV-HuBERT 1024)
    query_length: int = 120             # max_queries = queries_per_sec * 20 (* 2 with the speech-rate predictor)
    initializer_range: float = 0.02
    layer_norm_eps: float = 1e-12
    compute_dtype: str = "auto"         # "auto": fp32 module -> split-precision fp32 mode, half / bf16 -> bf16 mode


class _SelfParams(nn.Module):
    def __init__(self, dim, kv_dim):
        super().__init__()
        self.query = nn.Linear(dim, dim)
        self.key = nn.Linear(kv_dim, dim)
        self.value = nn.Linear(kv_dim, dim)


class _OutParams(nn.Module):
    def __init__(self, din, dim, eps):
        super().__init__()
        self.dense = nn.Linear(din, dim)
        self.LayerNorm = nn.LayerNorm(dim, eps=eps)


class _AttnParams(nn.Module):
    def __init__(self, dim, kv_dim, eps):
        super().__init__()
        self.self = _SelfParams(dim, kv_dim)
        self.output = _OutParams(dim, dim, eps)


class _Dense(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.dense = nn.Linear(din, dout)


class _LayerParams(nn.Module):          # BertLayer, query branch (Qformer.py:379-401)
    def __init__(self, c):
        super().__init__()
        self.attention = _AttnParams(c.hidden_size, c.hidden_size, c.layer_norm_eps)
        self.crossattention = _AttnParams(c.hidden_size, c.encoder_width, c.layer_norm_eps)
        self.intermediate_query = _Dense(c.hidden_size, c.intermediate_size)
        self.output_query = _OutParams(c.intermediate_size, c.hidden_size, c.layer_norm_eps)


class _Embeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.LayerNorm = nn.LayerNorm(c.hidden_size, eps=c.layer_norm_eps)


class _Encoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.layer = nn.ModuleList([_LayerParams(c) for _ in range(c.num_hidden_layers)])


class _Bert(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.embeddings = _Embeddings(c)
        self.encoder = _Encoder(c)


class _QformerParams(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bert = _Bert(c)


class QFormerCompressor(nn.Module):
    def __init__(self, cfg: QFormerConfig):
        super().__init__()
        if cfg.hidden_size != 64 * cfg.num_attention_heads:
            raise NotImplementedError("the device path needs attention heads of 64 channels (1024 / 16 in every shipped config)")
        if cfg.encoder_width % 64 or cfg.intermediate_size % 64:
            raise ValueError("encoder_width and intermediate_size must be multiples of 64")
        self.cfg = cfg
        self.Qformer = _QformerParams(cfg)
        self.query_tokens = nn.Parameter(torch.zeros(1, cfg.query_length, cfg.hidden_size))
        self.query_tokens.data.normal_(mean=0.0, std=cfg.initializer_range)
        for m in self.Qformer.modules():            # BertPreTrainedModel._init_weights (Qformer.py:665-675)
            if isinstance(m, nn.Linear):
                m.weight.data.normal_(mean=0.0, std=cfg.initializer_range)
                m.bias.data.zero_()
        self._handle = None
        self._handle_key = None
        self._dirty = True
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._mark_dirty())

    def _mark_dirty(self):
        self._dirty = True

    def _apply(self, fn, *args, **kwargs):
        out = super()._apply(fn, *args, **kwargs)
        self._dirty = True
        return out

    def refresh_weights(self):
        self._dirty = True

    def __del__(self):
        try:
            if self._handle is not None:
                _lib.load().avh_destroy(self._handle)
        except Exception:
            pass

    def _ensure_handle(self):
        p = self.query_tokens
        if p.device.type != "cuda":
            raise RuntimeError("QFormerCompressor computes on a B200 only: move the module to a CUDA device")
        cd = self.cfg.compute_dtype
        mode = ((_lib.AVH_COMPUTE_FP32 if p.dtype == torch.float32 else _lib.AVH_COMPUTE_BF16) if cd == "auto"
                else (_lib.AVH_COMPUTE_FP32 if cd in ("fp32", "float32") else _lib.AVH_COMPUTE_BF16))
        key = (p.device.index if p.device.index is not None else torch.cuda.current_device(), mode)
        if self._handle is not None and key == self._handle_key and not self._dirty:
            return self._handle
        lib = _lib.load()
        c = self.cfg
        if self._handle is None or key != self._handle_key:
            if self._handle is not None:
                lib.avh_destroy(self._handle)
                self._handle = None
            cc = _lib.AvhConfig(
                encoder_layers=c.num_hidden_layers, encoder_embed_dim=c.hidden_size,
                encoder_ffn_embed_dim=c.intermediate_size, encoder_attention_heads=c.num_attention_heads,
                audio_feat_dim=0, modality_fuse=_lib.AVH_FUSE_ADD, layer_norm_first=0, conv_pos=128, conv_pos_groups=16,
                compute_mode=mode, frontend_chunk_frames=0, capture_stages=0)
            cc.reserved[0], cc.reserved[1], cc.reserved[2] = 2, c.encoder_width, c.query_length
            hp = ctypes.c_void_p()
            _lib.check(lib.avh_create(ctypes.byref(cc), key[0], ctypes.byref(hp)))
            self._handle, self._handle_key = hp, key
        with torch.no_grad():
            tensors = {k[len("Qformer.bert."):]: v for k, v in self.state_dict().items() if k.startswith("Qformer.bert.")}
            tensors["query_tokens"] = self.query_tokens
            for name, t in tensors.items():
                t = t.detach().contiguous()
                if t.dtype not in _DTYPES:
                    t = t.float()
                shape = (ctypes.c_int64 * max(t.dim(), 1))(*t.shape)
                _lib.check(lib.avh_load_tensor(self._handle, name.encode(), ctypes.c_void_p(t.data_ptr()),
                                               _DTYPES[t.dtype], shape, t.dim()))
            torch.cuda.synchronize(p.device)
        _lib.check(lib.avh_finalize_weights(self._handle))
        _lib.check(lib.avh_drop_host_weights(self._handle))
        self._dirty = False
        return self._handle

    @torch.no_grad()
    def bert(self, len_queries: Sequence[int], encoder_hidden_states, encoder_attention_mask=None):
        """Qformer.bert(query_embeds=query_tokens.expand(B,-1,-1)[:, :max(len_queries)], attention_mask = ones up to
        len_queries[b], encoder_hidden_states, encoder_attention_mask (1 = valid) )['last_hidden_state'] -> [B, Lq, H]."""
        enc = encoder_hidden_states
        if enc.dim() != 3 or enc.size(2) != self.cfg.encoder_width:
            raise ValueError(f"encoder_hidden_states must be [B,T,{self.cfg.encoder_width}], got {tuple(enc.shape)}")
        B, Lk = enc.size(0), enc.size(1)
        if len(len_queries) != B:
            raise ValueError("one query length per sample expected")
        lq = [int(n) for n in len_queries]
        Lq = max(lq)
        if min(lq) < 1 or Lq > self.cfg.query_length:
            raise ValueError(f"query lengths must be in [1, {self.cfg.query_length}]")
        handle = self._ensure_handle()
        dev = self.query_tokens.device
        if enc.device != dev:
            raise RuntimeError("encoder_hidden_states must live on the module's device")
        if enc.dtype not in _DTYPES:
            enc = enc.float()
        enc = enc.contiguous()
        pad = None
        if encoder_attention_mask is not None:
            if tuple(encoder_attention_mask.shape) != (B, Lk):
                raise ValueError(f"encoder_attention_mask must be [{B},{Lk}]")
            pad = (encoder_attention_mask == 0).to(device=dev).contiguous().view(torch.uint8)
        out_dtype = self.query_tokens.dtype if self.query_tokens.dtype in _DTYPES else torch.float32
        out = torch.empty(B, Lq, self.cfg.hidden_size, device=dev, dtype=out_dtype)
        vp = ctypes.c_void_p
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream(dev).cuda_stream
            _lib.check(_lib.load().avh_qformer_forward(
                handle, vp(enc.data_ptr()), _DTYPES[enc.dtype], vp(pad.data_ptr()) if pad is not None else None,
                (ctypes.c_int32 * B)(*lq), B, Lq, Lk, vp(out.data_ptr()), _DTYPES[out_dtype], vp(stream)))
        return out

    @torch.no_grad()
    def compression_using_qformer(self, len_queries, resized_len_list, len_feat, av_feat):
        """src/model.py:584-619: per-clip linear resize of av_feat [B,T,C] (valid rows len_feat[b]) to
        int(resized_len_list[b]) rows, zero-padded batch + mask, then the Q-Former: returns query_output
        [B, max(len_queries), hidden]."""
        resized, mask = resize_av_features(av_feat, len_feat, resized_len_list)
        return self.bert(len_queries, resized, mask)
