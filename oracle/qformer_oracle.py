"""TEST INFRASTRUCTURE ONLY — plain-PyTorch fp32 restatement of the Q-Former stage of MMS-LLaMA on the query-only path:
``compression_using_qformer`` (src/model.py:584-619) = per-clip ``F.interpolate`` into a zero-padded batch, then
``Qformer.bert(query_embeds, attention_mask, encoder_hidden_states, encoder_attention_mask)``
(src/sub_model/Qformer.py:805-968: embeddings LayerNorm; per layer self-attention, cross-attention with K / V from the
AV features, intermediate_query / output_query; post-LN, eps 1e-12, erf GELU, additive -10000 key masks).

Parameter names are the reference's state-dict keys, so the REAL class's weights load with strict=False.  Pinned by
tests/golden/qformer_*.npz = outputs of the real classes (oracle/make_golden_qformer.py).  Never imported by the product.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Self(nn.Module):
    def __init__(self, dim, kv_dim, heads):
        super().__init__()
        self.query, self.key, self.value = nn.Linear(dim, dim), nn.Linear(kv_dim, dim), nn.Linear(kv_dim, dim)
        self.heads = heads

    def forward(self, x, kv, add_mask):
        B, Lq, D = x.shape
        hd = D // self.heads

        def split(t):
            return t.view(B, -1, self.heads, hd).permute(0, 2, 1, 3)
        q, k, v = split(self.query(x)), split(self.key(kv)), split(self.value(kv))
        s = q @ k.transpose(-1, -2) / math.sqrt(hd) + add_mask
        return (s.softmax(-1) @ v).permute(0, 2, 1, 3).reshape(B, Lq, D)


class _Out(nn.Module):
    def __init__(self, din, dim, eps):
        super().__init__()
        self.dense, self.LayerNorm = nn.Linear(din, dim), nn.LayerNorm(dim, eps=eps)

    def forward(self, h, residual):
        return self.LayerNorm(self.dense(h) + residual)


class _Attn(nn.Module):
    def __init__(self, dim, kv_dim, heads, eps):
        super().__init__()
        self.self, self.output = _Self(dim, kv_dim, heads), _Out(dim, dim, eps)

    def forward(self, x, kv, add_mask):
        return self.output(self.self(x, kv, add_mask), x)


class _Dense(nn.Module):
    def __init__(self, din, dout):
        super().__init__()
        self.dense = nn.Linear(din, dout)


class _Layer(nn.Module):
    def __init__(self, dim, inter, enc_width, heads, eps):
        super().__init__()
        self.attention = _Attn(dim, dim, heads, eps)
        self.crossattention = _Attn(dim, enc_width, heads, eps)
        self.intermediate_query, self.output_query = _Dense(dim, inter), _Out(inter, dim, eps)

    def forward(self, x, self_mask, enc, enc_mask):
        x = self.attention(x, x, self_mask)
        x = self.crossattention(x, enc, enc_mask)
        return self.output_query(F.gelu(self.intermediate_query.dense(x)), x)


class _Emb(nn.Module):
    def __init__(self, dim, eps):
        super().__init__()
        self.LayerNorm = nn.LayerNorm(dim, eps=eps)


class _Enc(nn.Module):
    def __init__(self, n, *a):
        super().__init__()
        self.layer = nn.ModuleList([_Layer(*a) for _ in range(n)])


class _Bert(nn.Module):
    def __init__(self, hidden, heads, inter, layers, enc_width, eps):
        super().__init__()
        self.embeddings = _Emb(hidden, eps)
        self.encoder = _Enc(layers, hidden, inter, enc_width, heads, eps)


class _Q(nn.Module):
    def __init__(self, *a):
        super().__init__()
        self.bert = _Bert(*a)


class OracleQFormer(nn.Module):
    def __init__(self, hidden, heads, inter, layers, enc_width, query_length, eps=1e-12):
        super().__init__()
        self.Qformer = _Q(hidden, heads, inter, layers, enc_width, eps)
        self.query_tokens = nn.Parameter(torch.zeros(1, query_length, hidden))

    @torch.no_grad()
    def bert(self, len_queries, enc, enc_mask):
        B, Lq = len(len_queries), max(len_queries)
        qmask = torch.zeros(B, Lq)
        for i, n in enumerate(len_queries):
            qmask[i, :n] = 1
        x = self.Qformer.bert.embeddings.LayerNorm(self.query_tokens.expand(B, -1, -1)[:, :Lq])
        self_add = (1.0 - qmask)[:, None, None, :] * -10000.0
        enc_add = (1.0 - enc_mask.float())[:, None, None, :] * -10000.0
        for layer in self.Qformer.bert.encoder.layer:
            x = layer(x, self_add, enc.float(), enc_add)
        return x

    @torch.no_grad()
    def compression_using_qformer(self, len_queries, resized_len_list, len_feat, av_feat):
        B = len(len_queries)
        Tm = int(max(resized_len_list))
        feats = torch.zeros(B, Tm, av_feat.size(2))
        mask = torch.zeros(B, Tm, dtype=torch.long)
        for b, n in enumerate(len_feat):
            r = F.interpolate(av_feat[b][:n].float().t()[None], size=int(resized_len_list[b]), mode="linear")[0].t()
            feats[b, :r.size(0)] = r
            mask[b, :int(resized_len_list[b])] = 1
        return self.bert(len_queries, feats, mask)


def synthetic_case(seed=3, B=3, T=40, C=192):
    g = torch.Generator().manual_seed(seed)
    av = torch.randn(B, T, C, generator=g)
    len_feat = [T, T - 11, max(T // 3, 1)][:B] + [T] * max(B - 3, 0)
    rates = [1.0, 1.62, 2.0][:B] + [1.3] * max(B - 3, 0)
    len_queries = [max(int(n / 25 * 3 * r), 3) for n, r in zip(len_feat, rates)]
    resized = [r * n for n, r in zip(len_feat, rates)]
    return av, len_feat, resized, len_queries
