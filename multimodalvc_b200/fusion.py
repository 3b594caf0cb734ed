"""The first consumers of the AV-HuBERT output inside MMS-LLaMA (src/model.py:296-330,561-609), as
device ops of ``libavh_b200.so`` — everything between the two encoders and the Q-Former's cross-attention:

* ``AudioFeatureConv`` = ``afeat_1d_conv``: ``nn.Conv1d(D, D, kernel_size=k, stride=k)`` over the Whisper features
  (k = 2 with the Q-Former, 4 without: src/model.py:115,152).  Non-overlapping windows make it ONE GEMM on the
  tcgen05 kernel: [B*T/k, k*D] x W'^T with W'[o, j*D + i] = weight[o, i, j];
* ``fuse_av`` = slice to the video length + ``concat`` / ``add`` (src/model.py:321-327);
* ``query_lengths`` = ``query_length_calculation``'s host arithmetic (src/model.py:563-581) on the predicted rates;
* ``resize_av_features`` = the per-sample ``F.interpolate(..., mode='linear')`` loop that builds the zero-padded
  ``resized_av_feats`` / ``resized_padding_masks`` (src/model.py:596-609) as one launch (``avh_interp_linear``).

The Q-Former itself (src/sub_model/Qformer.py) is not built."""
import ctypes
from typing import List, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib
from .hubert import _DTYPES
from .hubert_asr import _project


class AudioFeatureConv(nn.Module):
    """Parameter container with the reference's attribute layout (``weight [D, D, k]``, ``bias [D]`` — the state-dict
    keys of ``afeat_1d_conv``); forward on the library's GEMM."""

    def __init__(self, dim, kernel_size=2):
        super().__init__()
        conv = nn.Conv1d(dim, dim, kernel_size=kernel_size, stride=kernel_size, padding=0)
        self.weight, self.bias = conv.weight, conv.bias
        self.kernel_size, self.dim = kernel_size, dim
        self._lin = None

    def _as_linear(self):
        key = (self.weight._version, self.bias._version, self.weight.data_ptr())
        if self._lin is None or self._lin[0] != key:
            lin = nn.Linear(self.kernel_size * self.dim, self.dim, device=self.weight.device, dtype=self.weight.dtype)
            with torch.no_grad():
                lin.weight.copy_(self.weight.permute(0, 2, 1).reshape(self.dim, self.kernel_size * self.dim))
                lin.bias.copy_(self.bias)
            self._lin = (key, lin)
        return self._lin[1]

    @torch.no_grad()
    def forward(self, x):
        """x [B, T, D] -> [B, T // k, D]  (= afeat_1d_conv(x.transpose(1, 2)).transpose(1, 2), src/model.py:304)."""
        B, T, D = x.shape
        k = self.kernel_size
        To = T // k
        if To < 1:
            raise ValueError("sequence shorter than the convolution kernel")
        return _project(x[:, :To * k].reshape(B, To, k * D), self._as_linear())


def fuse_av(whisper_feat, av_out, mode="concat"):
    """src/model.py:318-327: whisper features cut to the video length, then concat (feature dim) or add."""
    T_v = av_out.size(1)
    w = whisper_feat[:, :T_v, :]
    if mode == "concat":
        return torch.cat([w, av_out], dim=2)
    if mode == "add":
        return w + av_out
    raise ValueError(f"unknown modality fusion type {mode}")


def query_lengths(sr_predictions: Sequence[float], video_lengths: Sequence[int], queries_per_sec: int) -> Tuple[List[int], List[float]]:
    """src/model.py:566-581: rate clipped to [1, 2]; queries = max(int(len / 25 * qps * rate), qps); resized length = rate * len."""
    len_queries, resized = [], []
    for rate, vid_len in zip(sr_predictions, video_lengths):
        factor = float(rate)
        if factor < 1:
            factor = 1
        elif factor > 2:
            factor = 2
        len_queries.append(max(int(vid_len / 25 * queries_per_sec * factor), queries_per_sec))
        resized.append(factor * vid_len)
    return len_queries, resized


@torch.no_grad()
def resize_av_features(av_feat, len_feat: Sequence[int], resized_len_list: Sequence[float]):
    """src/model.py:596-609.  av_feat [B,T,C] (CUDA), len_feat valid frames per sample, resized_len_list target
    lengths (int() of each, as the reference).  Returns (resized [B, int(max(resized)), C], padding_mask int64)."""
    if not av_feat.is_cuda:
        raise RuntimeError("resize_av_features computes on a B200 only (there is no CPU path)")
    B, T, C = av_feat.shape
    if len(len_feat) != B or len(resized_len_list) != B:
        raise ValueError("one length per sample expected")
    if av_feat.dtype not in _DTYPES:
        av_feat = av_feat.float()
    av_feat = av_feat.contiguous()
    dev = av_feat.device
    n_in = [int(n) for n in len_feat]
    n_out = [int(n) for n in resized_len_list]
    if min(n_in) < 1 or max(n_in) > T or min(n_out) < 1:
        raise ValueError("lengths out of range")
    Tout = int(max(resized_len_list))
    li = torch.tensor(n_in, dtype=torch.int32).to(dev)
    lo = torch.tensor(n_out, dtype=torch.int32).to(dev)
    out = torch.empty(B, Tout, C, device=dev, dtype=av_feat.dtype)
    mask = torch.empty(B, Tout, device=dev, dtype=torch.int64)
    vp = ctypes.c_void_p
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        _lib.check(_lib.load().avh_interp_linear(vp(av_feat.data_ptr()), _DTYPES[av_feat.dtype], B, T, C, vp(li.data_ptr()),
                                                 vp(lo.data_ptr()), Tout, vp(out.data_ptr()), vp(mask.data_ptr()), vp(stream)))
    return out, mask
