"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain torch) of the ops between the encoders and the Q-Former in
MMS-LLaMA: src/model.py:304 (afeat_1d_conv), :318-327 (slice + concat / add), :563-581 (query_length_calculation),
:596-609 (per-sample F.interpolate into zero-padded tensors + masks).  The arithmetic is torch's own (nn.Conv1d,
F.interpolate), called exactly as the reference calls it, so there is nothing to pin beyond the call pattern."""
import torch
import torch.nn.functional as F


def afeat_conv(conv, x):                       # src/model.py:304
    return conv(x.transpose(1, 2)).transpose(1, 2)


def fuse(whisper_feat, av_out, mode):          # src/model.py:318-327
    w = whisper_feat[:, :av_out.size(1), :]
    return torch.cat([w, av_out], dim=2) if mode == "concat" else w + av_out


def query_lengths(sr_predictions, video_lengths, queries_per_sec):      # src/model.py:566-581
    len_queries, resized = [], []
    for i, vid_len in enumerate(video_lengths):
        base_queries = vid_len / 25 * queries_per_sec
        factor = sr_predictions[i]
        if factor < 1:
            factor = 1
        elif factor > 2:
            factor = 2
        len_queries.append(max(int(base_queries * factor), queries_per_sec))
        resized.append(factor * vid_len)
    return len_queries, resized


def resize(av_feat, len_feat, resized_len_list):                          # src/model.py:596-609
    B = av_feat.size(0)
    out = torch.zeros(B, int(max(resized_len_list)), av_feat.size(2), dtype=av_feat.dtype)
    mask = torch.zeros(B, int(max(resized_len_list)), dtype=av_feat.dtype)
    for bs, n in enumerate(len_feat):
        x = av_feat[bs][:n].transpose(0, 1).unsqueeze(0)
        r = F.interpolate(x, size=int(resized_len_list[bs]), mode="linear").squeeze(0).transpose(0, 1)
        out[bs, :r.size(0)] = r
        mask[bs, :int(resized_len_list[bs])] = 1
    return out, mask.long()
