"""Device-side audio featurisation: the reference's per-sample numpy work in DataLoader workers
(avhubert/hubert_dataset.py:253-296,317-346,351-353,376-456) as batched sm_100a kernels.

``logfbank_stack_collate`` = logfbank(wav, 16000) -> float32 -> stacker(.,4) -> length alignment to the video ->
per-frame F.layer_norm -> collater_audio (zero pad to the batch size, padding_mask, [B,104,T] transposed view).
``add_noise`` = AVHubertDataset.add_noise with the noise clip already chosen.
"""
import ctypes
import math
from typing import List, Optional, Sequence

import torch

from . import _lib

FRAME_LEN, FRAME_STEP, STACK, NFILT = 400, 160, 4, 26


def num_frames(n_samples: int) -> int:
    """python_speech_features.sigproc.framesig frame count (25 ms / 10 ms at 16 kHz)."""
    if n_samples <= FRAME_LEN:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - FRAME_LEN) / FRAME_STEP))


def stacked_len(n_samples: int) -> int:
    """rows after stacker(., 4) (avhubert/hubert_dataset.py:259-274)."""
    return (num_frames(n_samples) + STACK - 1) // STACK


def _pack(wavs: Sequence[torch.Tensor], device):
    lens = [int(w.numel()) for w in wavs]
    offsets = torch.zeros(len(wavs) + 1, dtype=torch.int64)
    offsets[1:] = torch.tensor(lens, dtype=torch.int64).cumsum(0)
    flat = torch.cat([w.reshape(-1).to(torch.int16) for w in wavs]) if wavs else torch.zeros(0, dtype=torch.int16)
    return flat.to(device, non_blocking=True), offsets.to(device, non_blocking=True), lens


def logfbank_stack_collate(wavs: Sequence[torch.Tensor], video_lens: Optional[Sequence[int]] = None,
                           max_sample_size: Optional[int] = None, normalize: bool = True, device=None,
                           pad_audio: bool = True):
    """wavs: list of int16 1-D tensors (16 kHz).  Returns (audio [B,104,T] float32 — a transposed view exactly
    like collater_audio's, hubert_dataset.py:452-453 —, padding_mask bool [B,T]).

    T = min(max(len_i), max_sample_size) with len_i = video_lens[i] if given else the clip's own stacked length
    (collater, hubert_dataset.py:386-393, pad_audio=True; longer clips keep their head, random_crop=False)."""
    if not pad_audio:
        raise NotImplementedError("pad_audio=False (crop to the shortest clip) is not used on the inference path")
    device = torch.device(device if device is not None else "cuda")
    B = len(wavs)
    lens = [int(w.numel()) for w in wavs]
    own = [int(v) for v in video_lens] if video_lens is not None else [stacked_len(n) for n in lens]
    T = max(own) if own else 0
    if max_sample_size is not None:
        T = min(T, int(max_sample_size))
    out = torch.empty(B, T, STACK * NFILT, device=device, dtype=torch.float32)
    pm = torch.empty(B, T, device=device, dtype=torch.uint8)
    if B == 0 or T == 0:
        return out.transpose(1, 2), pm.bool()
    flat, offsets, _ = _pack(wavs, device)
    vl = torch.tensor(own, dtype=torch.int32).to(device) if video_lens is not None else None
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().avh_fbank(
            ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(offsets.data_ptr()),
            ctypes.c_void_p(vl.data_ptr()) if vl is not None else None, B, T, int(bool(normalize)),
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(pm.data_ptr()), ctypes.c_void_p(stream)))
    return out.transpose(1, 2), pm.view(torch.bool)


def logfbank_stack_collate_packed(flat: torch.Tensor, offsets: torch.Tensor, T: int,
                                  video_lens: Optional[torch.Tensor] = None, normalize: bool = True):
    """Same computation for clips that are already packed on the device (no host work, nothing allocated but the
    outputs): ``flat`` int16 CUDA tensor holding the clips back to back, ``offsets`` int64 CUDA tensor [B+1],
    ``video_lens`` optional int32 CUDA tensor [B].  Returns (audio [B,104,T] float32 view, padding_mask [B,T])."""
    if flat.dtype != torch.int16 or not flat.is_cuda or offsets.dtype != torch.int64 or not offsets.is_cuda:
        raise ValueError("flat must be an int16 CUDA tensor and offsets an int64 CUDA tensor")
    B = int(offsets.numel()) - 1
    device = flat.device
    out = torch.empty(B, T, STACK * NFILT, device=device, dtype=torch.float32)
    pm = torch.empty(B, T, device=device, dtype=torch.uint8)
    if B <= 0 or T <= 0:
        return out.transpose(1, 2), pm.bool()
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().avh_fbank(
            ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(offsets.data_ptr()),
            ctypes.c_void_p(video_lens.data_ptr()) if video_lens is not None else None, B, int(T), int(bool(normalize)),
            ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(pm.data_ptr()), ctypes.c_void_p(stream)))
    return out.transpose(1, 2), pm.view(torch.bool)


def add_noise_packed(flat: torch.Tensor, offsets: torch.Tensor, noise: torch.Tensor, snr_db: float) -> torch.Tensor:
    """add_noise for clips already packed on the device: ``flat`` int16 CUDA tensor (clips back to back), ``offsets``
    int64 CUDA tensor [B+1], ``noise`` float32 CUDA tensor.  Returns the mixed int16 clips in the same packing."""
    if flat.dtype != torch.int16 or not flat.is_cuda or offsets.dtype != torch.int64 or not offsets.is_cuda:
        raise ValueError("flat must be an int16 CUDA tensor and offsets an int64 CUDA tensor")
    device = flat.device
    B = int(offsets.numel()) - 1
    out = torch.empty_like(flat)
    if B <= 0:
        return out
    nz = noise.to(device=device, dtype=torch.float32).contiguous()
    scratch = torch.empty(4 * B, device=device, dtype=torch.float64)
    with torch.cuda.device(device):
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(_lib.load().avh_add_noise(
            ctypes.c_void_p(flat.data_ptr()), ctypes.c_void_p(offsets.data_ptr()), B,
            ctypes.c_void_p(nz.data_ptr()), int(nz.numel()), float(snr_db), ctypes.c_void_p(out.data_ptr()),
            ctypes.c_void_p(scratch.data_ptr()), ctypes.c_void_p(stream)))
    return out


def add_noise(wavs: Sequence[torch.Tensor], noise: torch.Tensor, snr_db: float, device=None) -> List[torch.Tensor]:
    """avhubert/hubert_dataset.py:317-346 for a batch of clips sharing one noise clip and SNR.
    wavs: int16 1-D tensors; noise: float32 1-D tensor (tiled when shorter, cropped from 0 when longer).
    Returns int16 tensors on `device`."""
    device = torch.device(device if device is not None else "cuda")
    if len(wavs) == 0:
        return []
    flat, offsets, lens = _pack(wavs, device)
    out = add_noise_packed(flat, offsets, noise.to(device), snr_db)
    res, o = [], 0
    for n in lens:
        res.append(out[o:o + n])
        o += n
    return res


def collate_video(videos: Sequence[torch.Tensor], T: int):
    """collater_audio for the video stream (hubert_dataset.py:430-456): list of [T_i,88,88,1] (or [T_i,88,88])
    -> ([B,1,T,88,88] contiguous, padding_mask bool [B,T]); pad frames are zeros, longer clips keep their head."""
    B = len(videos)
    dev = videos[0].device
    out = torch.zeros(B, 1, T, 88, 88, dtype=videos[0].dtype, device=dev)
    pm = torch.zeros(B, T, dtype=torch.bool, device=dev)
    for i, v in enumerate(videos):
        v = v.reshape(v.shape[0], 88, 88)
        n = min(v.shape[0], T)
        out[i, 0, :n] = v[:n]
        pm[i, n:] = True
    return out, pm
