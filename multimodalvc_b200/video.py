"""Video pre-processing of the dataset on the device: the eval transform of ``AVHubertDataset.load_video``
(avhubert/hubert_dataset.py:222-226,298-302 with avhubert/utils.py:56-95): ``Normalize(0, 255)`` ->
``CenterCrop(88)`` -> ``Normalize(image_mean, image_std)`` on raw uint8 gray mouth-ROI frames, and the collater's
``[B,T,H,W,1] -> [B,1,T,H,W]`` layout (hubert_dataset.py:455).  Raw frames are 4x smaller than the fp32 tensor the
reference ships host->device; ``AVHubertModel.extract_finetune`` also accepts them directly (uint8 video)."""
import ctypes

import torch

from . import _lib

IMAGE_MEAN, IMAGE_STD, IMAGE_CROP = 0.421, 0.165, 88      # hubert_pretraining.py:144-149

_DT = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def center_crop_offsets(h, w, crop=IMAGE_CROP):
    """CenterCrop.__call__ (avhubert/utils.py:84-90): delta = int(round(w - tw) / 2.)."""
    return int(round(h - crop) / 2.0), int(round(w - crop) / 2.0)


def normalize_crop(frames, crop=IMAGE_CROP, mean=IMAGE_MEAN, std=IMAGE_STD, dtype=torch.float32):
    """frames: uint8 CUDA tensor [..., H, W] (any leading dims, contiguous) -> [..., crop, crop] of ``dtype``.
    float32 results are bit-identical to the reference's float64 numpy transform cast to float32."""
    if frames.dtype != torch.uint8 or not frames.is_cuda:
        raise ValueError("normalize_crop takes a uint8 CUDA tensor (there is no CPU path)")
    if frames.dim() < 2:
        raise ValueError("frames must be [..., H, W]")
    frames = frames.contiguous()
    H, W = frames.shape[-2:]
    n = frames.numel() // (H * W) if H * W else 0
    out = torch.empty(*frames.shape[:-2], crop, crop, device=frames.device, dtype=dtype)
    if n == 0:
        return out
    with torch.cuda.device(frames.device):
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        _lib.check(_lib.load().avh_video_preprocess(
            ctypes.c_void_p(frames.data_ptr()), n, H, W, crop, float(mean), float(std),
            ctypes.c_void_p(out.data_ptr()), _DT[dtype], ctypes.c_void_p(stream)))
    return out


def collate_video(clips, crop=IMAGE_CROP, mean=IMAGE_MEAN, std=IMAGE_STD, dtype=torch.float32, max_frames=None):
    """List of uint8 CUDA clips [T_i, H, W] -> (video [B,1,T,crop,crop], padding_mask [B,T]) as the collater builds
    them (zero-padded tail frames, True = padded; hubert_dataset.py:430-456)."""
    T = max(int(c.size(0)) for c in clips)
    if max_frames is not None:
        T = min(T, int(max_frames))
    dev = clips[0].device
    video = torch.zeros(len(clips), 1, T, crop, crop, device=dev, dtype=dtype)
    pm = torch.zeros(len(clips), T, dtype=torch.bool, device=dev)
    for i, c in enumerate(clips):
        n = min(int(c.size(0)), T)
        video[i, 0, :n] = normalize_crop(c[:n], crop, mean, std, dtype)
        pm[i, n:] = True
    return video, pm
