"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) — numpy restatement of the video pre-processing of the
reference dataset's eval path.  Pinned: tests/golden/video_reference.npz holds outputs of the REAL
avhubert/utils.py transform classes and of the real collater (oracle/make_golden.py), compared bit for bit in
tests/test_video_oracle.py.

Follows avhubert/hubert_dataset.py:222-226 (transform = Compose([Normalize(0.0, 255.0), CenterCrop((crop, crop)),
Normalize(image_mean, image_std)])), :298-302 (load_video: transform + expand_dims(-1)), :350 (astype(float32)),
:430-456 (collater_audio for 5-D items: zero-padded tail, permute to [B,C,T,H,W]) and avhubert/utils.py:56-95.
"""
import numpy as np

IMAGE_MEAN, IMAGE_STD, IMAGE_CROP = 0.421, 0.165, 88       # avhubert/hubert_pretraining.py:144-149


def normalize(frames, mean, std):
    """utils.py:66-75: (frames - mean) / std — float64 for uint8 input and python-float constants."""
    return (frames - mean) / std


def center_crop(frames, size):
    """utils.py:84-90."""
    t, h, w = frames.shape
    th, tw = size
    delta_w = int(round((w - tw)) / 2.)
    delta_h = int(round((h - th)) / 2.)
    return frames[:, delta_h:delta_h + th, delta_w:delta_w + tw]


def video_transform(frames_u8, crop=IMAGE_CROP, mean=IMAGE_MEAN, std=IMAGE_STD):
    """uint8 [T,H,W] -> float64 [T,crop,crop] (hubert_dataset.py:222-226)."""
    x = normalize(frames_u8, 0.0, 255.0)
    x = center_crop(x, (crop, crop))
    return normalize(x, mean, std)


def load_video_feats(frames_u8, **kw):
    """hubert_dataset.py:298-302 + :350: [T,crop,crop,1] float32 as __getitem__ hands it to the collater."""
    return np.expand_dims(video_transform(frames_u8, **kw), axis=-1).astype(np.float32)


def collater_video(items, size):
    """hubert_dataset.py:430-456 for 5-D items ([T_i,H,W,1] float32): (video [B,1,T,H,W], padding_mask [B,T])."""
    B = len(items)
    shape = list(items[0].shape[1:])
    out = np.zeros([B, size] + shape, dtype=items[0].dtype)
    mask = np.zeros((B, size), dtype=bool)
    for i, v in enumerate(items):
        n = min(len(v), size)
        out[i, :n] = v[:n]
        mask[i, n:] = True
    return np.ascontiguousarray(out.transpose(0, 4, 1, 2, 3)), mask


def synthetic_frames(T, H=96, W=96, seed=0):
    """uint8 mouth-ROI-like frames: smooth gradient + noise, full 0..255 range exercised."""
    rs = np.random.RandomState(seed)
    base = np.linspace(0, 255, H * W).reshape(H, W)[None] * rs.uniform(0.5, 1.0, (T, 1, 1))
    x = base + rs.normal(0, 40, (T, H, W))
    x[0, 0, :256 if W >= 256 else W] = np.arange(min(W, 256)) * (255.0 / max(1, min(W, 256) - 1))
    return np.clip(np.round(x), 0, 255).astype(np.uint8)
