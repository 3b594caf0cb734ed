"""GPU: raw uint8 video path — the dataset's eval transform on the device and uint8 frames accepted
directly by extract_finetune / extract_finetune_host.  fp32 results are bit-exact against the float64 numpy
restatement cast to float32 (what the reference hands to the model)."""
import os

import numpy as np
import pytest
import torch

from multimodalvc_b200 import video
from oracle import video_oracle as vo

from helpers import load_encoder_case, make_device_model, rel_err

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("T,H,W", [(7, 96, 96), (3, 97, 101), (2, 88, 88), (1, 120, 90), (300, 96, 96)])
def test_normalize_crop_bit_exact_fp32_and_rounded_bf16(T, H, W):
    frames = vo.synthetic_frames(T, H, W, seed=T)
    ref = vo.video_transform(frames).astype(np.float32)
    dev = torch.from_numpy(frames).cuda()
    out = video.normalize_crop(dev)
    assert out.shape == (T, 88, 88) and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), ref)
    out_bf = video.normalize_crop(dev, dtype=torch.bfloat16)
    assert torch.equal(out_bf.cpu(), torch.from_numpy(ref).to(torch.bfloat16))
    out_h = video.normalize_crop(dev, dtype=torch.float16)
    assert torch.equal(out_h.cpu(), torch.from_numpy(ref).to(torch.float16))
    assert torch.equal(torch.from_numpy(frames), dev.cpu())                   # input untouched


def test_normalize_crop_against_reference_golden_and_collate():
    z = np.load(os.path.join(GOLDEN, "video_reference.npz"))
    for name in ("roi96", "odd_97x101", "exact88"):
        out = video.normalize_crop(torch.from_numpy(z["frames_" + name]).cuda())
        assert np.array_equal(out.cpu().numpy(), z["out_" + name].astype(np.float32)), name
    lens = z["coll_lens"]
    clips, o = [], 0
    for n in lens:
        clips.append(torch.from_numpy(z["coll_frames"][o:o + n]).cuda())
        o += n
    v, pm = video.collate_video(clips)
    assert np.array_equal(v.cpu().numpy(), z["coll_out"]) and np.array_equal(pm.cpu().numpy(), z["coll_mask"])


def test_leading_dims_and_empty():
    frames = vo.synthetic_frames(6, 96, 96, seed=3).reshape(2, 1, 3, 96, 96)
    out = video.normalize_crop(torch.from_numpy(frames).cuda())
    assert out.shape == (2, 1, 3, 88, 88)
    assert np.array_equal(out.cpu().numpy().reshape(6, 88, 88), vo.video_transform(frames.reshape(6, 96, 96)).astype(np.float32))
    assert video.normalize_crop(torch.zeros(0, 96, 96, dtype=torch.uint8, device="cuda")).shape == (0, 88, 88)
    with pytest.raises(ValueError):
        video.normalize_crop(torch.zeros(2, 96, 96, dtype=torch.uint8))       # CPU tensor: no CPU path


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_extract_finetune_accepts_raw_uint8_video(dtype):
    c = load_encoder_case("tiny_av_ragged")
    B, T = c["pm"].shape
    m = make_device_model(c["oracle"], c["over"], c["size"], dtype)
    frames = vo.synthetic_frames(B * T, 96, 96, seed=11).reshape(B, 1, T, 96, 96)
    raw = torch.from_numpy(frames).cuda()
    norm = torch.from_numpy(vo.video_transform(frames.reshape(B * T, 96, 96)).astype(np.float32)).view(B, 1, T, 88, 88)
    norm = norm.masked_fill(c["pm"].view(B, 1, T, 1, 1), 0.0)      # the collater zero-pads AFTER Normalize
    audio = c["src"]["audio"].cuda().to(dtype)
    pm = c["pm"].cuda()
    y_raw, _ = m.extract_finetune({"audio": audio, "video": raw}, pm)
    y_norm, _ = m.extract_finetune({"audio": audio, "video": norm.cuda().to(dtype)}, pm)
    assert torch.equal(y_raw, y_norm)              # same normalised pixels reach the stem: identical output
    with torch.no_grad():
        y_ref, _ = c["oracle"].extract_finetune({"audio": c["src"]["audio"], "video": norm}, c["pm"])
    if dtype == torch.float32:
        assert rel_err(y_raw.cpu(), y_ref) < 2e-3
    # host-buffer entry point with pinned uint8 frames (1 byte per pixel over PCIe)
    y_host = m.extract_finetune_host(torch.from_numpy(frames).pin_memory(), audio.cpu().contiguous(), c["pm"])
    assert torch.equal(y_host, y_raw.cpu())
    with pytest.raises(ValueError):
        m.extract_finetune({"audio": audio, "video": raw[:, :, :, :80, :80]}, pm)   # smaller than the crop
