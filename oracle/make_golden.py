"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.npz by running the REAL, unmodified reference code
(/root/reference, through oracle/ref_import.py) in this container.  The GPU box has no /root/reference, so the
fixtures travel instead.  Run:  python -m oracle.make_golden

Encoder fixtures: weights are NOT stored (ResNet-18 alone is 45 MB); they are re-created deterministically by
oracle.avhubert_oracle.build_oracle(size, seed) — same torch build on both boxes — and copied into the real
reference AVHubertModel here; the fixture holds the inputs' seeds and the reference's outputs.  A checksum of
the state dict is stored so a silent RNG change shows up as a fixture mismatch rather than a parity failure.

Audio fixtures: stacker / alignment / add_noise / collater outputs of the real avhubert/hubert_dataset.py
(with python_speech_features.logfbank supplied by oracle.fbank_oracle.logfbank, the library being absent),
plus log-fbank known answers from the float64 restatement.
"""
import hashlib
import os
import sys

import numpy as np
import torch

from . import avhubert_oracle as ao
from . import fbank_oracle as fo
from . import ref_import

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

ENCODER_CASES = [
    # name, size, B, T, lengths, audio, video, output_layer, cfg overrides
    ("tiny_av_ragged", "tiny", 3, 20, [20, 13, 7], True, True, None, {}),
    ("tiny_video_only", "tiny", 2, 12, None, False, True, None, {}),
    ("tiny_audio_only", "tiny", 2, 16, [16, 9], True, False, None, {}),
    ("tiny_layer1", "tiny", 2, 12, [12, 8], True, True, 1, {}),
    ("tiny_postln", "tiny", 2, 12, [12, 10], True, True, None, {"layer_norm_first": False}),
    ("tiny_add", "tiny", 2, 12, [12, 10], True, True, None, {"modality_fuse": "add"}),
    ("base_b1_t50", "base", 1, 50, None, True, True, None, {}),          # BASELINE config 1
    ("large_b2_t40", "large", 2, 40, [40, 27], True, True, None, {}),
]


def state_checksum(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def make_encoder_case(name, size, B, T, lengths, audio, video, output_layer, over):
    oracle = ao.build_oracle(size, seed=1234, **over)
    ref, _ = ref_import.build_reference_model(size, **over)
    missing = ref.load_state_dict(oracle.state_dict(), strict=False)
    assert not missing.unexpected_keys, missing
    ref.eval()
    src, pm = ao.synthetic_inputs(B, T, lengths=lengths, seed=11, audio=audio, video=video)
    with torch.no_grad():
        y, pm_out = ref.extract_finetune(src, pm, output_layer=output_layer)
        y_o, _ = oracle.extract_finetune(src, pm, output_layer=output_layer)
    err = (y - y_o).abs().max().item()
    print(f"{name}: ref vs oracle max abs diff {err:.3e}, |y| max {y.abs().max().item():.3f}")
    assert err < 2e-4, "oracle restatement disagrees with the reference"
    np.savez_compressed(
        os.path.join(OUT, f"enc_{name}.npz"),
        y=y.numpy().astype(np.float32),
        pm_out=(pm_out.numpy() if pm_out is not None else np.zeros(0, dtype=bool)),
        meta=np.array([size, str(B), str(T), repr(lengths), str(int(audio)), str(int(video)), repr(output_layer),
                       repr(over), state_checksum(oracle.state_dict())]))


def make_audio_fixtures():
    ds = ref_import.install_dataset(fo.logfbank)
    rs = np.random.RandomState(5)
    # --- stacker via the real load_feature closure is not reachable without files; the real functions that are
    #     plain methods/closures are exercised through a minimal instance
    obj = ds.AVHubertDataset.__new__(ds.AVHubertDataset)
    obj.pad_audio, obj.random_crop, obj.max_sample_size = True, False, 500
    items = [torch.from_numpy(rs.randn(n, 104).astype(np.float32)) for n in (37, 50, 12)]
    coll, pmask, _ = obj.collater_audio([x.clone() for x in items], 50)
    vids = [torch.from_numpy(rs.randn(n, 4, 4, 1).astype(np.float32)) for n in (37, 50, 12)]
    collv, pmaskv, _ = obj.collater_audio([x.clone() for x in vids], 50)
    # add_noise with a fixed "selected" noise
    clean = fo.synthetic_wave(48000, 3)
    noise = fo.synthetic_babble(30000, 9)
    mixes = {}
    for snr in (-5, 0, 5, 40):
        obj.noise_snr = snr
        obj.select_noise = lambda noise=noise: noise.copy()
        mixes[snr] = obj.add_noise(clean.copy())
    loud = (clean.astype(np.float32) * 9).clip(-32768, 32767).astype(np.int16)   # forces the clipping branch
    obj.noise_snr = -5
    mix_loud = obj.add_noise(loud.copy())
    np.savez_compressed(
        os.path.join(OUT, "audio_reference.npz"),
        coll_items=np.concatenate([x.numpy() for x in items]), coll_lens=np.array([37, 50, 12]),
        coll_out=coll.numpy(), coll_mask=pmask.numpy(), collv_out=collv.numpy(), collv_mask=pmaskv.numpy(),
        clean=clean, noise=noise, loud=loud, mix_loud=mix_loud,
        **{f"mix_snr_{k}".replace("-", "m"): v for k, v in mixes.items()})
    # --- log-fbank known answers (float64 restatement; python_speech_features itself is absent => unpinned)
    waves = {"noise_6s": fo.synthetic_wave(96000, 1), "noise_ragged": fo.synthetic_wave(12345, 2),
             "short_300": fo.synthetic_wave(300, 4), "len_401": fo.synthetic_wave(401, 6)}
    t = np.arange(16000)
    waves["tone_1k"] = np.round(8000 * np.sin(2 * np.pi * 1000 * t / 16000)).astype(np.int16)
    waves["silence_tail"] = np.concatenate([fo.synthetic_wave(4000, 8), np.zeros(4000, dtype=np.int16)])
    out = {}
    for k, w in waves.items():
        out["wav_" + k] = w
        out["fbank_" + k] = fo.logfbank(w)                       # float64 [nframes, 26]
        out["feat_" + k] = fo.featurize_clip(w, normalize=True)  # float32 [T, 104]
    out["filterbank"] = fo.get_filterbanks()
    np.savez_compressed(os.path.join(OUT, "audio_fbank.npz"), **out)
    print("audio fixtures written")


def make_extract_features_case():
    """Outputs of the REAL AVHubertModel.extract_features (hubert.py:676-692) on the tiny ragged case:
    ret_conv features, layer-1 features, full output."""
    oracle = ao.build_oracle("tiny", seed=1234)
    ref, _ = ref_import.build_reference_model("tiny")
    ref.load_state_dict(oracle.state_dict(), strict=False)
    ref.eval()
    src, pm = ao.synthetic_inputs(3, 20, lengths=[20, 13, 7], seed=11)
    out = {}
    with torch.no_grad():
        for key, kw in {"conv": dict(ret_conv=True), "layer1": dict(output_layer=1), "full": dict()}.items():
            y, pm_out = ref.extract_features({k: v.clone() for k, v in src.items()}, pm.clone(), mask=False, **kw)
            y_o, _ = oracle.extract_features(src, pm, **kw)
            err = (y - y_o).abs().max().item()
            print(f"extract_features[{key}]: ref vs oracle max abs diff {err:.3e}")
            assert err < 2e-4
            out[key] = y.numpy().astype(np.float32)
        out["pm_out"] = pm_out.numpy()
    np.savez_compressed(os.path.join(OUT, "enc_extract_features.npz"), **out)


def make_video_fixtures():
    """Outputs of the REAL transform classes (avhubert/utils.py) and of the real collater for raw uint8 frames."""
    ds = ref_import.install_dataset(fo.logfbank)
    # the real avhubert/utils.py, loaded by path (ref_import stubs the module name for the model import)
    import importlib.util
    spec = importlib.util.spec_from_file_location("avhubert_utils_real", os.path.join(ref_import.REF, "avhubert", "utils.py"))
    cu = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cu)
    from oracle import video_oracle as vo
    out = {}
    for name, (T, H, W) in {"roi96": (7, 96, 96), "odd_97x101": (3, 97, 101), "exact88": (2, 88, 88)}.items():
        frames = vo.synthetic_frames(T, H, W, seed=T + H)
        tf = cu.Compose([cu.Normalize(0.0, 255.0), cu.CenterCrop((88, 88)), cu.Normalize(0.421, 0.165)])
        y = tf(frames)                                       # float64 [T,88,88]
        out["frames_" + name] = frames
        out["out_" + name] = y
        assert np.array_equal(y, vo.video_transform(frames)), "video restatement disagrees with the reference"
    obj = ds.AVHubertDataset.__new__(ds.AVHubertDataset)
    obj.pad_audio, obj.random_crop, obj.max_sample_size = True, False, 500
    clips = [vo.synthetic_frames(n, 96, 96, seed=50 + n) for n in (5, 9, 2)]
    items = [torch.from_numpy(np.expand_dims(tf(c), -1).astype(np.float32)) for c in clips]
    coll, pmask, _ = obj.collater_audio(items, 9)
    out["coll_lens"] = np.array([5, 9, 2])
    out["coll_frames"] = np.concatenate(clips)
    out["coll_out"] = coll.numpy()
    out["coll_mask"] = pmask.numpy()
    np.savez_compressed(os.path.join(OUT, "video_reference.npz"), **out)
    print("video fixtures written")


def main():
    os.makedirs(OUT, exist_ok=True)
    if not ref_import.available():
        sys.exit("reference tree not available; fixtures can only be generated where /root/reference exists")
    torch.set_num_threads(8)
    make_audio_fixtures()
    make_video_fixtures()
    make_extract_features_case()
    if "--video-only" in sys.argv:
        return
    for case in ENCODER_CASES:
        make_encoder_case(*case)


if __name__ == "__main__":
    main()
