"""TEST INFRASTRUCTURE ONLY — numpy (float64) oracle of the audio featurisation on the hot path.

The arithmetic lives in a THIRD-PARTY dependency that is not vendored under /root/reference, not
installed in the image and not fetchable (no network): `python_speech_features` (README.md:26, unpinned;
latest PyPI release 0.6).  Reference call sites: avhubert/hubert_dataset.py:20,286 and
avhubert/clustering/dump_hubert_feature.py:19,69 (`logfbank(wav_data, samplerate=16000)`, all other
arguments default).  `logfbank` below restates that library's published algorithm:

  logfbank -> fbank -> sigproc.preemphasis (coeff 0.97; y[0]=x[0])
                    -> sigproc.framesig   (frame_len=round_half_up(0.025*16000)=400, step 160,
                                           numframes = 1 if n<=400 else 1+ceil((n-400)/160), zero-pad tail,
                                           rectangular window)
                    -> sigproc.powspec    (|rfft(frame, 512)|^2 / 512)
                    -> get_filterbanks    (26 triangular filters, mel points linspace(hz2mel(0), hz2mel(8000), 28),
                                           bin = floor(513 * mel2hz(mel) / 16000))
                    -> dot, where(==0, eps), natural log

PARITY UNPINNED for logfbank: the reference holds no golden vectors for it and the library cannot be
run here.  The known-answer tests in tests/test_fbank_oracle.py (frame-count table, filterbank structure,
pure-tone response) are the only pins.  Everything else in this file restates code that IS under
/root/reference and is pinned against it (tests/test_oracle_vs_reference.py):

  stacker          avhubert/hubert_dataset.py:259-274
  align_to_video   avhubert/hubert_dataset.py:290-295
  frame_layer_norm avhubert/hubert_dataset.py:351-353 (F.layer_norm over the 104 features, eps 1e-5, no affine)
  add_noise        avhubert/hubert_dataset.py:317-346
  collater_audio   avhubert/hubert_dataset.py:430-456
"""
import math

import numpy as np

EPS = np.finfo(float).eps


def hz2mel(hz):
    return 2595.0 * np.log10(1.0 + hz / 700.0)


def mel2hz(mel):
    return 700.0 * (10.0 ** (mel / 2595.0) - 1.0)


def mel_bins(nfilt=26, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    highfreq = highfreq or samplerate / 2
    melpoints = np.linspace(hz2mel(lowfreq), hz2mel(highfreq), nfilt + 2)
    return np.floor((nfft + 1) * mel2hz(melpoints) / samplerate)


def get_filterbanks(nfilt=26, nfft=512, samplerate=16000, lowfreq=0, highfreq=None):
    b = mel_bins(nfilt, nfft, samplerate, lowfreq, highfreq)
    fb = np.zeros([nfilt, nfft // 2 + 1])
    for j in range(nfilt):
        for i in range(int(b[j]), int(b[j + 1])):
            fb[j, i] = (i - b[j]) / (b[j + 1] - b[j])
        for i in range(int(b[j + 1]), int(b[j + 2])):
            fb[j, i] = (b[j + 2] - i) / (b[j + 2] - b[j + 1])
    return fb


def num_frames(n_samples, frame_len=400, frame_step=160):
    if n_samples <= frame_len:
        return 1
    return 1 + int(math.ceil((1.0 * n_samples - frame_len) / frame_step))


def preemphasis(signal, coeff=0.97):
    return np.append(signal[0], signal[1:] - coeff * signal[:-1])


def framesig(sig, frame_len=400, frame_step=160):
    n = len(sig)
    nf = num_frames(n, frame_len, frame_step)
    padlen = (nf - 1) * frame_step + frame_len
    padded = np.concatenate((sig, np.zeros(padlen - n)))
    idx = np.arange(frame_len)[None, :] + (np.arange(nf) * frame_step)[:, None]
    return padded[idx]


def powspec(frames, nfft=512):
    return 1.0 / nfft * np.square(np.absolute(np.fft.rfft(frames, nfft)))


def logfbank(signal, samplerate=16000):
    assert samplerate == 16000
    signal = np.asarray(signal)
    pspec = powspec(framesig(preemphasis(signal)))
    feat = np.dot(pspec, get_filterbanks().T)
    feat = np.where(feat == 0, EPS, feat)
    return np.log(feat)


def stacker(feats, stack_order=4):
    feat_dim = feats.shape[1]
    if len(feats) % stack_order != 0:
        res = stack_order - len(feats) % stack_order
        feats = np.concatenate([feats, np.zeros([res, feat_dim]).astype(feats.dtype)], axis=0)
    return feats.reshape((-1, stack_order, feat_dim)).reshape(-1, stack_order * feat_dim)


def align_to_video(audio_feats, n_video):
    diff = len(audio_feats) - n_video
    if diff < 0:
        audio_feats = np.concatenate(
            [audio_feats, np.zeros([-diff, audio_feats.shape[-1]], dtype=audio_feats.dtype)])
    elif diff > 0:
        audio_feats = audio_feats[:-diff]
    return audio_feats


def frame_layer_norm(feats, eps=1e-5):
    x = feats.astype(np.float64)
    mu = x.mean(axis=1, keepdims=True)
    var = x.var(axis=1, keepdims=True)
    return ((x - mu) / np.sqrt(var + eps)).astype(np.float32)


def featurize_clip(wav, n_video=None, normalize=True, stack_order=4):
    """wav int16 [n] -> float32 [T,104]: logfbank -> float32 -> stack -> align -> per-frame LN."""
    f = logfbank(wav).astype(np.float32)
    f = stacker(f, stack_order)
    if n_video is not None:
        f = align_to_video(f, n_video)
    if normalize:
        f = frame_layer_norm(f)
    return f


def add_noise(clean_wav, noise_wav, snr):
    """hubert_dataset.py:317-346 with the noise already selected (float32 array)."""
    clean_wav = clean_wav.astype(np.float32)
    noise_wav = noise_wav.astype(np.float32)
    clean_rms = np.sqrt(np.mean(np.square(clean_wav), axis=-1))
    if len(clean_wav) > len(noise_wav):
        ratio = int(np.ceil(len(clean_wav) / len(noise_wav)))
        noise_wav = np.concatenate([noise_wav for _ in range(ratio)])
    if len(clean_wav) < len(noise_wav):
        noise_wav = noise_wav[0:len(clean_wav)]
    noise_rms = np.sqrt(np.mean(np.square(noise_wav), axis=-1))
    adjusted_noise_rms = clean_rms / (10 ** (snr / 20))
    mixed = clean_wav + noise_wav * (adjusted_noise_rms / noise_rms)
    max_i, min_i = np.iinfo(np.int16).max, np.iinfo(np.int16).min
    if mixed.max(axis=0) > max_i or mixed.min(axis=0) < min_i:
        if mixed.max(axis=0) >= abs(mixed.min(axis=0)):
            rate = max_i / mixed.max(axis=0)
        else:
            rate = min_i / mixed.min(axis=0)
        mixed = mixed * rate
    return mixed.astype(np.int16)


def collater_audio(items, size):
    """items: list of float32 arrays [T_i, ...]; zero-pad to `size`; mask True on padded steps.
    Returns (collated [B,size,...], padding_mask bool [B,size]); crop branch (T_i > size) keeps the head."""
    shape = list(items[0].shape[1:])
    out = np.zeros([len(items), size] + shape, dtype=items[0].dtype)
    mask = np.zeros([len(items), size], dtype=bool)
    for i, a in enumerate(items):
        n = min(len(a), size)
        out[i, :n] = a[:n]
        mask[i, n:] = True
    return out, mask


def synthetic_wave(n_samples, seed):
    """int16 from clip(round(3000*randn))."""
    r = np.random.RandomState(seed)
    return np.clip(np.round(3000.0 * r.randn(n_samples)), -32768, 32767).astype(np.int16)


def synthetic_babble(n_samples, seed, n_streams=30):
    """babble = floor(mean of independent streams) (avhubert/hubert_dataset.py:304-315 with noise_num>1)."""
    streams = [synthetic_wave(n_samples, seed * 1000 + i).astype(np.float32) for i in range(n_streams)]
    return np.floor(np.stack(streams).mean(axis=0)).astype(np.float32)
