"""TEST INFRASTRUCTURE: CPU restatement of the reference's batch assembly for the device kernel sb_prepare_batch_fwd.

Follows src/datasets/general_multisrc_dataset_dis_embed.py:112-218 (mixture / target / one-hot) and
src/datasets/perturbations/{SampleShift,ChannelGain,ChannelDrop,PeakNorm}Perturbation.py with the random draws passed
in (the reference draws them inside the classes).  Only tests may import this module."""
import torch


def pcm(x: torch.Tensor) -> torch.Tensor:
    return x.to(torch.float32) / 32768.0                     # torchaudio / librosa int16 normalisation (src/utils.py)


def assemble(mix, voices, inside, radius_idx=None, gain=None, shift=None, drop=None, peak_scale=None):
    B, M, N = mix.shape
    mixture = pcm(mix).clone()
    target = torch.zeros(B, 1, N)
    for b in range(B):
        if voices is not None:
            for v in range(voices.shape[1]):                  # :141-171: voices with dis <= radius, reference microphone
                if inside[b, v]:
                    target[b, 0] += pcm(voices[b, v])
        if shift is not None:                                 # SampleShiftPerturbation.py:24-31 (reference channel 0)
            for m in range(M):
                mixture[b, m] = torch.roll(mixture[b, m], int(shift[b, m]), dims=-1)
            target[b, 0] = torch.roll(target[b, 0], int(shift[b, 0]), dims=-1)
        if gain is not None:                                  # ChannelGainPerturbation.py:25-33
            for m in range(M):
                mixture[b, m] = mixture[b, m] * gain[b, m]
            target[b, 0] = target[b, 0] * gain[b, 0]
        if drop is not None:                                  # ChannelDropPerturbation.py:15-17
            for m in range(M):
                if drop[b, m]:
                    mixture[b, m] *= 0
        if peak_scale is not None and float(peak_scale[b]) != 0.0:        # PeakNormPerturbation.py:9-16
            peak = torch.abs(mixture[b]).max()
            s = float(peak_scale[b]) / (peak + 1e-6)
            mixture[b] = mixture[b] * s
            target[b] = target[b] * s
    idx = radius_idx if radius_idx is not None else torch.zeros(B, dtype=torch.int32)
    table = torch.tensor([[0., 0., 1.], [0., 1., 0.], [1., 0., 0.]])       # :194-201
    return mixture, target, table[idx.long()]
