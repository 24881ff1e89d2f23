"""Named parity cases shared by oracle/make_golden.py, the tests and bench.py.  TEST INFRASTRUCTURE ONLY."""

# /root/reference/syn_experiments/{pretrain,finetune}_stage.json:8-27  (the "TFG_S" configuration)
SYN = dict(stft_chunk_size=192, stft_pad_size=96, num_ch=6, D=32, L=4, I=1, J=1, B=6, H=64, E=2,
           local_atten_len=100, use_attn=False, lookahead=True, chunk_causal=True, use_first_ln=True,
           merge_method="early_cat", conv_lstm=False, dis_type="conv3")

# /root/reference/real_experiments/raspberrypi_model_finetune.json:8-26
RPI = dict(stft_chunk_size=192, stft_pad_size=96, num_ch=6, D=16, L=4, I=1, J=1, B=3, H=64, E=2, conv_lstm=True,
           lstm_down=5, local_atten_len=50, use_attn=False, lookahead=True, chunk_causal=True, use_first_ln=True,
           merge_method="early_cat")

# /root/reference/real_experiments/orangpi_model_finetune.json:8-25
OPI = dict(stft_chunk_size=192, stft_pad_size=96, num_ch=6, D=32, L=4, I=1, J=1, B=6, H=64, E=2,
           local_atten_len=100, use_attn=False, lookahead=True, chunk_causal=True, use_first_ln=True,
           merge_method="early_cat", conv_lstm=False)


def _with(base, **kw):
    d = dict(base)
    d.update(kw)
    return d


CASES = {
    # offline call with mod-pad + look-ahead pad, then a second, state-carrying call (pad=False)
    "syn_offline": dict(variant="dis_embed", kwargs=SYN, batch=2, n_samples=192 * 12 - 46, second_call=192 * 3 + 96),
    # exact multiple of the chunk, single frame per call is exercised by the streaming tests against this
    "syn_nopad": dict(variant="dis_embed", kwargs=SYN, batch=3, n_samples=192 * 7 + 96, pad=False),
    "syn_attn": dict(variant="dis_embed", kwargs=_with(SYN, use_attn=True, local_atten_len=10), batch=2,
                     n_samples=192 * 16 + 96, pad=False, second_call=192 * 2 + 96),
    "syn_convlstm": dict(variant="dis_embed", kwargs=_with(SYN, conv_lstm=True), batch=2, n_samples=192 * 6 + 96,
                         pad=False),
    "syn_directional": dict(variant="dis_embed", kwargs=_with(SYN, directional=True), batch=1,
                            n_samples=192 * 5 + 96, pad=False),
    "syn_masking": dict(variant="dis_embed", kwargs=_with(SYN, spectral_masking=True, B=2), batch=1,
                        n_samples=192 * 5 + 96, pad=False),
    "syn_plain": dict(variant="dis_embed", kwargs=_with(SYN, merge_method="None", use_first_ln=False, B=2), batch=1,
                      n_samples=192 * 5 + 96, pad=False),
    "opi_offline": dict(variant="optim", kwargs=OPI, batch=2, n_samples=192 * 6 + 96, pad=False),
    "rpi_offline": dict(variant="optim", kwargs=RPI, batch=2, n_samples=192 * 10 - 7, second_call=192 * 2 + 96),
    "rpi_k4": dict(variant="optim", kwargs=_with(RPI, lstm_down=4), batch=1, n_samples=192 * 5 + 96, pad=False),
    "rpi_attn": dict(variant="optim", kwargs=_with(RPI, use_attn=True, local_atten_len=7), batch=2,
                     n_samples=192 * 12 + 96, pad=False),
    # BASELINE config 1: a real fixture clip (first 0.5 s), radius 1 m
    "wav_syn_1m": dict(variant="dis_embed", kwargs=SYN, wav="test_samples/syn_1m/00002/mixture.wav",
                       n_samples=12000, radius=[0.0, 0.0, 1.0]),
}

# gradient fixtures (oracle/make_golden_grads.py): the reference module in train() mode, L = sum(output * R)
GRAD_CASES = {
    # TFG_S architecture with 2 blocks (FiLM + distance embedding, first LayerNorm, mod-pad path of Net.forward)
    "grad_syn_b2": dict(variant="dis_embed", kwargs=_with(SYN, B=2), batch=2, n_samples=192 * 4 - 30, loss_seed=77),
    # the OPT variant (no distance embedding) at D = 16 with plain BiLSTM blocks
    "grad_opi_d16": dict(variant="optim", kwargs=_with(OPI, B=1, D=16), batch=1, n_samples=192 * 3, loss_seed=78),
}
GRAD_CASES.update({
    # the Raspberry-Pi model (conv-LSTM intra path with output_padding, k = 5, D = 16, 3 blocks): BASELINE config 5's model
    "grad_rpi": dict(variant="optim", kwargs=RPI, batch=2, n_samples=192 * 3 - 11, loss_seed=79),
    # DE3 conv-LSTM (pad-and-crop tail, k = 4, D = 32) with FiLM, 2 blocks
    "grad_syn_convlstm": dict(variant="dis_embed", kwargs=_with(SYN, conv_lstm=True, B=2), batch=1, n_samples=192 * 3, loss_seed=80),
})
GRAD_CASES.update({
    # TFG_S block with the sliding-window attention switched on (dormant in the shipped configs), window 5 over 6 frames
    "grad_syn_attn": dict(variant="dis_embed", kwargs=_with(SYN, use_attn=True, local_atten_len=5, B=2), batch=2, n_samples=192 * 6,
                          loss_seed=81),
})
GRAD_CASES.update({
    # the benchmark architecture itself (TFG_S: 6 blocks, FiLM on 5 of them), every one of its 149 parameter tensors
    "grad_tfg_s": dict(variant="dis_embed", kwargs=SYN, batch=1, n_samples=192 * 3 - 50, loss_seed=82),
})
