"""Drop-in for the reference's ``src/models/tfgridnet_realtime_clean_optim/net.py::Net`` (embedded / "optim" variant:
no distance embedding, ``lstm_down`` exposed, output_padding in the conv-LSTM tail): same constructor kwargs
(net.py:21-26), same ``forward / predict / init_buffers`` signatures (:67-92), same checkpoint keys; the forward pass
runs in hand-written sm_100a CUDA kernels."""
from .._net_base import NetBase, mod_pad  # noqa: F401
from ..packing import ModelConfig


class Net(NetBase):
    variant = "optim"

    def __init__(self, stft_chunk_size=160, stft_pad_size=120, stft_back_pad=0,
                 num_ch=2, D=64, B=6, I=1, J=1, L=0, H=128,
                 use_attn=False, lookahead=True, local_atten_len=100,
                 E=4, chunk_causal=False, num_src=1,
                 spectral_masking=False, use_first_ln=False, merge_method="None",
                 directional=False, conv_lstm=True, lstm_down=5, fb_type='stft'):
        super(Net, self).__init__()
        self._setup(ModelConfig(
            variant="optim", stft_chunk_size=stft_chunk_size, stft_pad_size=stft_pad_size,
            stft_back_pad=stft_back_pad, num_ch=num_ch, D=D, B=B, I=I, J=J, L=L, H=H, use_attn=use_attn,
            lookahead=lookahead, local_atten_len=local_atten_len, E=E, chunk_causal=chunk_causal, num_src=num_src,
            spectral_masking=spectral_masking, use_first_ln=use_first_ln, merge_method=merge_method,
            directional=directional, conv_lstm=conv_lstm, fb_type=fb_type, lstm_down=lstm_down))

    def predict(self, x, input_state, pad=True):
        return self._predict(x, None, input_state, pad)

    def forward(self, inputs, input_state=None, pad=True):
        x = inputs['mixture']
        if self._wants_grad(input_state):
            return self._train_forward(x, None, pad)
        if input_state is None:
            input_state = self.init_buffers(x.shape[0], x.device)
        x, next_state = self.predict(x, input_state, pad)
        return {'output': x, 'next_state': next_state}
