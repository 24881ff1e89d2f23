"""Drop-in for the reference's ``src/models/tfgridnet_realtime_clean_dis_embd3/net.py::Net`` (distance-embedding
variant): same constructor kwargs (net.py:21-26), same ``forward / predict / init_buffers`` signatures (:67-93), same
state-dict-of-tensors schema and ``state_dict`` keys, but the forward pass runs in hand-written sm_100a CUDA kernels.
Point ``pl_module_args.model`` at ``sound_bubble_b200.tfgridnet_realtime_clean_dis_embd3.net.Net``."""
from .._net_base import NetBase, mod_pad  # noqa: F401
from ..packing import ModelConfig


class Net(NetBase):
    variant = "dis_embed"

    def __init__(self, stft_chunk_size=160, stft_pad_size=120, stft_back_pad=0,
                 num_ch=2, D=64, B=6, I=1, J=1, L=0, H=128,
                 use_attn=False, lookahead=True, local_atten_len=100,
                 E=4, chunk_causal=False, num_src=1,
                 spectral_masking=False, use_first_ln=False, merge_method="None",
                 directional=False, conv_lstm=True, fb_type='stft', dis_type="conv3"):
        super(Net, self).__init__()
        self._setup(ModelConfig(
            variant="dis_embed", stft_chunk_size=stft_chunk_size, stft_pad_size=stft_pad_size,
            stft_back_pad=stft_back_pad, num_ch=num_ch, D=D, B=B, I=I, J=J, L=L, H=H, use_attn=use_attn,
            lookahead=lookahead, local_atten_len=local_atten_len, E=E, chunk_causal=chunk_causal, num_src=num_src,
            spectral_masking=spectral_masking, use_first_ln=use_first_ln, merge_method=merge_method,
            directional=directional, conv_lstm=conv_lstm, fb_type=fb_type, dis_type=dis_type))

    def predict(self, x, dis_embed, input_state, pad=True):
        return self._predict(x, dis_embed, input_state, pad)

    def forward(self, inputs, input_state=None, pad=True):
        x = inputs['mixture']
        dis_embed = inputs['dis_embed']
        if self._wants_grad(input_state):
            return self._train_forward(x, dis_embed, pad)
        if input_state is None:
            input_state = self.init_buffers(x.shape[0], x.device)
        x, next_state = self.predict(x, dis_embed, input_state, pad)
        return {'output': x, 'next_state': next_state}
