"""ctypes mirror of include/soundbubble.h (struct layouts, constants, prototypes).  Pure declarations: no library is
loaded here (see _lib.py for the product loader; tests/emu has its own loader for the host-emulated test build)."""
from __future__ import annotations

import ctypes as C

SB_VERSION = 100
SB_MAX_BLOCKS = 16
SB_MAX_MICS = 8

SB_ALGO_AUTO, SB_ALGO_TILE, SB_ALGO_LANE1, SB_ALGO_LANE2, SB_ALGO_LANE4, SB_ALGO_WS, SB_ALGO_TILE4, SB_ALGO_TC = 0, 1, 2, 3, 4, 5, 6, 7
SB_ALGO_WS2 = 8
SB_ALGO_TCP = 9
SB_ALGO_TCQ = 10
SB_FEAT_NONE, SB_FEAT_OMNI, SB_FEAT_DIRECTIONAL = 0, 1, 2
SB_EMB_CONV, SB_EMB_LINEAR = 0, 1
SB_CONVLSTM_PADCROP, SB_CONVLSTM_OUTPAD = 0, 1
SB_OPT_PDL = 1
SB_OPT_ATTN_TC = 2
SB_OPT_TRAIN_ONE_ROW = 3
SB_OPT_TRAIN_FFMA2 = 4
SB_OPT_TC_V1 = 5
SB_OPT_TC_CELL7 = 6
SB_OPT_TRAIN_TC = 7
SB_OPT_TC_PIPE = 8
SB_OPT_TC_CW16 = 9
SB_OPT_FRONT_TC = 10
SB_STAGES = ("stft_features", "conv_in", "film_params", "intra", "inter", "attention", "backend")

fp = C.c_void_p          # device float* (raw address)


class LstmDir(C.Structure):
    _fields_ = [(n, fp) for n in ("w_tile", "b_tile", "w_lane", "b_lane", "w_rec", "w_xp", "w_prj", "tc_w", "tc_b", "lin_t", "lin_n", "lin_b", "ln_g", "ln_b")]


class StftArgs(C.Structure):
    _fields_ = [("wave", fp), ("filt", fp), ("feats", fp), ("spec", fp),
                ("B", C.c_int), ("M", C.c_int), ("n_samples", C.c_int), ("T", C.c_int),
                ("n_fft", C.c_int), ("stride", C.c_int), ("F", C.c_int),
                ("feat_mode", C.c_int), ("Cin", C.c_int), ("n_src", C.c_int)]


class ConvInArgs(C.Structure):
    _fields_ = [("feats", fp), ("conv_buf_in", fp), ("conv_buf_out", fp), ("w_pack", fp), ("bias", fp),
                ("ln_g", fp), ("ln_b", fp), ("x", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("Cin", C.c_int), ("C", C.c_int)]


class FilmArgs(C.Structure):
    _fields_ = [("dis", fp), ("emb_w", fp), ("emb_ln_g", fp), ("emb_ln_b", fp),
                ("w_w", fp), ("w_b", fp), ("b_w", fp), ("b_b", fp), ("film", fp),
                ("B", C.c_int), ("F", C.c_int), ("C", C.c_int), ("Din", C.c_int), ("n_layers", C.c_int),
                ("emb_mode", C.c_int)]


class IntraArgs(C.Structure):
    _fields_ = [("x", fp), ("film_scale", fp), ("film_shift", fp), ("y_fwd", fp), ("y_bwd", fp),
                ("dir", LstmDir * 2),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("H", C.c_int), ("algo", C.c_int)]


class IntraConvArgs(C.Structure):
    _fields_ = [("x", fp), ("film_scale", fp), ("film_shift", fp), ("y", fp),
                ("conv_w", fp), ("conv_b", fp), ("prelu", fp), ("deconv_w", fp), ("deconv_b", fp),
                ("dir", LstmDir * 2), ("ws", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("H", C.c_int),
                ("down", C.c_int), ("tail_mode", C.c_int), ("algo", C.c_int)]


class InterArgs(C.Structure):
    _fields_ = [("x0", fp), ("x1", fp), ("y", fp), ("h0", fp), ("c0", fp), ("hN", fp), ("cN", fp),
                ("dir", LstmDir),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("H", C.c_int), ("algo", C.c_int)]


class AttnProj(C.Structure):
    _fields_ = [(n, fp) for n in ("w", "b", "prelu", "ln_g", "ln_b")]


class AttnArgs(C.Structure):
    _fields_ = [("x", fp), ("y", fp), ("q", AttnProj), ("k", AttnProj), ("v", AttnProj), ("o", AttnProj),
                ("K_buf_in", fp), ("K_buf_out", fp), ("V_buf_in", fp), ("V_buf_out", fp), ("ws", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("L", C.c_int), ("E", C.c_int),
                ("W", C.c_int)]


class BackendArgs(C.Structure):
    _fields_ = [("x", fp), ("deconv_buf_in", fp), ("deconv_buf_out", fp), ("istft_buf_in", fp), ("istft_buf_out", fp),
                ("w", fp), ("bias", fp), ("filt", fp), ("mask_spec", fp), ("wave_out", fp), ("ws", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("n_src", C.c_int),
                ("n_fft", C.c_int), ("stride", C.c_int)]


class BlockDesc(C.Structure):
    _fields_ = [("intra", LstmDir * 2), ("inter", LstmDir),
                ("cl_conv_w", fp), ("cl_conv_b", fp), ("cl_prelu", fp), ("cl_deconv_w", fp), ("cl_deconv_b", fp),
                ("attn_q", AttnProj), ("attn_k", AttnProj), ("attn_v", AttnProj), ("attn_o", AttnProj)]


class NetDesc(C.Structure):
    _fields_ = [("M", C.c_int), ("n_fft", C.c_int), ("stride", C.c_int), ("F", C.c_int),
                ("C", C.c_int), ("H", C.c_int), ("n_blocks", C.c_int), ("n_src", C.c_int),
                ("feat_mode", C.c_int), ("Cin", C.c_int), ("film_din", C.c_int), ("emb_mode", C.c_int),
                ("spectral_masking", C.c_int),
                ("conv_lstm", C.c_int), ("lstm_down", C.c_int), ("tail_mode", C.c_int),
                ("use_attn", C.c_int), ("L", C.c_int), ("E", C.c_int), ("W", C.c_int),
                ("enc_filt", fp), ("dec_filt", fp), ("conv_w_pack", fp), ("conv_bias", fp),
                ("conv_ln_g", fp), ("conv_ln_b", fp),
                ("emb_w", fp), ("emb_ln_g", fp), ("emb_ln_b", fp),
                ("film_w_w", fp), ("film_w_b", fp), ("film_b_w", fp), ("film_b_b", fp),
                ("deconv_w", fp), ("deconv_bias", fp),
                ("blocks", BlockDesc * SB_MAX_BLOCKS)]


class NetIO(C.Structure):
    _fields_ = [("wave", fp), ("dis_embed", fp), ("film", fp), ("wave_out", fp),
                ("conv_buf_in", fp), ("conv_buf_out", fp),
                ("deconv_buf_in", fp), ("deconv_buf_out", fp),
                ("istft_buf_in", fp), ("istft_buf_out", fp),
                ("h_in", fp * SB_MAX_BLOCKS), ("h_out", fp * SB_MAX_BLOCKS),
                ("c_in", fp * SB_MAX_BLOCKS), ("c_out", fp * SB_MAX_BLOCKS),
                ("K_in", fp * SB_MAX_BLOCKS), ("K_out", fp * SB_MAX_BLOCKS),
                ("V_in", fp * SB_MAX_BLOCKS), ("V_out", fp * SB_MAX_BLOCKS),
                ("workspace", fp),
                ("B", C.c_int), ("T", C.c_int), ("intra_algo", C.c_int), ("inter_algo", C.c_int)]


class PrepareArgs(C.Structure):
    _fields_ = [("mix", C.c_void_p), ("voices", C.c_void_p), ("inside", C.c_void_p), ("gain", fp), ("shift", C.c_void_p),
                ("drop", C.c_void_p), ("peak_scale", fp), ("radius_idx", C.c_void_p),
                ("mixture", fp), ("target", fp), ("dis_embed", fp), ("peak_ws", fp),
                ("B", C.c_int), ("M", C.c_int), ("V", C.c_int), ("N", C.c_int)]


class PathTrainArgs(C.Structure):
    _fields_ = [("x", fp), ("y", fp), ("ln_g", fp), ("ln_b", fp),
                ("w_ih", fp * 2), ("w_hh", fp * 2), ("b_ih", fp * 2), ("b_hh", fp * 2),
                ("lin_w", fp), ("lin_b", fp), ("saved", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("H", C.c_int), ("inter", C.c_int)]


class PathBwdArgs(C.Structure):
    _fields_ = [("f", PathTrainArgs), ("gy", fp), ("gx", fp), ("g_ln_g", fp), ("g_ln_b", fp),
                ("g_w_ih", fp * 2), ("g_w_hh", fp * 2), ("g_b_ih", fp * 2), ("g_b_hh", fp * 2),
                ("g_lin_w", fp), ("g_lin_b", fp), ("ws", fp)]


class FilmApplyArgs(C.Structure):
    _fields_ = [("x", fp), ("film_scale", fp), ("film_shift", fp), ("y", fp), ("gy", fp), ("gx", fp),
                ("g_scale", fp), ("g_shift", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int)]


class FilmBwdArgs(C.Structure):
    _fields_ = [("f", FilmArgs), ("g_film", fp), ("g_emb_w", fp), ("g_emb_ln_g", fp), ("g_emb_ln_b", fp),
                ("g_w_w", fp), ("g_w_b", fp), ("g_b_w", fp), ("g_b_b", fp)]


class ConvInTrainArgs(C.Structure):
    _fields_ = [("feats", fp), ("w", fp), ("bias", fp), ("ln_g", fp), ("ln_b", fp), ("x", fp), ("saved", fp),
                ("gx", fp), ("g_w", fp), ("g_bias", fp), ("g_ln_g", fp), ("g_ln_b", fp), ("ws", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("Cin", C.c_int), ("C", C.c_int)]


class BackendBwdArgs(C.Structure):
    _fields_ = [("x", fp), ("g_wave", fp), ("w", fp), ("filt", fp), ("mask_spec", fp), ("gx", fp), ("g_w", fp),
                ("g_bias", fp), ("ws", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("n_src", C.c_int),
                ("n_fft", C.c_int), ("stride", C.c_int)]


class ConvPathTrainArgs(C.Structure):
    _fields_ = [("x", fp), ("y", fp), ("conv_w", fp), ("conv_b", fp), ("prelu", fp), ("ln_g", fp), ("ln_b", fp),
                ("w_ih", fp * 2), ("w_hh", fp * 2), ("b_ih", fp * 2), ("b_hh", fp * 2),
                ("deconv_w", fp), ("deconv_b", fp), ("saved", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("H", C.c_int),
                ("down", C.c_int), ("tail_mode", C.c_int)]


class ConvPathBwdArgs(C.Structure):
    _fields_ = [("f", ConvPathTrainArgs), ("gy", fp), ("gx", fp),
                ("g_conv_w", fp), ("g_conv_b", fp), ("g_prelu", fp), ("g_ln_g", fp), ("g_ln_b", fp),
                ("g_w_ih", fp * 2), ("g_w_hh", fp * 2), ("g_b_ih", fp * 2), ("g_b_hh", fp * 2),
                ("g_deconv_w", fp), ("g_deconv_b", fp), ("ws", fp)]


class AttnTrainArgs(C.Structure):
    _fields_ = [("x", fp), ("y", fp), ("q", AttnProj), ("k", AttnProj), ("v", AttnProj), ("o", AttnProj), ("saved", fp),
                ("B", C.c_int), ("T", C.c_int), ("F", C.c_int), ("C", C.c_int), ("L", C.c_int), ("E", C.c_int), ("W", C.c_int)]


class AttnProjGrad(C.Structure):
    _fields_ = [(n, fp) for n in ("w", "b", "prelu", "ln_g", "ln_b")]


class AttnBwdArgs(C.Structure):
    _fields_ = [("f", AttnTrainArgs), ("gy", fp), ("gx", fp),
                ("gq", AttnProjGrad), ("gk", AttnProjGrad), ("gv", AttnProjGrad), ("go", AttnProjGrad), ("ws", fp)]


# index used by sb_abi_sizeof(which)
ABI_STRUCTS = {0: LstmDir, 1: StftArgs, 2: ConvInArgs, 3: FilmArgs, 4: IntraArgs, 5: InterArgs, 6: BackendArgs,
               7: NetDesc, 8: NetIO, 9: IntraConvArgs, 10: AttnProj, 11: AttnArgs, 12: BlockDesc, 13: PrepareArgs,
               14: PathTrainArgs, 15: PathBwdArgs, 16: FilmApplyArgs, 17: FilmBwdArgs, 18: ConvInTrainArgs, 19: BackendBwdArgs,
               20: ConvPathTrainArgs, 21: ConvPathBwdArgs, 22: AttnTrainArgs, 23: AttnProjGrad, 24: AttnBwdArgs}

# every symbol include/soundbubble.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "sb_set_option": (C.c_int, [C.c_int, C.c_int]),
    "sb_stft_features_fwd": (C.c_int, [C.POINTER(StftArgs), C.c_void_p]),
    "sb_conv_in_fwd": (C.c_int, [C.POINTER(ConvInArgs), C.c_void_p]),
    "sb_film_params_fwd": (C.c_int, [C.POINTER(FilmArgs), C.c_void_p]),
    "sb_intra_lstm_fwd": (C.c_int, [C.POINTER(IntraArgs), C.c_void_p]),
    "sb_intra_sum_supported": (C.c_int, [C.POINTER(IntraArgs)]),
    "sb_intra_convlstm_fwd": (C.c_int, [C.POINTER(IntraConvArgs), C.c_void_p]),
    "sb_inter_lstm_fwd": (C.c_int, [C.POINTER(InterArgs), C.c_void_p]),
    "sb_attn_workspace_floats": (C.c_size_t, [C.c_int] * 7),
    "sb_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), C.c_void_p]),
    "sb_backend_fwd": (C.c_int, [C.POINTER(BackendArgs), C.c_void_p]),
    "sb_workspace_floats": (C.c_size_t, [C.POINTER(NetDesc), C.c_int, C.c_int]),
    "sb_net_forward": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetIO), C.c_void_p]),
    "sb_net_forward_range": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetIO), C.c_int, C.c_int, C.c_void_p]),
    "sb_prepare_workspace_floats": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "sb_prepare_batch_fwd": (C.c_int, [C.POINTER(PrepareArgs), C.c_void_p]),
    "sb_path_train_saved_floats": (C.c_size_t, [C.c_int] * 6),
    "sb_path_bwd_workspace_floats": (C.c_size_t, [C.c_int] * 6),
    "sb_intra_lstm_train_fwd": (C.c_int, [C.POINTER(PathTrainArgs), C.c_void_p]),
    "sb_inter_lstm_train_fwd": (C.c_int, [C.POINTER(PathTrainArgs), C.c_void_p]),
    "sb_intra_lstm_bwd": (C.c_int, [C.POINTER(PathBwdArgs), C.c_void_p]),
    "sb_inter_lstm_bwd": (C.c_int, [C.POINTER(PathBwdArgs), C.c_void_p]),
    "sb_convpath_train_saved_floats": (C.c_size_t, [C.c_int] * 6),
    "sb_convpath_bwd_workspace_floats": (C.c_size_t, [C.c_int] * 6),
    "sb_intra_convlstm_train_fwd": (C.c_int, [C.POINTER(ConvPathTrainArgs), C.c_void_p]),
    "sb_intra_convlstm_bwd": (C.c_int, [C.POINTER(ConvPathBwdArgs), C.c_void_p]),
    "sb_attn_train_saved_floats": (C.c_size_t, [C.POINTER(AttnTrainArgs)]),
    "sb_attn_bwd_workspace_floats": (C.c_size_t, [C.POINTER(AttnTrainArgs)]),
    "sb_attn_train_fwd": (C.c_int, [C.POINTER(AttnTrainArgs), C.c_void_p]),
    "sb_attn_bwd": (C.c_int, [C.POINTER(AttnBwdArgs), C.c_void_p]),
    "sb_film_apply_fwd": (C.c_int, [C.POINTER(FilmApplyArgs), C.c_void_p]),
    "sb_film_apply_bwd": (C.c_int, [C.POINTER(FilmApplyArgs), C.c_void_p]),
    "sb_film_params_bwd": (C.c_int, [C.POINTER(FilmBwdArgs), C.c_void_p]),
    "sb_conv_in_train_fwd": (C.c_int, [C.POINTER(ConvInTrainArgs), C.c_void_p]),
    "sb_conv_in_bwd": (C.c_int, [C.POINTER(ConvInTrainArgs), C.c_void_p]),
    "sb_backend_bwd": (C.c_int, [C.POINTER(BackendBwdArgs), C.c_void_p]),
    "sb_pipe_create": (C.c_int, [C.POINTER(NetDesc), C.POINTER(NetIO), C.c_int, C.c_int, C.POINTER(C.c_int),
                                 C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]),
    "sb_pipe_destroy": (C.c_int, [C.c_void_p]),
    "sb_pipe_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_pipe_feed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_pipe_feed_chunk": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "sb_pipe_flush": (C.c_int, [C.c_void_p]),
    "sb_pipe_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "sb_pipe_reset": (C.c_int, [C.c_void_p]),
    "sb_pipe_calls": (C.c_longlong, [C.c_void_p]),
    "sb_profile_begin": (C.c_int, []),
    "sb_profile_end": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "sb_version": (C.c_int, []),
    "sb_last_error_string": (C.c_char_p, []),
    "sb_launch_count": (C.c_uint64, []),
    "sb_abi_sizeof": (C.c_int, [C.c_int]),
}


class SoundBubbleError(RuntimeError):
    pass


def bind(cdll):
    """Attach prototypes to a loaded CDLL and verify version + struct layouts.  Raises if any symbol is missing."""
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(cdll, name)            # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if cdll.sb_version() != SB_VERSION:
        raise SoundBubbleError("libsoundbubble version %d != binding version %d" % (cdll.sb_version(), SB_VERSION))
    for which, st in ABI_STRUCTS.items():
        got = cdll.sb_abi_sizeof(which)
        if got != C.sizeof(st):
            raise SoundBubbleError("ABI mismatch: %s is %d bytes in the library, %d in the binding"
                                   % (st.__name__, got, C.sizeof(st)))
    return cdll


def check(cdll, rc: int, what: str):
    if rc != 0:
        msg = cdll.sb_last_error_string()
        raise SoundBubbleError("%s failed (rc=%d): %s" % (what, rc, msg.decode() if msg else "?"))
