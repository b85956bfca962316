"""ctypes binding of libb200ir.so (include/b200ir.h).  No fallback: if the library is missing the import of
any compute path raises, and on a box without an sm_100 GPU every entry point returns an error."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200IR_LIB: another build of the same library (A/B of two builds inside one GPU box; tools/ab_env.sh)
LIB_PATH = os.environ.get('B200IR_LIB') or os.path.join(HERE, 'libb200ir.so')

MAX_TAPS = 16
MAX_VIEWS = 4


class View(C.Structure):
    _fields_ = [('ptr', C.c_void_p), ('c', C.c_int32), ('w', C.c_int32), ('h', C.c_int32), ('b', C.c_int32),
                ('stride_w', C.c_int64), ('stride_h', C.c_int64), ('stride_b', C.c_int64)]


class ConvDesc(C.Structure):
    _fields_ = [
        ('a', View * MAX_VIEWS), ('num_views', C.c_int32),
        ('weight', C.c_void_p), ('cin', C.c_int32), ('cout', C.c_int32), ('num_taps', C.c_int32),
        ('tap_view', C.c_int8 * MAX_TAPS), ('tap_dx', C.c_int8 * MAX_TAPS), ('tap_dy', C.c_int8 * MAX_TAPS),
        ('m_w', C.c_int32), ('m_h', C.c_int32), ('m_b', C.c_int32),
        ('tile_w', C.c_int32), ('tile_h', C.c_int32), ('tile_b', C.c_int32), ('block_n', C.c_int32),
        ('out', C.c_void_p), ('out_fp32', C.c_int32),
        ('out_stride_x', C.c_int64), ('out_stride_y', C.c_int64), ('out_stride_b', C.c_int64),
        ('out_c_off', C.c_int32), ('out_x_mul', C.c_int32), ('out_x_off', C.c_int32),
        ('out_y_mul', C.c_int32), ('out_y_off', C.c_int32),
        ('bias', C.c_void_p), ('demod', C.c_void_p), ('noise', C.c_void_p), ('noise_gain', C.c_void_p),
        ('noise_stride_b', C.c_int64), ('noise_stride_y', C.c_int64),
        ('act', C.c_int32), ('res_mode', C.c_int32), ('res', C.c_void_p),
        ('res_stride_x', C.c_int64), ('res_stride_y', C.c_int64), ('res_stride_b', C.c_int64),
        ('res_w', C.c_int32), ('res_h', C.c_int32), ('res_scale', C.c_float), ('max_ctas', C.c_int32),
        ('row_mode', C.c_int32),
        ('out_scale', C.c_void_p), ('rgb_w', C.c_void_p), ('rgb_part', C.c_void_p),
        ('rgb_w_px', C.c_int32), ('rgb_h', C.c_int32), ('no_store', C.c_int32),
        ('act_slope', C.c_float), ('res_mul', C.c_float), ('ps_r', C.c_int32),
        ('ps_c', C.c_int32), ('demod_c', C.c_int32), ('use_tap_mask', C.c_int32), ('tap_mask', C.c_uint32 * 8),
        ('corr_top', C.c_void_p), ('corr_bot', C.c_void_p), ('corr_left', C.c_void_p), ('corr_right', C.c_void_p),
        ('w_per_image', C.c_int32),
    ]


class ModLayer(C.Structure):
    _fields_ = [('w', C.c_void_p), ('bias', C.c_void_p), ('s', C.c_void_p), ('lat_idx', C.c_int32), ('cin', C.c_int32)]


class DemodLayer(C.Structure):
    _fields_ = [('s', C.c_void_p), ('wsq', C.c_void_p), ('d', C.c_void_p), ('scale2', C.c_float),
                ('cin', C.c_int32), ('cout', C.c_int32)]


class DegradeCrop(C.Structure):
    """b200ir_degrade_crop (include/b200ir.h); numpy view: DEGRADE_CROP_DTYPE."""
    _fields_ = [('blur_mode', C.c_int32), ('ksize', C.c_int32), ('blur_f64', C.c_int32), ('lr_w', C.c_int32), ('lr_h', C.c_int32),
                ('jpeg_quality', C.c_int32), ('gray', C.c_int32), ('jitter', C.c_float * 3), ('bilateral_sigma', C.c_float), ('cj_count', C.c_int32),
                ('cj_order', C.c_int32 * 4), ('cj_factor', C.c_float * 4), ('cj_one_minus', C.c_float * 4), ('mask_mode', C.c_int32)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float

# name -> argtypes (return type is int unless listed in _RESTYPES); mirrors include/b200ir.h one to one
SIGNATURES = {
    'b200ir_last_error': [],
    'b200ir_abi_version': [],
    'b200ir_launch_count': [],
    'b200ir_device_check': [],
    'b200ir_conv_igemm': [C.POINTER(ConvDesc), _P],
    'b200ir_conv_plan_create': [C.POINTER(ConvDesc), C.POINTER(_P)],
    'b200ir_conv_plan_launch': [_P, _P],
    'b200ir_conv_plan_destroy': [_P],
    'b200ir_pack_weights': [_P, _P, _I, _I, _I, _I, _F, _I, _I, _P],
    'b200ir_first_conv': [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_replicate_border': [_P, _I, _I, _I, _I, _P],
    'b200ir_upfold_corners': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_tiles_gather': [_P, _P, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    'b200ir_tiles_blend': [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _I, _P],
    'b200ir_u8_to_input': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_conv_wgrad': [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_conv_wgrad_view': [C.POINTER(View), _P, _P, _I, _I, _I, _I, C.c_uint32, _P],
    'b200ir_lrelu_bias_bwd': [_P, _P, _P, _P, _L, _I, _F, _F, _P],
    'b200ir_fir_pad11': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'b200ir_fir_down2_adjoint': [_P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_bilinear_up2_adjoint': [_P, _P, _I, _I, _I, _I, _F, _P],
    'b200ir_first_conv_wgrad': [_P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_minibatch_stddev_bwd': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_adam_step': [_P, _P, _P, _P, _L, _F, _F, _F, _F, _F, _I, _F, _P, _F, _P],
    'b200ir_minibatch_stddev': [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_f32_to_input': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_image_to_u8': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_fir_pad22': [_P, _P, _I, _I, _I, _I, _I, _I, _P],
    'b200ir_fir_down2': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_bilinear_up2': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_add': [_P, _P, _P, _L, _P],
    'b200ir_upfir_act': [_P, _P, _I, _I, _I, _I, _I, _I, _P, _L, _P, _P, _P, _P, _I, _P, _P],
    'b200ir_to_rgb': [_P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P],
    'b200ir_rgb_combine': [_P, _I, _P, _P, _P, _I, _I, _I, _P],
    'b200ir_rgb_wmod': [_P, _P, _P, _I, _I, _P],
    'b200ir_modulate_const': [_P, _P, _P, _I, _I, _I, _P],
    'b200ir_mod_linear': [_P, _I, _I, _I, _P, _P, _F, _P, _I, _I, _P],
    'b200ir_demod': [_P, _P, _F, _P, _I, _I, _I, _P],
    'b200ir_mod_linear_multi': [_P, _I, _I, _P, _I, _I, _F, _I, _P],
    'b200ir_demod_multi': [_P, _I, _I, _I, _P],
    'b200ir_style_mlp': [_P, _P, _P, _P, _I, _I, _I, _F, _P],
    'b200ir_nhwc_to_nchw_f32': [_P, _P, _I, _I, _I, _P],
    'b200ir_nchw_to_nhwc_pad': [_P, _P, _I, _I, _I, _I, _I, _P, _F, _I, _P],
    'b200ir_nearest_up2': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_sr_output': [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _I, _P],
    'b200ir_channel_mean': [_P, _P, _I, _I, _I, _P],
    'b200ir_ca_mlp': [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P],
    'b200ir_ca_scale_add': [_P, _P, _P, _P, _F, _I, _I, _I, _L, _L, _P],
    'b200ir_sft_mod': [_P, _P, _P, _I, _P, _P, _I, _L, _I, _P],
    'b200ir_sft_mod_bwd': [_P, _P, _L, _P, _P, _I, _P, _P, _I, _P, _P, _P, _I, _L, _I, _P],
    'b200ir_style_act_bwd': [_P, _P, _P, _L, _P, _P, _P, _F, _P, _P, _I, _L, _I, _P],
    'b200ir_to_rgb_bwd': [_P, _P, _P, _P, _P, _I, _P, _I, _L, _I, _P],
    'b200ir_style_act_bwd_params': [_P, _P, _P, _L, _P, _P, _P, _F, _P, _P, _P, _P, _I, _L, _I, _P],
    'b200ir_to_rgb_bwd_params': [_P, _P, _P, _P, _P, _I, _P, _P, _I, _L, _I, _P],
    'b200ir_plane_sums': [_P, _P, _I, _I, _L, _P],
    'b200ir_wgrad_unfold': [_P, _P, _I, _I, _I, _P],
    'b200ir_table_colsum': [_P, _I, _P, _I, _F, _P, _I, _L, _P],
    'b200ir_mod_linear_wgrad': [_P, _P, _F, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_modconv_wgrad': [_P, _I, _P, _P, _P, _P, _F, _P, _I, _I, _I, _I, _P],
    'b200ir_rgb_up_adjoint': [_P, _P, _I, _I, _I, _P],
    'b200ir_demod_bwd': [_P, _P, _P, _P, _P, _F, _I, _I, _I, _P],
    'b200ir_mod_linear_bwd': [_P, _P, _F, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_first_conv_dgrad': [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_head_to_nchw': [_P, _P, _I, _L, _I, _P],
    'b200ir_nchw_to_head': [_P, _P, _I, _L, _I, _P],
    'b200ir_gram_batched': [_P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_maxpool2_relu': [_P, _P, _I, _I, _I, _I, _P],
    'b200ir_maxpool2_relu_bwd': [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_l1_loss_f16': [_P, _P, _L, _F, _F, _P, _P, _P],
    'b200ir_sum_squares': [_P, _L, _F, _P, _P],
    'b200ir_minibatch_stddev_jvp': [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    'b200ir_minibatch_stddev_hvp': [_P, _P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_l1_loss': [_P, _P, _L, _F, _F, _P, _P, _P],
    'b200ir_softplus_loss': [_P, _I, _I, _F, _F, _F, _P, _P, _P],
    'b200ir_degrade': [_P, _P, _P, _P, _I, _P, _P, _P, _I, _I, _P, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_degrade_full': [_P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _I, _I, _I, _I, _P],
    'b200ir_degrade_full_ex': [_P, _P, _P, _I, _P, _P, _I, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P],
}
_RESTYPES = {'b200ir_last_error': C.c_char_p, 'b200ir_launch_count': C.c_uint64, 'b200ir_conv_plan_destroy': None}

_lib = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} not found: build it with `python -m image_restoration_b200.build` '
                               '(or __graft_entry__.build()); image_restoration_b200 has no fallback path')
        handle = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        _lib = handle
    return _lib


def require_cuda(t, who):
    """Every compute entry point of the package takes CUDA tensors; there is no CPU path."""
    if not t.is_cuda:
        raise RuntimeError(f'image_restoration_b200.{who} needs CUDA tensors (no CPU path)')


def device_ctx(dev):
    """`with` context selecting the CUDA device of a tensor for the launches inside."""
    import torch
    return torch.cuda.device(dev)


class B200irError(RuntimeError):
    pass


def check(status, what=''):
    if status != 0:
        raise B200irError(f'{what}: {lib().b200ir_last_error().decode()}')
