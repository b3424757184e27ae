"""ctypes binding of libvfmops.so (the C ABI declared in include/vfm_ops.h).

There is no fallback: if the library is missing or a call fails, a RuntimeError is raised.  Nothing in this
package computes the ops on the CPU or through stock PyTorch kernels.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('VFM_LIB_PATH') or os.path.join(_HERE, 'lib', 'libvfmops.so')   # override: A/B builds of the library only

VFM_F16, VFM_F32, VFM_F64 = 0, 1, 2
VFM_OK, VFM_ERR_NO_KERNEL, VFM_ERR_INVALID, VFM_ERR_CUDA, VFM_ERR_WORKSPACE = 0, -1, -2, -3, -4
NOISE_NONE, NOISE_HW, NOISE_N1HW = 0, 1, 2

_i32, _i64, _f32, _f64, _vp, _sz = C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_void_p, C.c_size_t


class BiasActParams(C.Structure):
    _fields_ = [('x', _vp), ('b', _vp), ('xref', _vp), ('yref', _vp), ('dy', _vp), ('y', _vp), ('db', _vp),
                ('dtype', _i32), ('grad', _i32), ('act', _i32), ('alpha', _f64), ('gain', _f64), ('clamp', _f64),
                ('size_x', _i64), ('size_b', _i64), ('step_b', _i64)]


class Upfirdn2dParams(C.Structure):
    _fields_ = [('x', _vp), ('f', _vp), ('y', _vp), ('dtype', _i32),
                ('upx', _i32), ('upy', _i32), ('downx', _i32), ('downy', _i32),
                ('padx0', _i32), ('pady0', _i32), ('flip', _i32), ('gain', _f64),
                ('in_w', _i32), ('in_h', _i32), ('channels', _i32), ('batch', _i32),
                ('in_stride_w', _i64), ('in_stride_h', _i64), ('in_stride_c', _i64), ('in_stride_n', _i64),
                ('fw', _i32), ('fh', _i32), ('f_stride_w', _i64), ('f_stride_h', _i64),
                ('out_w', _i32), ('out_h', _i32),
                ('out_stride_w', _i64), ('out_stride_h', _i64), ('out_stride_c', _i64), ('out_stride_n', _i64),
                ('add', _vp), ('add_stride_h', _i64), ('add_stride_n', _i64),
                ('ep_enable', _i32), ('ep_act', _i32), ('ep_alpha', _f64), ('ep_gain', _f64), ('ep_clamp', _f64), ('ep_bias', _vp),
                ('pad_mode', _i32), ('f_stride_c', _i64)]


class FilteredLreluParams(C.Structure):
    _fields_ = [('x', _vp), ('y', _vp), ('b', _vp), ('s', _vp), ('fu', _vp), ('fd', _vp), ('dtype', _i32),
                ('up', _i32), ('down', _i32), ('fu_w', _i32), ('fu_h', _i32), ('fd_w', _i32), ('fd_h', _i32),
                ('fu_stride_w', _i64), ('fu_stride_h', _i64), ('fd_stride_w', _i64), ('fd_stride_h', _i64),
                ('pad_x0', _i32), ('pad_y0', _i32), ('gain', _f32), ('slope', _f32), ('clamp', _f32),
                ('flip', _i32), ('write_signs', _i32), ('read_signs', _i32),
                ('x_w', _i32), ('x_h', _i32), ('channels', _i32), ('batch', _i32),
                ('x_stride_w', _i64), ('x_stride_h', _i64), ('x_stride_c', _i64), ('x_stride_n', _i64),
                ('y_w', _i32), ('y_h', _i32),
                ('y_stride_w', _i64), ('y_stride_h', _i64), ('y_stride_c', _i64), ('y_stride_n', _i64),
                ('b_stride', _i64), ('s_w_bytes', _i32), ('s_h', _i32), ('s_ofs_x', _i32), ('s_ofs_y', _i32),
                ('s_w_active', _i32), ('y_sum', _vp)]


class FilteredLreluActParams(C.Structure):
    _fields_ = [('x', _vp), ('s', _vp), ('dtype', _i32), ('gain', _f64), ('slope', _f64), ('clamp', _f64),
                ('write_signs', _i32), ('read_signs', _i32),
                ('x_w', _i32), ('x_h', _i32), ('channels', _i32), ('batch', _i32),
                ('x_stride_w', _i64), ('x_stride_h', _i64), ('x_stride_c', _i64), ('x_stride_n', _i64),
                ('s_w', _i32), ('s_h', _i32), ('s_ofs_x', _i32), ('s_ofs_y', _i32)]


class ModconvDesc(C.Structure):
    _fields_ = [('dtype', _i32), ('batch', _i32), ('in_channels', _i32), ('out_channels', _i32),
                ('in_h', _i32), ('in_w', _i32), ('kh', _i32), ('kw', _i32), ('up', _i32), ('padding', _i32),
                ('demodulate', _i32), ('flip_weight', _i32), ('noise_mode', _i32),
                ('resample_filter', _vp), ('fw', _i32), ('fh', _i32), ('out_h', _i32), ('out_w', _i32),
                ('force_generic', _i32)]


class ModconvFwdParams(C.Structure):
    _fields_ = [('d', ModconvDesc), ('x', _vp), ('weight', _vp), ('styles', _vp), ('noise', _vp), ('y', _vp),
                ('dcoefs', _vp), ('workspace', _vp), ('workspace_bytes', _sz),
                ('ep_enable', _i32), ('ep_act', _i32), ('ep_alpha', _f64), ('ep_gain', _f64), ('ep_clamp', _f64),
                ('ep_bias', _vp), ('ep_residual', _vp), ('ep_gamma', _vp), ('ep_res_scale', _f64),
                ('x_scale', _vp), ('x_shift', _vp), ('ep_res_affine', _i32), ('keep_operand', _i32)]


class GroupNormAffineParams(C.Structure):
    _fields_ = [('x', _vp), ('gamma', _vp), ('beta', _vp), ('scale', _vp), ('shift', _vp), ('dtype', _i32),
                ('batch', _i32), ('channels', _i32), ('groups', _i32), ('hw', _i64), ('eps', _f64)]


class GroupNormParams(C.Structure):
    _fields_ = [('x', _vp), ('gamma', _vp), ('beta', _vp), ('y', _vp), ('mean', _vp), ('rstd', _vp), ('scratch', _vp), ('dy', _vp), ('dx', _vp),
                ('dgamma_nc', _vp), ('dbeta_nc', _vp), ('dtype', _i32), ('batch', _i32), ('channels', _i32), ('groups', _i32), ('hw', _i64), ('eps', _f64)]


class RowsParams(C.Structure):
    _fields_ = [('a', _vp), ('b', _vp), ('P', _vp), ('Q', _vp), ('R', _vp), ('out', _vp), ('dtype', _i32), ('rows', _i64), ('hw', _i64)]


class PixelShuffle2Params(C.Structure):
    _fields_ = [('x', _vp), ('y', _vp), ('dtype', _i32), ('batch', _i32), ('out_channels', _i32), ('in_h', _i32), ('in_w', _i32), ('inverse', _i32)]


class ReplicateBlurEdgesParams(C.Structure):
    _fields_ = [('dy', _vp), ('dx', _vp), ('f', _vp), ('dtype', _i32), ('k', _i32), ('planes', C.c_int64), ('h', _i32), ('w', _i32)]


class DepthwiseWgradParams(C.Structure):
    _fields_ = [('x', _vp), ('dy', _vp), ('dweight', _vp), ('dbias', _vp), ('dtype', _i32), ('batch', _i32), ('channels', _i32), ('h', _i32),
                ('w', _i32), ('k', _i32)]


class GradFinalizeParams(C.Structure):
    _fields_ = [('grads', _vp), ('numel', _i64), ('world_size', _i32), ('use_gain', _i32), ('gain', _f64), ('nan', _f64), ('posinf', _f64),
                ('neginf', _f64)]


class ImageToU8Params(C.Structure):
    _fields_ = [('x', _vp), ('y', _vp), ('dtype', _i32), ('batch', _i32), ('channels', _i32), ('height', _i32), ('width', _i32),
                ('pre_add', _f64), ('pre_div', _f64), ('scale', _f64)]


class ModconvBwdParams(C.Structure):
    _fields_ = [('d', ModconvDesc), ('dy', _vp), ('x', _vp), ('y', _vp), ('weight', _vp), ('styles', _vp),
                ('noise', _vp), ('dcoefs', _vp), ('dx', _vp), ('dweight', _vp), ('dstyles', _vp), ('dnoise', _vp),
                ('workspace', _vp), ('workspace_bytes', _sz), ('saved_operand', _vp), ('saved_operand_lo', _vp),
                ('ep_enable', _i32), ('ep_act', _i32), ('ep_alpha', _f64), ('ep_gain', _f64), ('ep_clamp', _f64), ('ep_bias', _vp), ('dbias_no', _vp)]


class KernelStat(C.Structure):
    _fields_ = [('name', C.c_char * 64), ('launches', _i64), ('total_ms', _f64), ('flops', _f64), ('bytes', _f64)]


#: every symbol include/vfm_ops.h declares, with (restype, argtypes)
SYMBOLS = {
    'vfm_timing_enable': (None, [C.c_int]),
    'vfm_timing_report': (C.c_int, [C.POINTER(KernelStat), C.c_int]),
    'vfm_last_error': (C.c_char_p, []),
    'vfm_abi_version': (C.c_int, []),
    'vfm_launch_count': (C.c_uint64, []),
    'vfm_bias_act': (C.c_int, [C.POINTER(BiasActParams), _vp]),
    'vfm_upfirdn2d': (C.c_int, [C.POINTER(Upfirdn2dParams), _vp]),
    'vfm_filtered_lrelu': (C.c_int, [C.POINTER(FilteredLreluParams), _vp]),
    'vfm_filtered_lrelu_act': (C.c_int, [C.POINTER(FilteredLreluActParams), _vp]),
    'vfm_modconv_workspace_bytes': (_sz, [C.POINTER(ModconvDesc), C.c_int]),
    'vfm_modconv_forward': (C.c_int, [C.POINTER(ModconvFwdParams), _vp]),
    'vfm_modconv_backward': (C.c_int, [C.POINTER(ModconvBwdParams), _vp]),
    'vfm_modconv_uses_tensor_cores': (C.c_int, [C.POINTER(ModconvDesc)]),
    'vfm_modconv_fused_backward_supported': (C.c_int, [C.POINTER(ModconvDesc)]),
    'vfm_modconv_forward_operand': (C.c_int, [C.POINTER(ModconvDesc), _vp, _sz, C.POINTER(_vp), C.POINTER(_vp)]),
    'vfm_group_norm_affine': (C.c_int, [C.POINTER(GroupNormAffineParams), _vp]),
    'vfm_group_norm_forward': (C.c_int, [C.POINTER(GroupNormParams), _vp]),
    'vfm_group_norm_backward': (C.c_int, [C.POINTER(GroupNormParams), _vp]),
    'vfm_rows_affine': (C.c_int, [C.POINTER(RowsParams), _vp]),
    'vfm_rows_dot': (C.c_int, [C.POINTER(RowsParams), _vp]),
    'vfm_pixel_shuffle2': (C.c_int, [C.POINTER(PixelShuffle2Params), _vp]),
    'vfm_replicate_blur_edges': (C.c_int, [C.POINTER(ReplicateBlurEdgesParams), _vp]),
    'vfm_depthwise_wgrad': (C.c_int, [C.POINTER(DepthwiseWgradParams), _vp]),
    'vfm_grad_finalize': (C.c_int, [C.POINTER(GradFinalizeParams), _vp]),
    'vfm_image_to_u8': (C.c_int, [C.POINTER(ImageToU8Params), _vp]),
}

_lib = None


def load():
    """Load libvfmops.so (once).  Raises RuntimeError with build instructions if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'vfm_vae_b200: CUDA library not found at {LIB_PATH}. Build it with '
            f'`python -c "import __graft_entry__ as g; g.build()"` or `make -C vfm_vae_b200/csrc`. '
            f'There is no CPU or PyTorch fallback for these ops.')
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)       # AttributeError here = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.vfm_abi_version() != 10:
        raise RuntimeError(f'vfm_vae_b200: ABI version mismatch ({lib.vfm_abi_version()} != 10)')
    _lib = lib
    return lib


def last_error():
    return load().vfm_last_error().decode()


def check(status, what):
    """Translate a C-ABI status into the reference's error behaviour (TORCH_CHECK -> RuntimeError)."""
    if status == VFM_OK:
        return
    raise RuntimeError(f'{what}: {last_error()} (status {status})')


def launch_count():
    return int(load().vfm_launch_count())
