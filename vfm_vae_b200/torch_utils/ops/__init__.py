from . import bias_act, upfirdn2d, filtered_lrelu, conv2d_resample, fma, modulated_conv2d  # noqa: F401
