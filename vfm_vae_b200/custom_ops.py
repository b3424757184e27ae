"""Mirror of the reference plugin loader (torch_utils/custom_ops.py:59 ``get_plugin``).

The reference JIT-compiles each plugin with nvcc on first use and caches the module.  Here the kernels are
pre-built into libvfmops.so, so ``get_plugin`` just hands out the matching plugin object; the signature is kept
so that the reference's op wrappers (which call ``custom_ops.get_plugin(module_name=..., sources=..., ...)``) work
unchanged when this module is patched over theirs (see integration.py)."""
from . import _lib
from .plugins import PLUGINS

verbosity = 'brief'  # kept for API compatibility; unused

_cached_plugins = dict()


def get_plugin(module_name, sources=None, headers=None, source_dir=None, **build_kwargs):
    if module_name in _cached_plugins:
        return _cached_plugins[module_name]
    if module_name not in PLUGINS:
        raise RuntimeError(f'vfm_vae_b200.custom_ops: unknown plugin "{module_name}" (have: {sorted(PLUGINS)})')
    _lib.load()  # raises if libvfmops.so is missing -- no silent fallback
    _cached_plugins[module_name] = PLUGINS[module_name]
    return _cached_plugins[module_name]
