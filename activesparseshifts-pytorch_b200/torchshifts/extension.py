"""Loader of the native library and registration of the ``torch.ops.torchshifts`` operators.

Mirror of the reference's ``torchshifts/extension.py`` (``_HAS_OPS``, ``_assert_has_ops``,
``_check_cuda_version``), except that what gets loaded is the sm_100a C-ABI library
``libtorchshifts_b200.so`` (ctypes) and the operators are defined from Python with
``torch.library`` (see ``_ops.py``) instead of a torch C++ extension.
"""
_HAS_OPS = False
error_str = ''
_NATIVE = None


def _has_ops():
    return False


def _register_extensions():
    global _NATIVE
    from ._cabi import NativeLibrary
    _NATIVE = NativeLibrary()
    from . import _ops
    _ops.register(_NATIVE)


try:
    _register_extensions()
    _HAS_OPS = True

    def _has_ops():  # noqa: F811
        return True
except (ImportError, OSError) as e:
    error_str = str(e)


def native():
    """The loaded :class:`torchshifts._cabi.NativeLibrary` (raises when it is missing)."""
    _assert_has_ops()
    return _NATIVE


def _assert_has_ops():
    if not _has_ops():
        raise RuntimeError(
            "Couldn't load the torchshifts-b200 native library (libtorchshifts_b200.so). There is no "
            "CPU or eager fallback: build it for sm_100a with `python __graft_entry__.py` (runs nvcc "
            "-gencode arch=compute_100a,code=sm_100a) and make sure the CUDA 12.8+ runtime is present."
            f"\n\nImport error details:\n\t{error_str}"
        )


def _check_cuda_version():
    """CUDA (runtime API) version the native library was built with, or -1 without the library.

    Like the reference, raise when torch and the library were built against different CUDA major
    versions.  (The reference also rejects a newer torch minor version; that is not needed here:
    the CUDA runtime is linked statically and no torch type crosses the C ABI.)
    """
    if not _HAS_OPS:
        return -1
    import torch
    _version = torch.ops.torchshifts._cuda_version()
    if _version != -1 and torch.version.cuda is not None:
        ts_major, ts_minor = _version // 1000, (_version % 1000) // 10
        t_major, t_minor = (int(v) for v in torch.version.cuda.split('.')[:2])
        if t_major != ts_major:
            raise RuntimeError("Detected that PyTorch and torchshifts were compiled with different CUDA versions. "
                               "PyTorch has CUDA Version={}.{} and torchshifts has CUDA Version={}.{}. "
                               "Please reinstall the torchshifts that matches your PyTorch install."
                               .format(t_major, t_minor, ts_major, ts_minor))
    return _version


_check_cuda_version()
