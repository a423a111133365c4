"""The C-ABI shared library loads and exports exactly the symbols declared in include/pantea_b200.h.
No compute call is made (there is no GPU in the CPU test run)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared_symbols():
    text = (ROOT / "include" / "pantea_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pantea_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from pantea_b200 import _lib
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in pantea_b200.h but not exported"
    assert sorted(_lib.PROTOTYPES) == declared      # the Python binding covers the whole header, nothing else


def test_error_convention_without_gpu():
    import torch
    from pantea_b200 import _lib
    lib = _lib.load()
    assert lib.pantea_version().startswith(b"pantea_b200")
    assert lib.pantea_launch_count() >= 0
    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    assert lib.pantea_device_count() == 0
    handle = ctypes.c_void_p()
    code = lib.pantea_workspace_create(None, 64, 32, 64, ctypes.byref(handle))
    assert code == _lib.PANTEA_ECUDA and b"no CUDA device" in lib.pantea_last_error()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_cuda()
    assert lib.pantea_workspace_create(None, 64, 32, 16, ctypes.byref(handle)) == _lib.PANTEA_EINVAL
    with pytest.raises(ValueError):
        _lib.check(lib.pantea_workspace_set_owned_range(None, 0, 1))
    # process-wide switches are plain setters (no device needed) and return the previous value
    assert lib.pantea_set_fast_path(0) == 1 and lib.pantea_set_fast_path(1) == 0
    assert lib.pantea_set_gauss_screen(0.0) == 40.0 and lib.pantea_set_gauss_screen(40.0) == 0.0
    assert lib.pantea_mgpu_handle_bytes() == 64 and lib.pantea_mgpu_destroy(None) == 0
    # argument validation of the later entry points happens before any CUDA call
    counts = (ctypes.c_int64 * 2)()
    stats = ctypes.c_void_p(16)  # never dereferenced: the checks below fail first
    for call, needle in (
            (lambda: lib.pantea_workspace_set_skin(None, 0.5), b"NULL workspace"),
            (lambda: lib.pantea_neighbor_rebuilds(None, counts, None), b"NULL argument"),
            (lambda: lib.pantea_scaler_stats(None, 4, 3, 3, 64, stats, None), b"NULL argument"),
            (lambda: lib.pantea_scaler_stats(stats, 0, 3, 3, 64, stats, None), b"n_rows >= 1"),
            (lambda: lib.pantea_scaler_stats(stats, 4, 3, 2, 64, stats, None), b"n_cols <= ld"),
            (lambda: lib.pantea_lj_energy_forces(None, 1.0, 1.0, None, None, None, None), b"NULL workspace"),
            (lambda: lib.pantea_halo_pack(None, None, 4, None, None, 0, None, 0.5, None, 64, None), b"NULL array"),
            (lambda: lib.pantea_halo_pack(None, None, 0, None, None, 0, None, 0.5, None, 16, None), b"dtype"),
            (lambda: lib.pantea_halo_pack(stats, stats, 0, stats, stats, 4, None, 0.5, None, 64, None), b"needs positions and a flag"),
            (lambda: lib.pantea_halo_unpack_add(None, None, None, None, 4, 64, None), b"NULL array"),
            (lambda: lib.pantea_md_update_positions_mass(None, None, None, None, 0, 4, None, 0.25, 64, None), b"NULL argument"),
            (lambda: lib.pantea_md_update_velocities_mass(stats, stats, stats, None, 0, 4, 0.25, 7, None), b"dtype"),
            (lambda: lib.pantea_energy_forces(None, None, None, None, 1, None), b"no potential"),
            (lambda: lib.pantea_neighbor_build(None, None, None, 0, None, 1.0, None), b"NULL argument"),
            (lambda: lib.pantea_workspace_set_compute_precision(None, 32), b"NULL workspace"),
            (lambda: lib.pantea_mgpu_create(None, 0, 1, 10, None, None, None, None, None, 12.0, 0.25, 0, ctypes.byref(handle)), b"NULL argument"),
            (lambda: lib.pantea_mgpu_connect(None, None), b"NULL argument"),
            (lambda: lib.pantea_mgpu_set_state(None, None, None, None, None), b"NULL argument"),
            (lambda: lib.pantea_mgpu_run(None, 1, 1, None), b"NULL argument"),
            (lambda: lib.pantea_mgpu_read(None, None, None, None, None, None, None), b"NULL argument"),
            (lambda: lib.pantea_mgpu_export_handle(None, None), b"NULL argument")):
        code = call()
        assert code == _lib.PANTEA_EINVAL and needle in lib.pantea_last_error(), (code, lib.pantea_last_error())


def test_product_package_never_imports_the_oracle():
    offenders = [p for p in (ROOT / "pantea_b200").rglob("*.py") if re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), re.M)]
    assert offenders == []


def test_xla_ffi_shim_is_optional_and_says_why():
    """The XLA FFI binding (north star) wraps the same C ABI but needs jaxlib's headers: without jax the shim is not
    built and the JAX module refuses to import, both with the reason -- never a silent fallback."""
    import importlib.util
    shim = (ROOT / "pantea_b200" / "csrc" / "xla_ffi_shim.cc").read_text()
    for sym in ("pantea_neighbor_build", "pantea_energy_forces", "pantea_acsf_compute", "XLA_FFI_DEFINE_HANDLER_SYMBOL"):
        assert sym in shim
    if importlib.util.find_spec("jax") is not None:
        pytest.skip("jax present: the shim can be built (python -m pantea_b200.csrc.build --ffi)")
    from pantea_b200.csrc import build
    with pytest.raises(RuntimeError, match="jax.ffi is not importable"):
        build.build_ffi_shim()
    with pytest.raises(ImportError, match="ctypes \\+ torch"):
        importlib.import_module("pantea_b200.jax_ffi")
