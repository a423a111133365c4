// XLA FFI custom-call targets over the C ABI of libpantea_b200.so (north star: "a JAX FFI custom-call, so arrays pass
// zero-copy").  Every handler forwards XLA's device buffers and XLA's own CUDA stream to the same extern "C" entry
// points the ctypes binding uses (include/pantea_b200.h); nothing is copied and nothing else is computed here.
//
// NOT part of the default build: it needs jaxlib's headers (`jax.ffi.include_dir()` -> xla/ffi/api/ffi.h), which this
// image does not have (no jax, no network).  `python -m pantea_b200.csrc.build --ffi` compiles it into
// pantea_b200/libpantea_b200_ffi.so when `import jax.ffi` works and says why it was skipped otherwise;
// pantea_b200/jax_ffi.py registers the targets.  Replaces, on the reference side, the jitted kernels
// `_jitted_compute_energy` / `_jitted_grad_compute_energy` (potentials/nnp/energy.py:66, force.py:18-21) and
// `_jitted_calculate_acsf_descriptor` / `_jitted_calculate_grad_acsf_descriptor` (descriptors/acsf/acsf.py:205-228).
//
// Host-side inputs of the C ABI (workspace handle, lattice diagonal, cutoff) travel as attributes; the workspace is
// created once per (potential, capacity) with pantea_workspace_create through ctypes and passed as an integer.
#include <cuda_runtime.h>

#include <cstdint>

#include "pantea_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

int32_t dtype_code(ffi::DataType t) {
    if (t == ffi::DataType::F64) return PANTEA_F64;
    if (t == ffi::DataType::F32) return PANTEA_F32;
    return 0;
}

ffi::Error bind_structure(pantea_workspace* ws, cudaStream_t stream, ffi::AnyBuffer positions, ffi::Buffer<ffi::S32> types,
                          int64_t has_box, double lx, double ly, double lz, double r_cutoff, int32_t* dtype_out) {
    const auto dims = positions.dimensions();
    if (dims.size() != 2 || dims[1] != 3) return ffi::Error::InvalidArgument("positions must be [n, 3]");
    if (static_cast<int64_t>(types.element_count()) != dims[0]) return ffi::Error::InvalidArgument("types must be [n]");
    *dtype_out = dtype_code(positions.element_type());
    if (*dtype_out == 0) return ffi::Error::InvalidArgument("positions must be float64 or float32");
    const double box[3] = {lx, ly, lz};
    if (pantea_neighbor_build(ws, positions.untyped_data(), types.typed_data(), dims[0], has_box ? box : nullptr, r_cutoff,
                              stream) != PANTEA_OK)
        return ffi::Error::Internal(pantea_last_error());
    return ffi::Error::Success();
}

// forces [n,3], e_atom [n], e_total [1]  <-  positions [n,3], types [n]
ffi::Error EnergyForcesImpl(cudaStream_t stream, int64_t workspace, double r_cutoff, int64_t has_box, double lx, double ly,
                            double lz, int64_t force_mode, ffi::AnyBuffer positions, ffi::Buffer<ffi::S32> types,
                            ffi::Result<ffi::AnyBuffer> forces, ffi::Result<ffi::AnyBuffer> e_atom,
                            ffi::Result<ffi::AnyBuffer> e_total) {
    auto* ws = reinterpret_cast<pantea_workspace*>(static_cast<intptr_t>(workspace));
    int32_t dtype = 0;
    ffi::Error err = bind_structure(ws, stream, positions, types, has_box, lx, ly, lz, r_cutoff, &dtype);
    if (err.failure()) return err;
    if (forces->element_type() != positions.element_type() || e_atom->element_type() != positions.element_type() ||
        e_total->element_type() != positions.element_type())
        return ffi::Error::InvalidArgument("outputs must have the dtype of positions");
    if (pantea_energy_forces(ws, e_atom->untyped_data(), forces->untyped_data(), e_total->untyped_data(),
                             static_cast<int32_t>(force_mode), stream) != PANTEA_OK)
        return ffi::Error::Internal(pantea_last_error());
    return ffi::Error::Success();  // row-capacity overflow is reported by pantea_neighbor_status (host side, after sync)
}

// G [n_c, n_sf], dG [n_c, n_sf, 3]  <-  positions [n,3], types [n], centres [n_c]
ffi::Error AcsfImpl(cudaStream_t stream, int64_t workspace, int64_t element, double r_cutoff, int64_t has_box, double lx,
                    double ly, double lz, ffi::AnyBuffer positions, ffi::Buffer<ffi::S32> types,
                    ffi::Buffer<ffi::S32> centres, ffi::Result<ffi::AnyBuffer> G, ffi::Result<ffi::AnyBuffer> dG) {
    auto* ws = reinterpret_cast<pantea_workspace*>(static_cast<intptr_t>(workspace));
    int32_t dtype = 0;
    ffi::Error err = bind_structure(ws, stream, positions, types, has_box, lx, ly, lz, r_cutoff, &dtype);
    if (err.failure()) return err;
    if (G->element_type() != positions.element_type() || dG->element_type() != positions.element_type())
        return ffi::Error::InvalidArgument("outputs must have the dtype of positions");
    const int64_t n_c = static_cast<int64_t>(centres.element_count());
    if (n_c == 0) return ffi::Error::Success();
    if (pantea_acsf_compute(ws, static_cast<int32_t>(element), centres.typed_data(), n_c, G->untyped_data(),
                            dG->untyped_data(), stream) != PANTEA_OK)
        return ffi::Error::Internal(pantea_last_error());
    return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(PanteaEnergyForces, EnergyForcesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("workspace")
                                  .Attr<double>("r_cutoff")
                                  .Attr<int64_t>("has_box")
                                  .Attr<double>("lx")
                                  .Attr<double>("ly")
                                  .Attr<double>("lz")
                                  .Attr<int64_t>("force_mode")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(PanteaAcsf, AcsfImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("workspace")
                                  .Attr<int64_t>("element")
                                  .Attr<double>("r_cutoff")
                                  .Attr<int64_t>("has_box")
                                  .Attr<double>("lx")
                                  .Attr<double>("ly")
                                  .Attr<double>("lz")
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::AnyBuffer>()
                                  .Ret<ffi::AnyBuffer>());
