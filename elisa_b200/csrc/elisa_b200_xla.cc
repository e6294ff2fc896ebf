// XLA FFI shim: registers the C ABI of include/elisa_b200.h as jax.ffi custom calls.
//
// Built by elisa_b200/build.py ONLY when the XLA FFI headers are present
// (`python -c "import jax; print(jax.ffi.include_dir())"` -> .../xla/ffi/api/ffi.h); the image this
// library is developed in has neither jax nor those headers, so this translation unit is
// committed as source, kept trivially thin (every handler forwards to one C function), and
// exercised through the same C functions from torch / ctypes (elisa_b200/autograd.py implements
// the custom-JVP contract of INTEGRATION.md section 4 with torch.autograd).
//
// Reference seam: infer/helper.py:313-323 (likelihood factories) inside numpyro_model
// (:332-367), batched by jax.vmap (models/model.py:406-411).  Handlers:
//   elisa_b200_loglike        (theta[N,P])                 -> loglike[N,S]
//   elisa_b200_loglike_grad   (theta[N,P])                 -> loglike[N,S], jac[N,S,P]
//   elisa_b200_loglike_grad_c (theta, counts_on, counts_off [N, sum C]) -> loglike, jac
//   elisa_b200_sites          (theta[N,P]) attr spectrum    -> rate, model_on, ll_on, model_off, ll_off [N,C_s]
// Attribute "plan": the ElisaPlan* as an int64 (created through ctypes on the Python side;
// plans are per-device -- one per jax device -- and own every replicated constant).
// Batch axes: the Python wrapper calls jax.ffi.ffi_call(..., vmap_method="broadcast_all"), so a
// vmapped call arrives with extra LEADING dimensions on theta; all leading dimensions are
// flattened into N here.
#include <cstdint>

#include "xla/ffi/api/ffi.h"

#include "elisa_b200.h"

namespace ffi = xla::ffi;

namespace {

// product of all dimensions but the last `keep`
template <typename Buf>
int64_t leading(const Buf& b, int keep) {
    const auto dims = b.dimensions();
    int64_t n = 1;
    for (size_t i = 0; i + keep < dims.size(); ++i) n *= dims[i];
    return n;
}

ffi::Error status(int rc) {
    if (rc == ELISA_OK) return ffi::Error::Success();
    const auto code = rc == ELISA_ERR_INVALID ? ffi::ErrorCode::kInvalidArgument
                      : rc == ELISA_ERR_NOMEM ? ffi::ErrorCode::kResourceExhausted
                      : rc == ELISA_ERR_UNSUPPORTED ? ffi::ErrorCode::kUnimplemented
                                                    : ffi::ErrorCode::kInternal;
    return ffi::Error(code, elisa_b200_last_error());
}

ffi::Error LoglikeImpl(void* stream, int64_t plan, ffi::Buffer<ffi::F64> theta, ffi::ResultBuffer<ffi::F64> ll) {
    return status(elisa_b200_loglike_dev(reinterpret_cast<ElisaPlan*>(plan), theta.typed_data(), nullptr, nullptr,
                                         leading(theta, 1), ll->typed_data(), stream));
}

ffi::Error LoglikeGradImpl(void* stream, int64_t plan, ffi::Buffer<ffi::F64> theta, ffi::ResultBuffer<ffi::F64> ll,
                           ffi::ResultBuffer<ffi::F64> jac) {
    return status(elisa_b200_loglike_grad_dev(reinterpret_cast<ElisaPlan*>(plan), theta.typed_data(), nullptr, nullptr,
                                              leading(theta, 1), ll->typed_data(), jac->typed_data(), stream));
}

// per-replica observed data: what substituting '{name}_Non_data' / '{name}_Noff_data' does
// upstream (infer/helper.py:608-613); a zero-sized counts_off buffer means "not given"
ffi::Error LoglikeGradCountsImpl(void* stream, int64_t plan, ffi::Buffer<ffi::F64> theta,
                                 ffi::Buffer<ffi::F64> counts_on, ffi::Buffer<ffi::F64> counts_off,
                                 ffi::ResultBuffer<ffi::F64> ll, ffi::ResultBuffer<ffi::F64> jac) {
    const double* on = counts_on.element_count() ? counts_on.typed_data() : nullptr;
    const double* off = counts_off.element_count() ? counts_off.typed_data() : nullptr;
    return status(elisa_b200_loglike_grad_dev(reinterpret_cast<ElisaPlan*>(plan), theta.typed_data(), on, off,
                                              leading(theta, 1), ll->typed_data(), jac->typed_data(), stream));
}

ffi::Error SitesImpl(void* stream, int64_t plan, int32_t spectrum, ffi::Buffer<ffi::F64> theta,
                     ffi::ResultBuffer<ffi::F64> rate, ffi::ResultBuffer<ffi::F64> model_on,
                     ffi::ResultBuffer<ffi::F64> ll_on, ffi::ResultBuffer<ffi::F64> model_off,
                     ffi::ResultBuffer<ffi::F64> ll_off) {
    return status(elisa_b200_sites_dev(reinterpret_cast<ElisaPlan*>(plan), spectrum, theta.typed_data(), nullptr,
                                       nullptr, leading(theta, 1), rate->typed_data(), model_on->typed_data(),
                                       ll_on->typed_data(), model_off->typed_data(), ll_off->typed_data(), stream));
}

}  // namespace

// PlatformStream<void*>: the CUstream / cudaStream_t XLA runs the custom call on; the library only
// enqueues on it (no allocation, no host synchronisation -- include/elisa_b200.h conventions),
// which is what XLA requires of a command-buffer compatible custom call.
XLA_FFI_DEFINE_HANDLER_SYMBOL(ElisaB200Loglike, LoglikeImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<void*>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(ElisaB200LoglikeGrad, LoglikeGradImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<void*>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(ElisaB200LoglikeGradCounts, LoglikeGradCountsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<void*>>()
                                  .Attr<int64_t>("plan")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());

XLA_FFI_DEFINE_HANDLER_SYMBOL(ElisaB200Sites, SitesImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<void*>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int32_t>("spectrum")
                                  .Arg<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>()
                                  .Ret<ffi::Buffer<ffi::F64>>());
