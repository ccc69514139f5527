// XLA custom-call (jax.ffi) handlers in front of the C ABI of include/nifty_b200.h.
//
// This is the binding a nifty.re maintainer adds: it turns the two bilinear operators of the correlated-field model
//     L(a, x)   = (1/V) hartley(a[power_distributor] * x)                       (nifty/re/correlated_field.py:909-912, :882-887)
//     L^T(a, c) = (a[pd] * g ,  segment_sum_pd(x * g)),  g = (1/V) hartley(c)    (what jax.linear_transpose derives, likelihood.py:619)
// into XLA FFI targets on the CUDA platform.  jax_ffi/_b200.py registers them and builds the JAX-transformable function
// (custom_jvp whose tangent is LINEAR FFI calls with registered transposes, vmap through the batched entry points).
//
// Build (needs the XLA FFI headers that ship with jaxlib; they are not in this repository's image, so the file is
// compile-gated and `nb200_jax_ffi_available()` reports which branch was built):
//     g++ -O2 -std=c++17 -shared -fPIC -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") -Iinclude \
//         jax_ffi/nifty_b200_jax.cc -Lnifty_b200/lib -lniftyb200 -o jax_ffi/libnifty_b200_jax.so
//
// Contract with XLA: buffers are caller-owned device memory, outputs pre-allocated by XLA, nothing is
// retained after return, work is enqueued on the stream XLA passes, no synchronisation, errors are returned (never thrown);
// handlers may run concurrently from one host thread per device (the library keeps no unsynchronised globals).
#include <cstdint>

#include "../include/nifty_b200.h"

#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define NB200_HAVE_XLA_FFI 1
#endif
#endif

#ifdef NB200_HAVE_XLA_FFI
#include <cuda_runtime_api.h>

#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

// The plan handle travels as an int64 attribute (the address of the nb200_plan created by the Python side and kept alive
// there for the lifetime of the jitted function).
inline nb200_plan* plan_of(int64_t handle) { return reinterpret_cast<nb200_plan*>(static_cast<intptr_t>(handle)); }

inline ffi::Error status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(ffi::ErrorCode::kInternal, nb200_last_error());
}

// leading batch extent of `buf` given the rank of one item (vmap_method="broadcast_all" prepends batch axes)
template <class Buf> inline int64_t batch_of(const Buf& buf, size_t item_rank) {
  int64_t b = 1;
  const auto dims = buf.dimensions();
  for (size_t i = 0; i + item_rank < dims.size(); ++i) b *= dims[i];
  return b;
}

// out = offset + L(amp, xi);  amp: [..., K] or [K] (shared by the batch), xi / out: [..., *grid]
ffi::Error CfApplyImpl(cudaStream_t stream, int64_t plan, int64_t grid_rank, double offset, ffi::AnyBuffer amp, ffi::AnyBuffer xi,
                       ffi::Result<ffi::AnyBuffer> out) {
  const int64_t batch = batch_of(xi, (size_t)grid_rank);
  const int64_t amp_batch = batch_of(amp, 1);
  if (amp_batch != 1 && amp_batch != batch) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "nb200_cf_apply: amp batch mismatch");
  const int64_t K = amp.dimensions().back();
  return status(nb200_cf_apply_batch(plan_of(plan), stream, amp.untyped_data(), amp_batch == 1 ? 0 : K, xi.untyped_data(), offset,
                                     out->untyped_data(), batch));
}

// (xi_bar, amp_bar) = L^T(amp, xi; cot)
ffi::Error CfAdjointImpl(cudaStream_t stream, int64_t plan, int64_t grid_rank, ffi::AnyBuffer amp, ffi::AnyBuffer xi, ffi::AnyBuffer cot,
                         ffi::Result<ffi::AnyBuffer> xi_bar, ffi::Result<ffi::AnyBuffer> amp_bar) {
  const int64_t batch = batch_of(cot, (size_t)grid_rank);
  const int64_t amp_batch = batch_of(amp, 1);
  if (amp_batch != 1 && amp_batch != batch) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "nb200_cf_adjoint: amp batch mismatch");
  const int64_t K = amp.dimensions().back();
  return status(nb200_cf_apply_adjoint_batch(plan_of(plan), stream, amp.untyped_data(), amp_batch == 1 ? 0 : K, xi.untyped_data(),
                                             cot.untyped_data(), xi_bar->untyped_data(), amp_bar->untyped_data(), K, batch));
}

// xi_bar only (the transpose of x -> L(a, x); no excitations needed)
ffi::Error CfAdjointXiImpl(cudaStream_t stream, int64_t plan, int64_t grid_rank, ffi::AnyBuffer amp, ffi::AnyBuffer cot,
                           ffi::Result<ffi::AnyBuffer> xi_bar) {
  const int64_t batch = batch_of(cot, (size_t)grid_rank);
  const int64_t amp_batch = batch_of(amp, 1);
  if (amp_batch != 1 && amp_batch != batch) return ffi::Error(ffi::ErrorCode::kInvalidArgument, "nb200_cf_adjoint_xi: amp batch mismatch");
  const int64_t K = amp.dimensions().back();
  return status(nb200_cf_apply_adjoint_batch(plan_of(plan), stream, amp.untyped_data(), amp_batch == 1 ? 0 : K, nullptr, cot.untyped_data(),
                                             xi_bar->untyped_data(), nullptr, K, batch));
}

// hartley(p) -- the module attribute bound at finalize() (nifty/re/correlated_field.py:24-30, :865).  Linear and self-adjoint: the
// JVP rule and the transpose rule of the Python side are this same target.  Batch axes (vmap) are looped: one transform per item.
ffi::Error HartleyImpl(cudaStream_t stream, int64_t plan, int64_t grid_rank, int64_t item_bytes, ffi::AnyBuffer in,
                       ffi::Result<ffi::AnyBuffer> out) {
  const int64_t batch = batch_of(in, (size_t)grid_rank);
  const char* src = static_cast<const char*>(in.untyped_data());
  char* dst = static_cast<char*>(out->untyped_data());
  for (int64_t b = 0; b < batch; ++b) {
    int rc = nb200_hartley(plan_of(plan), stream, src + b * item_bytes, dst + b * item_bytes);
    if (rc != 0) return status(rc);
  }
  return ffi::Error::Success();
}

// the same seam on a grid whose extents are not powers of two: chirp-z around the padded plan (nb200_hartley_chirpz).  `tab` and
// `work` are device buffers the Python side allocates once (tables) / per call (scratch, an extra XLA result that is dropped).
ffi::Error HartleyChirpzImpl(cudaStream_t stream, int64_t plan, int64_t grid_rank, int64_t item_bytes, int64_t n0, int64_t n1, int64_t n2,
                             ffi::AnyBuffer tab, ffi::AnyBuffer in, ffi::Result<ffi::AnyBuffer> out, ffi::Result<ffi::AnyBuffer> work) {
  const int64_t batch = batch_of(in, (size_t)grid_rank);
  const int64_t n[3] = {n0, n1, n2};
  const char* src = static_cast<const char*>(in.untyped_data());
  char* dst = static_cast<char*>(out->untyped_data());
  for (int64_t b = 0; b < batch; ++b) {
    int rc = nb200_hartley_chirpz(plan_of(plan), stream, n, tab.untyped_data(), src + b * item_bytes, work->untyped_data(), dst + b * item_bytes);
    if (rc != 0) return status(rc);
  }
  return ffi::Error::Success();
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(nb200_jax_cf_apply, CfApplyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("grid_rank")
                                  .Attr<double>("offset")
                                  .Arg<ffi::AnyBuffer>()    // amp
                                  .Arg<ffi::AnyBuffer>()    // xi
                                  .Ret<ffi::AnyBuffer>());  // out

XLA_FFI_DEFINE_HANDLER_SYMBOL(nb200_jax_cf_adjoint, CfAdjointImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("grid_rank")
                                  .Arg<ffi::AnyBuffer>()    // amp
                                  .Arg<ffi::AnyBuffer>()    // xi
                                  .Arg<ffi::AnyBuffer>()    // cot
                                  .Ret<ffi::AnyBuffer>()    // xi_bar
                                  .Ret<ffi::AnyBuffer>());  // amp_bar

XLA_FFI_DEFINE_HANDLER_SYMBOL(nb200_jax_cf_adjoint_xi, CfAdjointXiImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("grid_rank")
                                  .Arg<ffi::AnyBuffer>()    // amp
                                  .Arg<ffi::AnyBuffer>()    // cot
                                  .Ret<ffi::AnyBuffer>());  // xi_bar

XLA_FFI_DEFINE_HANDLER_SYMBOL(nb200_jax_hartley, HartleyImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("grid_rank")
                                  .Attr<int64_t>("item_bytes")
                                  .Arg<ffi::AnyBuffer>()    // in
                                  .Ret<ffi::AnyBuffer>());  // out

XLA_FFI_DEFINE_HANDLER_SYMBOL(nb200_jax_hartley_chirpz, HartleyChirpzImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Attr<int64_t>("plan")
                                  .Attr<int64_t>("grid_rank")
                                  .Attr<int64_t>("item_bytes")
                                  .Attr<int64_t>("n0")
                                  .Attr<int64_t>("n1")
                                  .Attr<int64_t>("n2")
                                  .Arg<ffi::AnyBuffer>()    // tab
                                  .Arg<ffi::AnyBuffer>()    // in
                                  .Ret<ffi::AnyBuffer>()    // out
                                  .Ret<ffi::AnyBuffer>());  // work (scratch)

extern "C" int nb200_jax_ffi_available() { return 1; }

#else  // no XLA FFI headers: the translation unit still builds and says so

extern "C" int nb200_jax_ffi_available() { return 0; }

#endif
