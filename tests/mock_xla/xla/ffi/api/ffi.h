// TEST INFRASTRUCTURE: a few-line stand-in for the surface of XLA's `xla/ffi/api/ffi.h` that jax_ffi/nifty_b200_jax.cc
// uses, so that the compile-gated branch of the binding is at least type-checked in an image without jaxlib
// (tests/test_jax_ffi_source.py).  It mirrors names and call shapes only; it is not XLA and is never shipped.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace xla { namespace ffi {

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string msg_;
};

template <class T> class Span {
 public:
  Span(const T* p, size_t n) : p_(p), n_(n) {}
  size_t size() const { return n_; }
  const T& operator[](size_t i) const { return p_[i]; }
  const T& back() const { return p_[n_ - 1]; }
 private:
  const T* p_; size_t n_;
};

class AnyBuffer {
 public:
  AnyBuffer(void* d, std::vector<int64_t> dims) : d_(d), dims_(std::move(dims)) {}
  void* untyped_data() const { return d_; }
  Span<int64_t> dimensions() const { return Span<int64_t>(dims_.data(), dims_.size()); }
 private:
  void* d_; std::vector<int64_t> dims_;
};
template <class T> class Result {
 public:
  explicit Result(T* t) : t_(t) {}
  T* operator->() const { return t_; }
 private:
  T* t_;
};
template <class S> struct PlatformStream {};

template <class... Stages> struct Binding {
  template <class T> Binding<Stages..., T> Ctx() const { return {}; }
  template <class T> Binding<Stages..., T> Attr(const char*) const { return {}; }
  template <class T> Binding<Stages..., T> Arg() const { return {}; }
  template <class T> Binding<Stages..., Result<T>> Ret() const { return {}; }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}}  // namespace xla::ffi

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                  \
  extern "C" XLA_FFI_Error* name(XLA_FFI_CallFrame*) { (void)&impl; (void)(binding); return nullptr; }
