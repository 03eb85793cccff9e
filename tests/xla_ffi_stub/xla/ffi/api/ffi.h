// Minimal stand-in for XLA's typed-FFI header (xla/ffi/api/ffi.h) -- TEST INFRASTRUCTURE ONLY.
//
// JAX / XLA headers are not available in this image, so ofdft_nflows_b200/jax_ffi/xla_ffi_shim.cc cannot be compiled
// against the real API here.  This stub declares the small subset of the API the shim uses with the same names and
// shapes (Buffer<dtype>, ResultBuffer, Span, Error, Ffi::Bind().Ctx/Arg/Ret/Attr, XLA_FFI_DEFINE_HANDLER_SYMBOL) and --
// unlike a bag of empty templates -- CHECKS at compile time that every handler is callable with exactly the argument
// list its binding declares.  tests/test_capi.py compiles the shim against it (g++ -fsyntax-only), which catches drift
// between the shim, its bindings and include/ofdft_b200.h.  It does not execute anything.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <type_traits>
#include <utility>

namespace xla {
namespace ffi {

enum DataType { U8, S32, F32, F64 };
template <DataType> struct NativeType;
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<F32> { using type = float; };
template <> struct NativeType<F64> { using type = double; };

template <typename T>
class Span {
 public:
  const T* begin() const { return data_; }
  const T* end() const { return data_ + size_; }
  size_t size() const { return size_; }
  const T& operator[](size_t i) const { return data_[i]; }
 private:
  const T* data_ = nullptr;
  size_t size_ = 0;
};

template <DataType dtype>
class Buffer {
 public:
  using T = typename NativeType<dtype>::type;
  T* typed_data() const { return data_; }
  Span<int64_t> dimensions() const { return dims_; }
  size_t element_count() const { return count_; }
 private:
  T* data_ = nullptr;
  Span<int64_t> dims_;
  size_t count_ = 0;
};

template <typename T>
class Result {
 public:
  T* operator->() { return &value_; }
  T& operator*() { return value_; }
 private:
  T value_;
};
template <DataType dtype> using ResultBuffer = Result<Buffer<dtype>>;

enum class ErrorCode { kOk, kInternal, kInvalidArgument };
class Error {
 public:
  Error() = default;
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
 private:
  ErrorCode code_ = ErrorCode::kOk;
  std::string message_;
};

template <typename Stream> struct PlatformStream { using type = Stream; };

template <typename... Ts>
struct Binding {
  template <typename C> Binding<Ts..., typename C::type> Ctx() const { return {}; }
  template <typename A> Binding<Ts..., A> Arg() const { return {}; }
  template <typename R> Binding<Ts..., Result<R>> Ret() const { return {}; }
  template <typename A> Binding<Ts..., A> Attr(const char*) const { return {}; }
  template <typename Fn>
  static constexpr bool Matches() { return std::is_invocable_r<Error, Fn, Ts...>::value; }
};
struct Ffi { static Binding<> Bind() { return {}; } };

}  // namespace ffi
}  // namespace xla

// the real macro defines an XLA_FFI_Handler; here it only forces the signature check and keeps the symbol name
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                                              \
  static_assert(decltype(binding)::template Matches<decltype(&impl)>(),                                                 \
                #impl " is not callable with the argument list its binding declares");                                  \
  extern "C" void* name(void*) { return reinterpret_cast<void*>(&impl); }
