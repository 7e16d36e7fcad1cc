// A LOCAL STAND-IN for JAX's xla/ffi/api/ffi.h -- test infrastructure, not the real header and not shipped.
//
// JAX is not installable in this image, so probdiffeq_b200/csrc/ffi/pdeq_xla_ffi.cc can never be built against the real
// XLA FFI here. This header models the handful of xla::ffi names the shim uses (Buffer, ResultBuffer, Error,
// PlatformStream, Ffi::Bind().Ctx/Attr/Arg/Ret and XLA_FFI_DEFINE_HANDLER_SYMBOL) closely enough that the compiler
// checks what can rot: the shim's use of the C ABI in include/probdiffeq_b200.h (argument order and types of every
// pdeq_* call, struct field names) and that each handler's C++ signature is exactly what its binding declares, in
// order (a static_assert inside the macro). tests/test_ffi_shim_compiles.py runs `g++ -fsyntax-only` against it.
#pragma once

#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <type_traits>

namespace xla::ffi {

enum DataType { F64, S32, U8 };
template <DataType>
struct NativeOf;
template <>
struct NativeOf<F64> { using type = double; };
template <>
struct NativeOf<S32> { using type = int32_t; };
template <>
struct NativeOf<U8> { using type = uint8_t; };

struct Dims {
  const int64_t* ptr;
  size_t len;
  size_t size() const { return len; }
  int64_t operator[](size_t i) const { return ptr[i]; }
  int64_t back() const { return ptr[len - 1]; }
};

template <DataType dt>
struct Buffer {
  typename NativeOf<dt>::type* typed_data() const;
  void* untyped_data() const;
  Dims dimensions() const;
  size_t element_count() const;
  size_t size_bytes() const;
};

template <class T>
struct Result {
  T* operator->() const;
  T& operator*() const;
};
template <DataType dt>
using ResultBuffer = Result<Buffer<dt>>;

enum class ErrorCode { kInternal, kInvalidArgument };
struct Error {
  Error();
  Error(ErrorCode, std::string);
  static Error Success();
};

template <class T>
struct PlatformStream {
  using type = T;
};

template <class... Ts>
struct Binding {
  template <class C>
  Binding<Ts..., typename C::type> Ctx() const;
  template <class A>
  Binding<Ts..., A> Attr(const char*) const;
  template <class A>
  Binding<Ts..., A> Arg() const;
  template <class R>
  Binding<Ts..., Result<R>> Ret() const;
  // the handler must take exactly the bound types, in binding order
  template <class F>
  static constexpr bool matches = std::is_same_v<F, Error (*)(Ts...)>;
};

struct Ffi {
  static Binding<> Bind();
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                                      \
  static_assert(std::remove_reference_t<decltype(binding)>::template matches<decltype(&fn)>,      \
                #fn ": handler signature does not match its binding");                           \
  extern "C" void* symbol() { return reinterpret_cast<void*>(&fn); }
