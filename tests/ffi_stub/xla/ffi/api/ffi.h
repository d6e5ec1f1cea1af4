// TEST STUB of the slice of xla/ffi/api/ffi.h that dgppo_ffi.cc uses: enough surface to syntax- and type-check the
// handlers (and their calls into the C ABI) in an image without jaxlib.  Not the real header; nothing links to it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace xla { namespace ffi {
enum DataType { F32, S32, U64, U8 };
enum class ErrorCode { kUnimplemented };
struct Error {
  Error() = default;
  Error(ErrorCode, std::string) {}
  static Error Success() { return {}; }
  static Error Internal(std::string) { return {}; }
  static Error InvalidArgument(std::string) { return {}; }
};
struct AnyBuffer { using Dimensions = std::vector<int64_t>; };
template <DataType> struct Native;
template <> struct Native<F32> { using type = float; };
template <> struct Native<S32> { using type = int32_t; };
template <> struct Native<U64> { using type = uint64_t; };
template <> struct Native<U8> { using type = uint8_t; };
template <DataType T> struct Buffer {
  typename Native<T>::type* typed_data() const { return nullptr; }
  AnyBuffer::Dimensions dimensions() const { return {}; }
  size_t element_count() const { return 0; }
  size_t size_bytes() const { return 0; }
};
template <DataType T> struct ResultBuffer {
  Buffer<T>* operator->() const { return nullptr; }
};
template <class S> struct PlatformStream {};
struct Binding {
  template <class T> Binding Ctx() const { return *this; }
  template <class T> Binding Arg() const { return *this; }
  template <class T> Binding Ret() const { return *this; }
  template <class T> Binding Attr(const char*) const { return *this; }
};
struct Ffi { static Binding Bind() { return {}; } };
}}  // namespace xla::ffi
#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding) \
  static auto* name##_impl_ref = &impl;                    \
  static auto name##_binding = binding
