// Minimal DECLARATION-ONLY stand-in for jaxlib's xla/ffi/api/ffi.h (jax.ffi.include_dir()), written for this repository:
// jaxlib is not installable in the build image (no network), so integration/pcx_xla_ffi.cc cannot be compiled against
// the real header here.  This file declares just the part of the public XLA FFI C++ API that the shim uses, with the
// same names and call shapes, so that `g++ -std=c++17 -fsyntax-only -Iintegration/stub -Iinclude integration/pcx_xla_ffi.cc`
// type-checks every handler and every binding (tests/test_integration_sources.py).  It is NOT the XLA header and must
// never be used to build the real shim: a maintainer compiles pcx_xla_ffi.cc with -I$(python -c "import jax.ffi as f;
// print(f.include_dir())") instead.
#ifndef PCX_STUB_XLA_FFI_API_FFI_H_
#define PCX_STUB_XLA_FFI_API_FFI_H_

#include <cstddef>
#include <cstdint>
#include <string>
#include <string_view>
#include <tuple>
#include <type_traits>
#include <utility>

struct XLA_FFI_CallFrame;
struct XLA_FFI_Error;

namespace xla::ffi {

enum class DataType { PRED, S8, S16, S32, S64, U8, U16, U32, U64, F16, F32, F64, BF16 };
inline constexpr DataType F64 = DataType::F64;
inline constexpr DataType F32 = DataType::F32;
inline constexpr DataType S32 = DataType::S32;
inline constexpr DataType S64 = DataType::S64;

template <DataType dtype> struct NativeTypeOf;
template <> struct NativeTypeOf<DataType::F64> { using type = double; };
template <> struct NativeTypeOf<DataType::F32> { using type = float; };
template <> struct NativeTypeOf<DataType::S32> { using type = int32_t; };
template <> struct NativeTypeOf<DataType::S64> { using type = int64_t; };

template <typename T>
class Span {
 public:
  const T* begin() const;
  const T* end() const;
  size_t size() const;
  const T& operator[](size_t i) const;
};

template <DataType dtype, size_t rank = static_cast<size_t>(-1)>
class Buffer {
 public:
  using native = typename NativeTypeOf<dtype>::type;
  native* typed_data() const;
  void* untyped_data() const;
  Span<int64_t> dimensions() const;
  size_t element_count() const;
  size_t size_bytes() const;
};

template <typename T>
class Result {
 public:
  T& operator*() const;
  T* operator->() const;
};
template <DataType dtype, size_t rank = static_cast<size_t>(-1)>
using ResultBuffer = Result<Buffer<dtype, rank>>;

enum class ErrorCode { kOk, kInvalidArgument, kInternal, kUnimplemented, kFailedPrecondition };
class Error {
 public:
  Error();
  Error(ErrorCode code, std::string message);
  static Error Success();
  static Error Internal(std::string message);
  static Error InvalidArgument(std::string message);
  bool success() const;
  bool failure() const;
};

template <typename T> struct PlatformStream {};

namespace internal {
template <typename T> struct ArgTag {};
template <typename T> struct RetTag {};
template <typename T> struct AttrTag {};
template <typename T> struct CtxTag {};
template <typename Tag> struct HandlerArg;
template <typename T> struct HandlerArg<ArgTag<T>> { using type = T; };
template <typename T> struct HandlerArg<RetTag<T>> { using type = Result<T>; };
template <typename T> struct HandlerArg<AttrTag<T>> { using type = T; };
template <typename T> struct HandlerArg<CtxTag<PlatformStream<T>>> { using type = T; };
}  // namespace internal

class HandlerBase {
 public:
  virtual ~HandlerBase() = default;
  XLA_FFI_Error* Call(const XLA_FFI_CallFrame* call_frame) const;
};
template <typename Fn, typename... Tags>
class Handler : public HandlerBase {};

template <typename... Tags>
class Binding {
 public:
  template <typename T> Binding<Tags..., internal::CtxTag<T>> Ctx() &&;
  template <typename T> Binding<Tags..., internal::ArgTag<T>> Arg() &&;
  template <typename T> Binding<Tags..., internal::RetTag<T>> Ret() &&;
  template <typename T> Binding<Tags..., internal::AttrTag<T>> Attr(std::string_view name) &&;
  // the callable must accept exactly the decoded (context, attribute, argument, result) types in binding order
  template <typename Fn>
  Handler<Fn, Tags...>* To(Fn fn) && {
    static_assert(std::is_invocable_r_v<Error, Fn, typename internal::HandlerArg<Tags>::type...>,
                  "FFI handler signature does not match its binding");
    return nullptr;
  }
};

class Ffi {
 public:
  static Binding<> Bind();
};

}  // namespace xla::ffi

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(fn, impl, binding)                                   \
  extern "C" XLA_FFI_Error* fn(XLA_FFI_CallFrame* call_frame) {                            \
    static auto* handler = (binding).To(impl);                                             \
    return handler->Call(call_frame);                                                      \
  }

#endif  // PCX_STUB_XLA_FFI_API_FFI_H_
