// MINIMAL STAND-IN for xla/ffi/api/ffi.h (TEST INFRASTRUCTURE -- not XLA's header).
//
// The real header ships with jaxlib (jax.ffi.include_dir()); neither JAX nor XLA is installable in this image.  This
// file declares just the part of the C++ binding API that integration/xla_ffi_shim.cc uses -- ffi::Ffi::Bind() with
// Ctx<PlatformStream<cudaStream_t>> / Attr<T>(name) / Arg<Buffer<dtype>> / Ret<Buffer<dtype>>, ffi::Buffer,
// ffi::ResultBuffer, ffi::Error, XLA_FFI_DEFINE_HANDLER_SYMBOL -- with the same names and call shapes, over a plain-C
// call frame of its own (StubCallFrame below) instead of XLA's XLA_FFI_CallFrame, so that the shim can be compiled,
// its exported symbols checked and a call driven through the handler from a test.  Nothing here is ABI-compatible
// with XLA; the shim source is what is shared.
#ifndef MPXB_XLA_FFI_STUB_H_
#define MPXB_XLA_FFI_STUB_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

typedef struct CUstream_st* cudaStream_t;

extern "C" {
// one buffer of the stub call frame
typedef struct StubBuffer {
  void* data;
  int32_t dtype;       // xla::ffi::DataType
  int32_t rank;
  const int64_t* dims;
} StubBuffer;
typedef struct StubAttr {
  const char* name;
  int64_t value;
} StubAttr;
typedef struct StubCallFrame {
  void* stream;
  int32_t nargs, nrets, nattrs;
  const StubBuffer* args;
  const StubBuffer* rets;
  const StubAttr* attrs;
  char error[256];     // filled when the handler fails
} StubCallFrame;
typedef struct XLA_FFI_Error XLA_FFI_Error;   // opaque in the real API; here: null = success, else the frame
}

namespace xla {
namespace ffi {

enum DataType : int32_t { U8 = 0, S32 = 1, F32 = 2 };
enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };

template <DataType dt> struct NativeType;
template <> struct NativeType<U8> { using type = uint8_t; };
template <> struct NativeType<S32> { using type = int32_t; };
template <> struct NativeType<F32> { using type = float; };

class Error {
 public:
  Error() : code_(ErrorCode::kOk) {}
  Error(ErrorCode code, std::string message) : code_(code), message_(std::move(message)) {}
  static Error Success() { return Error(); }
  bool success() const { return code_ == ErrorCode::kOk; }
  const std::string& message() const { return message_; }

 private:
  ErrorCode code_;
  std::string message_;
};

class Dimensions {
 public:
  Dimensions(const int64_t* d, size_t n) : d_(d), n_(n) {}
  size_t size() const { return n_; }
  int64_t operator[](size_t i) const { return d_[i]; }

 private:
  const int64_t* d_;
  size_t n_;
};

template <DataType dt> class Buffer {
 public:
  using T = typename NativeType<dt>::type;
  Buffer() : b_{nullptr, dt, 0, nullptr} {}
  explicit Buffer(const StubBuffer& b) : b_(b) {}
  T* typed_data() const { return static_cast<T*>(b_.data); }
  Dimensions dimensions() const { return Dimensions(b_.dims, static_cast<size_t>(b_.rank)); }
  size_t element_count() const {
    size_t c = 1;
    for (int i = 0; i < b_.rank; ++i) c *= static_cast<size_t>(b_.dims[i]);
    return c;
  }
  size_t size_bytes() const { return element_count() * sizeof(T); }

 private:
  StubBuffer b_;
};

template <class T> class Result {
 public:
  Result() {}
  explicit Result(T v) : v_(v) {}
  T* operator->() { return &v_; }
  T& operator*() { return v_; }

 private:
  T v_;
};
template <DataType dt> using ResultBuffer = Result<Buffer<dt>>;

template <class T> struct PlatformStream {};

// ---- binding machinery: a Binding records, per handler parameter, how to decode it from the call frame
namespace internal {
template <class T> struct CtxTag {};
template <class T> struct ArgTag {};
template <class T> struct RetTag {};
template <class T> struct AttrTag { const char* name; };

struct Cursor { int arg = 0, ret = 0; };

template <class T> struct Decode;
template <class T> struct Decode<CtxTag<PlatformStream<T>>> {
  using type = T;
  static bool get(const CtxTag<PlatformStream<T>>&, StubCallFrame* f, Cursor&, type* out) { *out = static_cast<T>(f->stream); return true; }
};
template <DataType dt> struct Decode<ArgTag<Buffer<dt>>> {
  using type = Buffer<dt>;
  static bool get(const ArgTag<Buffer<dt>>&, StubCallFrame* f, Cursor& c, type* out) {
    if (c.arg >= f->nargs || f->args[c.arg].dtype != dt) return false;
    *out = Buffer<dt>(f->args[c.arg++]);
    return true;
  }
};
template <DataType dt> struct Decode<RetTag<Buffer<dt>>> {
  using type = Result<Buffer<dt>>;
  static bool get(const RetTag<Buffer<dt>>&, StubCallFrame* f, Cursor& c, type* out) {
    if (c.ret >= f->nrets || f->rets[c.ret].dtype != dt) return false;
    *out = Result<Buffer<dt>>(Buffer<dt>(f->rets[c.ret++]));
    return true;
  }
};
template <class T> struct Decode<AttrTag<T>> {
  using type = T;
  static bool get(const AttrTag<T>& t, StubCallFrame* f, Cursor&, type* out) {
    for (int i = 0; i < f->nattrs; ++i)
      if (std::strcmp(f->attrs[i].name, t.name) == 0) { *out = static_cast<T>(f->attrs[i].value); return true; }
    return false;
  }
};
}  // namespace internal

template <class... Tags> class Binding {
 public:
  explicit Binding(std::tuple<Tags...> tags) : tags_(std::move(tags)) {}
  template <class T> Binding<Tags..., internal::CtxTag<T>> Ctx() const { return add(internal::CtxTag<T>{}); }
  template <class T> Binding<Tags..., internal::ArgTag<T>> Arg() const { return add(internal::ArgTag<T>{}); }
  template <class T> Binding<Tags..., internal::RetTag<T>> Ret() const { return add(internal::RetTag<T>{}); }
  template <class T> Binding<Tags..., internal::AttrTag<T>> Attr(const char* name) const { return add(internal::AttrTag<T>{name}); }

  template <class Fn> XLA_FFI_Error* Call(Fn&& fn, StubCallFrame* frame) const {
    if (frame == nullptr) return reinterpret_cast<XLA_FFI_Error*>(1);
    return call(std::forward<Fn>(fn), frame, std::index_sequence_for<Tags...>{});
  }

 private:
  template <class T> Binding<Tags..., T> add(T t) const { return Binding<Tags..., T>(std::tuple_cat(tags_, std::make_tuple(t))); }
  template <class Fn, size_t... I> XLA_FFI_Error* call(Fn&& fn, StubCallFrame* f, std::index_sequence<I...>) const {
    std::tuple<typename internal::Decode<Tags>::type...> vals;
    internal::Cursor cur;
    bool ok = true;
    // decode in declaration order (arguments and results are positional)
    int dummy[] = {0, (ok = ok && internal::Decode<Tags>::get(std::get<I>(tags_), f, cur, &std::get<I>(vals)), 0)...};
    (void)dummy;
    if (!ok) { std::strncpy(f->error, "call frame does not match the binding", sizeof(f->error) - 1); return reinterpret_cast<XLA_FFI_Error*>(f); }
    Error e = fn(std::get<I>(vals)...);
    if (e.success()) return nullptr;
    std::strncpy(f->error, e.message().c_str(), sizeof(f->error) - 1);
    return reinterpret_cast<XLA_FFI_Error*>(f);
  }
  std::tuple<Tags...> tags_;
};

class Ffi {
 public:
  static Binding<> Bind() { return Binding<>(std::tuple<>()); }
};

}  // namespace ffi
}  // namespace xla

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(symbol, fn, binding)                      \
  extern "C" XLA_FFI_Error* symbol(StubCallFrame* call_frame) {                 \
    static const auto b = (binding);                                            \
    return b.Call(fn, call_frame);                                              \
  }

#endif  // MPXB_XLA_FFI_STUB_H_
