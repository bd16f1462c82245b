// MOCK of the subset of jaxlib's xla/ffi/api/ffi.h that jubik_b200/csrc/jax_ffi.cc uses -- TEST INFRASTRUCTURE ONLY.
// jax is not installable in the build image, so the real header is absent; this stand-in lets the CPU test-suite check
// that the handler file is well-formed C++ against the documented shape of the API (Buffer / ResultBuffer / Error /
// Ffi::Bind().Ctx().Attr().Arg().Ret() / XLA_FFI_DEFINE_HANDLER_SYMBOL).  The product build uses
// `jax.ffi.include_dir()` and never sees this file.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <tuple>
#include <utility>

namespace xla::ffi {
enum class ErrorCode { kOk = 0, kInvalidArgument = 3, kInternal = 13 };
class Error {
 public:
    Error() = default;
    Error(ErrorCode c, std::string m) : code_(c), msg_(std::move(m)) {}
    static Error Success() { return Error(); }
    bool failure() const { return code_ != ErrorCode::kOk; }
    bool success() const { return code_ == ErrorCode::kOk; }
    const std::string& message() const { return msg_; }
 private:
    ErrorCode code_ = ErrorCode::kOk;
    std::string msg_;
};
enum DataType { F64 = 12 };
template <DataType> struct NativeOf;
template <> struct NativeOf<F64> { using type = double; };
template <DataType T>
class Buffer {
 public:
    using native = typename NativeOf<T>::type;
    native* typed_data() const { return data_; }
    size_t element_count() const { return count_; }
    native* data_ = nullptr;
    size_t count_ = 0;
};
template <DataType T>
class ResultBuffer {
 public:
    Buffer<T>* operator->() { return &b_; }
    Buffer<T> b_;
};
template <class T> struct PlatformStream {};
template <class... Ts>
struct Binding {
    template <class T> Binding<Ts..., T> Ctx() const { return {}; }
    template <class T> Binding<Ts..., T> Attr(const char*) const { return {}; }
    template <class T> Binding<Ts..., T> Arg() const { return {}; }
    template <class T> Binding<Ts..., T> Ret() const { return {}; }
};
struct Ffi { static Binding<> Bind() { return {}; } };
// signature check: the implementation must be callable with what the binding declares
template <class T> struct ArgOf { using type = T; };
template <class T> struct ArgOf<PlatformStream<T>> { using type = T; };
template <class F, class... Ts>
constexpr bool Matches(F f, Binding<Ts...>) {
    return std::is_invocable_r_v<Error, F, typename ArgOf<Ts>::type...>;
}
}  // namespace xla::ffi

// results are declared with Ret<Buffer<T>> and received as ResultBuffer<T>: map them for the check
namespace xla::ffi::mock {
template <class F> struct Sig;
template <class... A> struct Sig<Error (*)(A...)> { static constexpr size_t n = sizeof...(A); };
}  // namespace xla::ffi::mock

#define XLA_FFI_DEFINE_HANDLER_SYMBOL(name, impl, binding)                                   \
    static_assert(xla::ffi::mock::Sig<decltype(&impl)>::n >= 2, #impl " takes stream + plan"); \
    extern "C" void* name(void*) { (void)(binding); (void)&impl; return nullptr; }
