/* Minimal stand-ins for the MXNet / mshadow / dmlc / nnvm declarations that
 * /root/reference/gridifyop/k_nn-inl.h and ball_k_nn-inl.h mention, just enough for those two
 * headers to PARSE with a plain g++.  TEST INFRASTRUCTURE (oracle/_ref build): no MXNet code is
 * reproduced here, only empty types and no-op macros; the kernel bodies that get compiled and
 * executed (KNNKernel::Map, BallKNNKernel::Map) are the reference's own, read from
 * /root/reference at build time and never copied into this repository. */
#ifndef GRIDGCN_REF_SHIM_COMMON_H_
#define GRIDGCN_REF_SHIM_COMMON_H_
#include <cfloat>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#define MSHADOW_XINLINE inline
#define CHECK_EQ(a, b) ((void)0)
#define MSHADOW_TYPE_SWITCH(flag, DType, ...) \
    {                                           \
        typedef float DType;                    \
        { __VA_ARGS__ }                         \
    }
#define DMLC_DECLARE_PARAMETER(T) void __declare__()
#define DMLC_DECLARE_FIELD(f) ::dmlc::FieldStub()

namespace dmlc {
struct FieldStub {
    template <typename T> FieldStub &set_default(const T &) { return *this; }
    FieldStub &describe(const char *) { return *this; }
};
template <typename T> struct Parameter {};
}  // namespace dmlc

namespace mshadow {
struct cpu {};
struct gpu {};
template <typename xpu> struct Stream {};
}  // namespace mshadow

namespace nnvm {
struct NodeAttrs { int parsed; };
template <typename T> const T &get(const int &) { static T t; return t; }
}  // namespace nnvm

namespace mxnet {
typedef int index_t;
struct TShape {};
enum OpReqType { kNullOp };
struct TBlob {
    int type_flag_;
    int size(int) const { return 0; }
    template <typename T> T *dptr() const { return nullptr; }
};
struct OpContext {
    template <typename xpu> mshadow::Stream<xpu> *get_stream() const { return nullptr; }
};
namespace op {
namespace mxnet_op {
template <typename OP, typename xpu> struct Kernel {
    template <typename... Args> static void Launch(mshadow::Stream<xpu> *, int, Args...) {}
};
}  // namespace mxnet_op
}  // namespace op
}  // namespace mxnet
#endif
