// oracle/_ref/libknn_ref.so -- the REFERENCE's own KNN / BallKNN kernel bodies, compiled from the
// sources where they lie (/root/reference/gridifyop/k_nn-inl.h:40-92, ball_k_nn-inl.h:43-95) through
// the parse-only stand-ins of this directory.  TEST INFRASTRUCTURE: used by tests/ to pin the C
// restatement (oracle/gridgcn_oracle.c) against the real reference for the two operators that have a
// CPU path (k_nn.cc:59, ball_k_nn.cc:59).  The row loop mirrors MXNet's CPU launcher
// mxnet_op::Kernel<OP, cpu>::Launch (an OpenMP `for` over i, k_nn-inl.h:109).
#include "k_nn-inl.h"
#include "ball_k_nn-inl.h"

extern "C" void ref_knn(const float *unknown, const float *known, const int *downnum, const int *upnum,
                        int B, int n, int m, int k, int *idx) {
#pragma omp parallel for
    for (int i = 0; i < B * n; i++)
        mxnet::op::KNNKernel::Map<float>(i, n, m, k, unknown, known, downnum, upnum, idx);
}

extern "C" void ref_ball_knn(const float *unknown, const float *known, const int *downnum,
                             const int *upnum, int B, int n, int m, int k, float radius, int *idx) {
#pragma omp parallel for
    for (int i = 0; i < B * n; i++)
        mxnet::op::BallKNNKernel::Map<float>(i, n, m, k, radius, unknown, known, downnum, upnum, idx);
}
