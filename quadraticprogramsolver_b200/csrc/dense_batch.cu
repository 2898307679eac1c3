// dense_batch.cu -- batched small dense QPs (placeholder until the DMMA Cholesky path lands).
#include "host_common.h"

extern "C" {
int qpb200_batch_create(qpb200_batch **out, int64_t, int64_t, int64_t, const double *, const double *, const double *,
                        const double *, const double *, const qpb200_settings *) {
    if (out) *out = nullptr;
    return qpb::fail(QPB200_ERR_ARG, "qpb200_batch_create: not implemented in this build");
}
int qpb200_batch_solve(qpb200_batch *, double *, int32_t *, int64_t *, qpb200_info *) {
    return qpb::fail(QPB200_ERR_ARG, "qpb200_batch_solve: not implemented in this build");
}
void qpb200_batch_destroy(qpb200_batch *) {}
}
