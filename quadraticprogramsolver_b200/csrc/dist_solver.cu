// dist_solver.cu -- row-partitioned multi-GPU sparse path (placeholder until the NCCL path lands).
#include "host_common.h"

extern "C" {
int qpb200_dist_unique_id(void *) { return qpb::fail(QPB200_ERR_NCCL, "qpb200_dist_unique_id: not implemented in this build"); }
int qpb200_dist_create(qpb200_handle **out, int32_t, int32_t, const void *, int64_t, int64_t, const int64_t *, const int64_t *,
                       const double *, const int64_t *, const int64_t *, const double *, const double *, const double *,
                       const double *, const qpb200_settings *, int32_t) {
    if (out) *out = nullptr;
    return qpb::fail(QPB200_ERR_NCCL, "qpb200_dist_create: not implemented in this build");
}
int qpb200_dist_solve(qpb200_handle *, double *, double *, double *, qpb200_info *) {
    return qpb::fail(QPB200_ERR_NCCL, "qpb200_dist_solve: not implemented in this build");
}
}
