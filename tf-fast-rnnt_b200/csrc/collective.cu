// reduction='sum'|'mean' over a batch sharded by utterance: the one collective on the path, a sum
// all-reduce of a handful of floats over NVLink through NCCL.  NCCL is bound at run time: the TensorFlow shim's
// process (or a torch one) already carries its own libnccl.so.2, and dlopen by soname returns that instance.
#include <dlfcn.h>

#include "common.cuh"
#include "launchers.h"

namespace frn {
namespace {
// ncclResult_t ncclAllReduce(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t)
using AllReduceFn = int (*)(const void *, void *, size_t, int, int, void *, cudaStream_t);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;
AllReduceFn bind_allreduce() {
  static const AllReduceFn fn = [] {
    void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
    if (!sym) {
      if (void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL)) sym = dlsym(h, "ncclAllReduce");
    }
    return reinterpret_cast<AllReduceFn>(sym);
  }();
  return fn;
}
}  // namespace

int launch_allreduce_sum(float *buf, size_t n, void *comm, cudaStream_t stream) {
  const AllReduceFn fn = bind_allreduce();
  if (!fn) return FRN_EUNSUPPORTED;
  return fn(buf, buf, n, kNcclFloat32, kNcclSum, comm, stream) == 0 ? FRN_OK : FRN_ECUDA;
}
}  // namespace frn
