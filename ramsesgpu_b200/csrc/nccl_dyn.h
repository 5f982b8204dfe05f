// Minimal run-time binding of NCCL (dlopen), so that the library loads on a box without NCCL and
// shares the copy PyTorch has already mapped when it is used from a torchrun rank.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>

namespace rg {

struct NcclApi {
  typedef struct ncclComm* comm_t;
  struct UniqueId { char internal[128]; };
  enum { kInt8 = 0, kFloat32 = 7, kFloat64 = 8, kSum = 0, kMax = 2 };  // ncclDataType_t / ncclRedOp_t values (nccl.h)

  int (*GetUniqueId)(UniqueId*) = nullptr;
  int (*CommInitRank)(comm_t*, int, UniqueId, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;

  // returns nullptr (and sets *err) when libnccl cannot be loaded
  static const NcclApi* get(const char** err);
};

}  // namespace rg
