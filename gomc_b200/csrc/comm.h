// comm.h -- the engine's NCCL communicator (multi-GPU sharding).
//
// libnccl.so.2 is bound at run time (dlopen): a single-GPU process never needs it, and a
// process that already carries NCCL (torch) shares its copy.  One communicator per engine;
// every collective is queued on the engine's own stream, so the exchange is part of the
// device timeline (CUDA events around a step include it) and costs no host round trip.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <string>

namespace gbc {

struct Comm;

// 128-byte NCCL unique id (rank 0 creates it, every rank passes the same bytes to comm_create)
int comm_unique_id(void *out128, std::string &err);
Comm *comm_create(const void *id128, int rank, int world, std::string &err);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_world(const Comm *c);
// sum of n doubles over all ranks, in place
int comm_allreduce_sum(Comm *c, double *buf, size_t n, cudaStream_t st, std::string &err);
// in-place all-gather: rank r's bytesPerRank bytes sit at buf + r * bytesPerRank
int comm_allgather_inplace(Comm *c, void *buf, size_t bytesPerRank, cudaStream_t st,
                           std::string &err);

}  // namespace gbc
