// Multi-GPU accumulation of the film buffers: one NCCL communicator per process, ncclReduce(sum) of the
// radiance / weight (/ albedo / normal) planes onto a root rank over NVLink, on the renderer's own stream so that the
// resolve kernel that follows on the root needs no extra synchronisation.  The reference is single-GPU
// (cli/main.cpp:23 creates one cuda::Context); this is new work behind the same boundary.
//
// NCCL is loaded at first use (dlopen "libnccl.so.2"): a single-GPU build host needs no NCCL, and inside a process that
// already carries a NCCL (e.g. PyTorch's) the same library is shared.
#pragma once
#include <cstddef>
#include <cstdint>

#include <cuda_runtime.h>

namespace btrc {

constexpr size_t kNcclIdBytes = 128;  // sizeof(ncclUniqueId)

struct Comm;

void nccl_unique_id(uint8_t out[kNcclIdBytes]);
Comm *comm_create(int n_ranks, int rank, const uint8_t id[kNcclIdBytes]);
void comm_destroy(Comm *c);
int comm_rank(const Comm *c);
int comm_size(const Comm *c);
// in-place sum of `count` floats at `buf` onto rank `root` (other ranks' buffers are left unchanged), grouped over several planes
void comm_reduce_sum(Comm *c, float *const *bufs, const size_t *counts, int n_bufs, int root, cudaStream_t stream);

}  // namespace btrc
