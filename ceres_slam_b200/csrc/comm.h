// NCCL plumbing for the multi-GPU reduction of the partial reduced camera systems.
// libnccl is loaded at run time (dlopen) so a single-GPU process never needs it.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cslam {
void comm_unique_id(uint8_t id[128]);
void* comm_create(int n_ranks, int rank, const uint8_t id[128]);
void comm_destroy(void* comm);
void comm_allreduce_sum(void* comm, double* buf, size_t count, cudaStream_t s);
void comm_allreduce_max(void* comm, double* buf, size_t count, cudaStream_t s);
void comm_broadcast(void* comm, double* buf, size_t count, int root, cudaStream_t s);
// in-place all-gather of raw bytes: rank r's chunk is buf[r * chunk_bytes .. (r + 1) * chunk_bytes)
void comm_allgather_bytes(void* comm, void* buf, size_t chunk_bytes, int rank, cudaStream_t s);
}  // namespace cslam
