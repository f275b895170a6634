// dist.cu -- NVLink peer windows for the in-kernel exchange of the sharded fixed point.
//
// One process per GPU.  Each rank allocates a small window (rlvi_fp_dist_inbox_doubles(world) doubles),
// exports it as a CUDA IPC handle, the host side gathers the handles of all ranks (any transport: the
// Python layer uses torch.distributed), and every rank maps its peers' windows.  The fixed-point kernel
// then stores its per-pass partial sums straight into the peers' windows over NVLink
// (fixed_point.cu: grid_allreduce3) -- no NCCL call, no host round trip per pass.
#include "common.cuh"

extern "C" int rlvi_dist_window_create(rlvi_ctx* ctx, int world, void** window_out, unsigned char* handle_out) {
  RLVI_REQUIRE(ctx && window_out && handle_out, "null pointer");
  RLVI_REQUIRE(world >= 1 && world <= 32, "world must be in 1..32");
  RlviDeviceGuard guard(ctx->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == RLVI_IPC_HANDLE_BYTES, "IPC handle size");
  const size_t bytes = size_t(rlvi_fp_dist_inbox_doubles(world)) * sizeof(double);
  void* w = nullptr;
  RLVI_CUDA(cudaMalloc(&w, bytes));
  RLVI_CUDA(cudaMemset(w, 0, bytes));
  RLVI_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, w);
  if (e != cudaSuccess) {
    cudaFree(w);
    rlvi_set_error("cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    return RLVI_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  *window_out = w;
  return RLVI_OK;
}

extern "C" int rlvi_dist_window_open(rlvi_ctx* ctx, int rank, int world, void* own_window,
                                     const unsigned char* all_handles, void** peer_table_out) {
  RLVI_REQUIRE(ctx && own_window && all_handles && peer_table_out, "null pointer");
  RLVI_REQUIRE(world >= 1 && world <= 32 && rank >= 0 && rank < world, "bad rank/world");
  RlviDeviceGuard guard(ctx->device);
  void* host_table[32];
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      host_table[r] = own_window;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, all_handles + size_t(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank) cudaIpcCloseMemHandle(host_table[q]);
      rlvi_set_error("cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
      return RLVI_ERR_CUDA;
    }
    host_table[r] = p;
  }
  void* table = nullptr;
  RLVI_CUDA(cudaMalloc(&table, sizeof(void*) * 32));
  RLVI_CUDA(cudaMemcpy(table, host_table, sizeof(void*) * world, cudaMemcpyHostToDevice));
  *peer_table_out = table;
  return RLVI_OK;
}

extern "C" int rlvi_dist_window_close(rlvi_ctx* ctx, int rank, int world, void* own_window, void* peer_table) {
  RLVI_REQUIRE(ctx != nullptr, "null context");
  RlviDeviceGuard guard(ctx->device);
  cudaDeviceSynchronize();
  if (peer_table) {
    void* host_table[32];
    if (world >= 1 && world <= 32 &&
        cudaMemcpy(host_table, peer_table, sizeof(void*) * world, cudaMemcpyDeviceToHost) == cudaSuccess) {
      for (int r = 0; r < world; ++r)
        if (r != rank && host_table[r]) cudaIpcCloseMemHandle(host_table[r]);
    }
    cudaFree(peer_table);
  }
  if (own_window) cudaFree(own_window);
  return RLVI_OK;
}
