// Host -> device upload of 192 separate 1.5 MB pinned buffers: per-copy cudaMemcpyAsync on one / two queues,
// cudaMemcpyBatchAsync, and one contiguous copy.  nvcc -O2 -o h2d_batch_probe h2d_batch_probe.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
int main() {
  const int B = 192; const size_t bytes = 131072 * 12;
  char* host; CK(cudaMallocHost(&host, bytes * B));
  char* dev; CK(cudaMalloc(&dev, bytes * B));
  cudaStream_t s0, s1; CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
  cudaEvent_t a, b, j; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b)); CK(cudaEventCreateWithFlags(&j, cudaEventDisableTiming));
  std::vector<void*> dsts(B), srcs(B); std::vector<size_t> sizes(B, bytes);
  for (int i = 0; i < B; ++i) { dsts[i] = dev + i * bytes; srcs[i] = host + (size_t)((i * 7) % B) * bytes; }
  const int reps = 10;
  auto report = [&](const char* name) { float ms; cudaEventElapsedTime(&ms, a, b); std::printf("%-34s %.2f GB/s\n", name, bytes * B * reps / (ms * 1e-3) / 1e9); };
  for (int mode = 0; mode < 5; ++mode) {
    for (int warm = 0; warm < 2; ++warm) {
      CK(cudaEventRecord(a, s0));
      for (int r = 0; r < reps; ++r) {
        if (mode == 0) for (int i = 0; i < B; ++i) CK(cudaMemcpyAsync(dsts[i], srcs[i], bytes, cudaMemcpyHostToDevice, s0));
        if (mode == 1) {
          for (int i = 0; i < B; ++i) CK(cudaMemcpyAsync(dsts[i], srcs[i], bytes, cudaMemcpyHostToDevice, (i & 1) ? s1 : s0));
          CK(cudaEventRecord(j, s1)); CK(cudaStreamWaitEvent(s0, j, 0));
        }
        if (mode == 2 || mode == 3) {
          cudaMemcpyAttributes at{}; at.srcAccessOrder = cudaMemcpySrcAccessOrderStream;
          at.flags = mode == 3 ? cudaMemcpyFlagPreferOverlapWithCompute : 0;
          size_t idx = 0, fail = 0;
          CK(cudaMemcpyBatchAsync(dsts.data(), srcs.data(), sizes.data(), B, &at, &idx, 1, &fail, s0));
        }
        if (mode == 4) CK(cudaMemcpyAsync(dev, host, bytes * B, cudaMemcpyHostToDevice, s0));
      }
      CK(cudaEventRecord(b, s0)); CK(cudaEventSynchronize(b));
    }
    report(mode == 0 ? "192 copies, one queue" : mode == 1 ? "192 copies, two queues" : mode == 2 ? "cudaMemcpyBatchAsync" : mode == 3 ? "cudaMemcpyBatchAsync (overlap flag)" : "one contiguous copy");
  }
  return 0;
}
