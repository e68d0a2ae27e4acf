// Store-pattern probe for the depth kernel: how fast can this B200 WRITE 12.4 GB with
//   v0  a plain grid-stride STG.128 fill (what torch.fill_ does),
//   v1  the depth kernel's pattern: persistent warps, one 4 KB tile per warp at a time, 8 x (2 x STG.128) per lane,
//       tiles taken round-robin over all warps of the grid,
//   v2  the same with every CTA owning a CONTIGUOUS range of tiles,
//   v3  v1 through TMA bulk stores (4 KB from shared memory, double buffered),
//   v4  v1 + the depth kernel's shared-memory traffic (zero 4 KB, 8 x LDS.128, shuffles) but no events.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_probe store_probe.cu ; run: ./store_probe [GB]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

constexpr int TILE = 1024, THREADS = 256, WARPS = THREADS / 32;

__global__ void __launch_bounds__(THREADS) v0_fill(int4* __restrict__ out, int64_t n4) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += gridDim.x * (int64_t)blockDim.x)
    out[i] = make_int4(1, 2, 3, 4);
}

template <bool CONTIG>
__global__ void __launch_bounds__(THREADS) v1_tiles(int32_t* __restrict__ out, int64_t n_tiles) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS, w = (int64_t)blockIdx.x * WARPS + warp;
  const int64_t per = (n_tiles + n_warps - 1) / n_warps;
  for (int64_t k = 0; k < per; k++) {
    const int64_t tile = CONTIG ? w * per + k : w + k * n_warps;
    if (tile >= n_tiles) break;
    int32_t* o = out + tile * TILE;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int ia = it * 256 + lane * 4;
      *reinterpret_cast<int4*>(o + ia) = make_int4(it, lane, (int)k, 7);
      *reinterpret_cast<int4*>(o + ia + 128) = make_int4(it, lane, (int)k, 9);
    }
  }
}

__global__ void __launch_bounds__(THREADS) v3_tma(int32_t* __restrict__ out, int64_t n_tiles) {
  extern __shared__ __align__(128) int s_dyn[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* s_warp = s_dyn + warp * 2 * TILE;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS;
  int cur = 0;
  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < n_tiles; tile += n_warps, cur ^= 1) {
    int* buf = s_warp + cur * TILE;
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int v = lane; v < TILE / 4; v += 32) reinterpret_cast<int4*>(buf)[v] = make_int4(lane, v, (int)tile, 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;\n cp.async.bulk.commit_group;" ::"l"(
                       out + tile * TILE),
                   "r"((uint32_t)__cvta_generic_to_shared(buf))
                   : "memory");
  }
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS) v4_smem(int32_t* __restrict__ out, int64_t n_tiles) {
  __shared__ __align__(16) int s_all[WARPS * TILE];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* s = s_all + warp * TILE;
  const int64_t n_warps = (int64_t)gridDim.x * WARPS;
  int carry = 0;
  for (int64_t tile = (int64_t)blockIdx.x * WARPS + warp; tile < n_tiles; tile += n_warps) {
#pragma unroll
    for (int v = lane; v < TILE / 4; v += 32) reinterpret_cast<int4*>(s)[v] = make_int4(0, 0, 0, 0);
    __syncwarp();
    if (lane < 6) atomicAdd(&s[(lane * 173 + (int)tile * 7) & (TILE - 1)], (lane & 1) ? -1 : 1);
    __syncwarp();
    int32_t* o = out + tile * TILE;
#pragma unroll
    for (int it = 0; it < 4; it++) {
      const int ia = it * 256 + lane * 4, ib = ia + 128;
      const int4 va = *reinterpret_cast<const int4*>(&s[ia]);
      const int4 vb = *reinterpret_cast<const int4*>(&s[ib]);
      const int a3 = va.x + va.y + va.z + va.w, b3 = vb.x + vb.y + vb.z + vb.w;
      const unsigned nz = __ballot_sync(0xffffffffu, (a3 | b3) != 0);
      int ex = 0;
      if (nz) {
        int incl = a3 + b3;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        ex = incl - a3 - b3;
        carry += __shfl_sync(0xffffffffu, incl, 31);
      }
      *reinterpret_cast<int4*>(o + ia) = make_int4(carry + ex + va.x, carry + ex + va.y, carry + ex, carry);
      *reinterpret_cast<int4*>(o + ib) = make_int4(carry + ex + vb.x, carry + ex + vb.y, carry + ex, carry);
    }
    __syncwarp();
  }
}

template <class F>
static void run(const char* name, F launch, double bytes) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int rep = 0; rep < 6; rep++) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (rep && ms < best) best = ms;
  }
  const cudaError_t e = cudaGetLastError();
  printf("{\"variant\": \"%s\", \"ms\": %.4f, \"GBps\": %.1f, \"err\": \"%s\"}\n", name, best, bytes / best / 1e6,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main(int argc, char** argv) {
  const double gb = argc > 1 ? atof(argv[1]) : 12.4;
  const int64_t n_tiles = (int64_t)(gb * 1e9 / 4096);
  const double bytes = (double)n_tiles * 4096;
  int32_t* out;
  if (cudaMalloc(&out, (size_t)bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for (int per_sm : {2, 3, 4, 5, 8}) {
    const int grid = sms * per_sm;
    char nm[64];
    snprintf(nm, sizeof nm, "v0_fill_x%d", per_sm);
    run(nm, [&] { v0_fill<<<grid, THREADS>>>((int4*)out, n_tiles * 256); }, bytes);
    snprintf(nm, sizeof nm, "v1_tiles_roundrobin_x%d", per_sm);
    run(nm, [&] { v1_tiles<false><<<grid, THREADS>>>(out, n_tiles); }, bytes);
    snprintf(nm, sizeof nm, "v2_tiles_contiguous_x%d", per_sm);
    run(nm, [&] { v1_tiles<true><<<grid, THREADS>>>(out, n_tiles); }, bytes);
    if (per_sm <= 3) {
      cudaFuncSetAttribute(v3_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
      snprintf(nm, sizeof nm, "v3_tma_x%d", per_sm);
      run(nm, [&] { v3_tma<<<grid, THREADS, 65536>>>(out, n_tiles); }, bytes);
    }
    if (per_sm <= 5) {
      snprintf(nm, sizeof nm, "v4_smem_scan_x%d", per_sm);
      run(nm, [&] { v4_smem<<<grid, THREADS>>>(out, n_tiles); }, bytes);
    }
  }
  cudaFree(out);
  return 0;
}
