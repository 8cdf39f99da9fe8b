// tcgen05.mma issue rate by kind and shape: cycles per MMA (M = 128, K = 32 bytes) for kind::i8 / f8f6f4 / f16,
// N in {64, 128, 256}, one or seven rotating accumulators, operands K-major SWIZZLE_128B in shared memory
// (contents irrelevant).  One CTA per SM so that the numbers include chip-level effects (power, clocks).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_rate tools/umma_rate.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int KIND>
__device__ __forceinline__ void mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  if (KIND == 0)
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else if (KIND == 1)
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

// kind: 0 i8 (s8 x s8 -> s32), 1 f8f6f4 (e4m3 -> f32), 2 f16 (bf16 -> f32)
template <int KIND>
__global__ void __launch_bounds__(128, 1) k_rate(int n_cols, int n_acc, int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // [128 rows][128 B]
  uint8_t* sB = smem + 128 * 128;     // [256 rows][128 B]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 256) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = KIND == 0 ? 0x01010101u : 0u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&s_bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = s_tmem;
  uint32_t idesc;
  if (KIND == 0) idesc = (2u << 4) | (1u << 7) | (1u << 10);
  else if (KIND == 1) idesc = (1u << 4) | (0u << 7) | (0u << 10);
  else idesc = (1u << 4) | (1u << 7) | (1u << 10);
  idesc |= ((uint32_t)(n_cols >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 0) {   // the whole warp runs the loop (uniform control flow); one elected lane issues
    const uint64_t da = make_desc_sw128(smem_u32(sA)), db = make_desc_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 4) {
      const int g = (i >> 2) & (n_acc - 1);
#pragma unroll
      for (int k = 0; k < 4; ++k) mma<KIND>(tmem + g * n_cols, da + 2 * k, db + 2 * k, idesc, i >= 4 * n_acc ? 1u : 0u);
    }
    const long long t1 = clock64();
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(&s_bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred q;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u));
}

// kind::i8, N = 64: groups of 7 MMAs that share the A operand (collector::a fill / use / lastuse), 7 different B
// tiles and 7 accumulators - the digit-pair loop of k_score_i8 reordered so that A stays in the collector.
__device__ __forceinline__ void mma_i8_coll(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc, int mode) {
  if (mode == 0)
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else if (mode == 1)
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else if (mode == 2)
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
  else
    asm volatile("{\n\t.reg .pred p, e;\n\telect.sync _|e, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@e tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
                 :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

template <int REUSE>
__global__ void __launch_bounds__(128, 1) k_rate_coll(int n_mma, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // [128 rows][128 B]
  uint8_t* sB = smem + 128 * 128;     // [7][64 rows][128 B]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (128 + 7 * 64) * 128 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&s_bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = s_tmem;
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
  if (warp == 0) {
    const uint64_t da = make_desc_sw128(smem_u32(sA)), db = make_desc_sw128(smem_u32(sB));
    const long long t0 = clock64();
    for (int i = 0; i < n_mma; i += 28) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int g = 0; g < 7; ++g)
          mma_i8_coll(tmem + g * 64, da + 2 * k, db + g * (64 * 128 / 16) + 2 * k, idesc, 1u,
                      REUSE == 1 ? (g == 0 ? 0 : (g == 6 ? 2 : 1)) : (REUSE == 2 ? 3 : 0));
    }
    const long long t1 = clock64();
    asm volatile("{\n\t.reg .pred e;\n\telect.sync _|e, 0xffffffff;\n\t"
                 "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" :: "r"(smem_u32(&s_bar)) : "memory");
    asm volatile(
        "{\n\t.reg .pred q;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(smem_u32(&s_bar)), "r"(0u) : "memory");
    const long long t2 = clock64();
    if (blockIdx.x == 0 && tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(512u));
}

template <int REUSE>
void run_coll(int grid, long long* d_out) {
  const int reuse = REUSE;
  const int n_mma = 28 * 146;
  const int smem = (128 + 7 * 64) * 128 + 1024;
  CK(cudaFuncSetAttribute(k_rate_coll<REUSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_rate_coll<REUSE><<<grid, 128, smem>>>(n_mma, d_out);
  k_rate_coll<REUSE><<<grid, 128, smem>>>(n_mma, d_out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("i8 N=64, 7 B tiles / accumulators per A, collector reuse %d, grid=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA\n",
         reuse, grid, (double)h[0] / n_mma, (double)h[1] / n_mma);
}

template <int KIND>
void run(const char* name, int n_cols, int n_acc, int grid, long long* d_out) {
  const int n_mma = 4096;
  const int smem = (128 + 256) * 128 + 1024;
  CK(cudaFuncSetAttribute(k_rate<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_rate<KIND><<<grid, 128, smem>>>(n_cols, n_acc, n_mma, d_out);   // warm-up
  k_rate<KIND><<<grid, 128, smem>>>(n_cols, n_acc, n_mma, d_out);
  CK(cudaDeviceSynchronize());
  long long h[2];
  CK(cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost));
  printf("%-7s N=%3d acc=%d grid=%3d: issue %.1f clk/MMA, complete %.1f clk/MMA  (floor %d)\n", name, n_cols, n_acc, grid,
         (double)h[0] / n_mma, (double)h[1] / n_mma, 128 * n_cols / 256);
}

int main() {
  long long* d_out;
  CK(cudaMalloc(&d_out, 16));
  for (int grid : {1, 148}) {
    for (int n_cols : {64, 128, 256}) {
      run<0>("i8", n_cols, 1, grid, d_out);
      run<1>("f8f6f4", n_cols, 1, grid, d_out);
      run<2>("bf16", n_cols, 1, grid, d_out);
    }
    run<0>("i8", 64, 8, grid, d_out);
    run_coll<2>(grid, d_out);
    run_coll<0>(grid, d_out);
    run_coll<1>(grid, d_out);
  }
  return 0;
}
