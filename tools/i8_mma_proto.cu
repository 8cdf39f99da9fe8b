// tcgen05.mma kind::i8 mechanics in isolation: D[128 x 64] (s32, TMEM) = A[128 x K] * B[64 x K]^T with int8
// K-major operands in the 128-byte-swizzle shared-memory layout, K = 128 * nchunk.  Checks against the host.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o i8_mma_proto tools/i8_mma_proto.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_128B: rows of 128 bytes, 8-row groups of 1024 bytes (SBO), version 1, layout type 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)1 << 16;                            // LBO (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                  // SBO
  d |= (uint64_t)1 << 46;                            // version
  d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
  return d;
}

constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc), "r"(0u) : "memory");
}

__global__ void __launch_bounds__(128, 1) k_proto(const int8_t* A, const int8_t* B, int32_t* D, int nchunk) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                 // [128 rows][128 B]
  uint8_t* sB = smem + 128 * 128;     // [64 rows][128 B]
  __shared__ uint32_t s_tmem;
  __shared__ __align__(8) uint64_t s_bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int K = 128 * nchunk;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" :: "r"(smem_u32(&s_tmem)), "r"(64u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(&s_bar)), "r"(1u));
    asm volatile("fence.mbarrier_init.release.cluster;\n");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  const uint32_t tmem = s_tmem;
  uint32_t phase = 0;
  for (int c = 0; c < nchunk; ++c) {
    // stage chunk c: 16-byte pieces, piece p of row r lands at piece (p ^ (r & 7))
    for (int i = tid; i < 128 * 8; i += 128) {
      const int r = i >> 3, p = i & 7;
      *reinterpret_cast<int4*>(sA + r * 128 + ((p ^ (r & 7)) << 4)) =
          *reinterpret_cast<const int4*>(A + (size_t)r * K + c * 128 + p * 16);
    }
    for (int i = tid; i < 64 * 8; i += 128) {
      const int r = i >> 3, p = i & 7;
      *reinterpret_cast<int4*>(sB + r * 128 + ((p ^ (r & 7)) << 4)) =
          *reinterpret_cast<const int4*>(B + (size_t)r * K + c * 128 + p * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;\n");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n");
    if (tid == 0) {
      const uint64_t da = make_desc_sw128(smem_u32(sA)), db = make_desc_sw128(smem_u32(sB));
      for (int k = 0; k < 4; ++k) mma_i8(tmem, da + 2 * k, db + 2 * k, kIdesc, (c | k) ? 1u : 0u);
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(smem_u32(&s_bar)) : "memory");
    }
    // everybody waits for the MMAs of this chunk before the staging buffers are overwritten
    asm volatile(
        "{\n\t.reg .pred q;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 q, [%0], %1;\n\t@q bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        :: "r"(smem_u32(&s_bar)), "r"(phase) : "memory");
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n");
  // lane (row) = 32 * warp + lane; 64 columns in two loads of 32
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    const uint32_t addr = tmem + ((uint32_t)(32 * warp) << 16) + 32 * h;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(addr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 32; ++j) D[(size_t)(32 * warp + lane) * 64 + 32 * h + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem), "r"(64u));
}

int main() {
  const int nchunk = 3, K = 128 * nchunk;
  std::vector<int8_t> hA(128 * K), hB(64 * K);
  srand(1);
  for (auto& v : hA) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : hB) v = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB; int32_t* dD;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dD, 128 * 64 * 4));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xff, 128 * 64 * 4));
  const int smem = 128 * 128 + 64 * 128 + 1024;
  CK(cudaFuncSetAttribute(k_proto, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_proto<<<1, 128, smem>>>(dA, dB, dD, nchunk);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD(128 * 64);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 64; ++j) {
      int32_t s = 0;
      for (int k = 0; k < K; ++k) s += (int32_t)hA[i * K + k] * hB[j * K + k];
      if (s != hD[i * 64 + j]) { if (bad < 8) printf("mismatch (%d,%d): got %d want %d\n", i, j, hD[i * 64 + j], s); ++bad; }
    }
  printf("i8_mma_proto: %d mismatches of %d\n", bad, 128 * 64);
  return bad != 0;
}
