// Microbenchmark (development): issue rate of tcgen05.mma kind::i8 (M = 128, K = 32) on sm_100a for several N,
// A operand from tensor memory or shared memory, with / without concurrent tensor-memory traffic from other
// warps (tcgen05.st like the A expanders, tcgen05.ld like the epilogue).  Operands are garbage: only the pace
// of the tensor pipe is measured (clock64 around a batch of MMAs committed to one mbarrier).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/mma_rate.cu -o /tmp/mma_rate && /tmp/mma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
    if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum) : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}

struct Sm {
  unsigned long long bar;
  unsigned long long dummy[8];
  uint32_t tmem_base;
  volatile int stop;
};

// mode bits: 1 = A from shared memory (SS) instead of tensor memory (TS); 2 = four warps keep writing the A ring
// (tcgen05.st x32, one per ~pace clocks); 4 = four warps keep reading accumulator columns (tcgen05.ld x32)
__global__ void __launch_bounds__(320, 1) k_rate(int N, int mode, int batches, int per_batch, long long* out_cycles, uint32_t* sink) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;                   // 256 rows x 128 B
  uint8_t* sA = base + 256 * 128;       // 128 rows x 128 B
  Sm* sm = reinterpret_cast<Sm*>(base + 256 * 128 + 128 * 128);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  for (int i = tid; i < (256 + 128) * 128 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u;
  if (tid == 0) {
    mbar_init(smem_u32(&sm->bar), 1);
    for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&sm->dummy[i]), 1);
    sm->stop = 0;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&sm->tmem_base);
  const uint32_t idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t acol = tmem + 384;  // A ring: columns 384..511
  if (w == 9) {
    const uint64_t bd = sw128_desc(smem_u32(sB)), ad = sw128_desc(smem_u32(sA));
    long long t0 = clock64();
    uint32_t ph = 0;
    for (int b = 0; b < batches; ++b) {
      if (mode & 16) {
        // the issue loop of k_tc_score: per K block two barrier polls (always complete here: parity 1 of a fresh
        // barrier), fence, one elected lane issues 4 MMAs + 2 commits
        for (int i0 = 0; i0 < per_batch; i0 += 4) {
          mbar_wait(smem_u32(&sm->dummy[4]), 1);
          mbar_wait(smem_u32(&sm->dummy[5]), 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ts(tmem, acol + ((i0 + kk) & 15) * 8, bd + (uint64_t)(kk * 2), idesc, 1u);
            tc_commit(smem_u32(&sm->dummy[(i0 >> 2) & 1]));
            tc_commit(smem_u32(&sm->dummy[2 + ((i0 >> 2) & 1)]));
          }
          __syncwarp();
        }
      } else if (mode & 32) {
        // the same with the polls of the NEXT pair issued after the MMAs and two K blocks per elected block
        bool r0 = mbar_try(smem_u32(&sm->dummy[4]), 1), r1 = mbar_try(smem_u32(&sm->dummy[5]), 1);
        bool r2 = mbar_try(smem_u32(&sm->dummy[6]), 1), r3 = mbar_try(smem_u32(&sm->dummy[7]), 1);
        for (int i0 = 0; i0 < per_batch; i0 += 8) {
          if (!r0) mbar_wait(smem_u32(&sm->dummy[4]), 1);
          if (!r1) mbar_wait(smem_u32(&sm->dummy[5]), 1);
          if (!r2) mbar_wait(smem_u32(&sm->dummy[6]), 1);
          if (!r3) mbar_wait(smem_u32(&sm->dummy[7]), 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) mma_ts(tmem, acol + ((i0 + kk) & 15) * 8, bd + (uint64_t)((kk & 3) * 2), idesc, 1u);
            tc_commit(smem_u32(&sm->dummy[0]));
            tc_commit(smem_u32(&sm->dummy[1]));
            tc_commit(smem_u32(&sm->dummy[2]));
            tc_commit(smem_u32(&sm->dummy[3]));
          }
          __syncwarp();
          r0 = mbar_try(smem_u32(&sm->dummy[4]), 1); r1 = mbar_try(smem_u32(&sm->dummy[5]), 1);
          r2 = mbar_try(smem_u32(&sm->dummy[6]), 1); r3 = mbar_try(smem_u32(&sm->dummy[7]), 1);
        }
      } else if (mode & 64) {
        // 4 MMAs + 2 commits per elected block, no polls
        for (int i0 = 0; i0 < per_batch; i0 += 4) {
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) mma_ts(tmem, acol + ((i0 + kk) & 15) * 8, bd + (uint64_t)(kk * 2), idesc, 1u);
            tc_commit(smem_u32(&sm->dummy[(i0 >> 2) & 1]));
            tc_commit(smem_u32(&sm->dummy[2 + ((i0 >> 2) & 1)]));
          }
          __syncwarp();
        }
      } else {
      for (int i0 = 0; i0 < per_batch; i0 += 4) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            const int i = i0 + kk;
            const uint32_t dcol = tmem + (((mode & 8) && (i & 4)) ? 128 : 0);
            if (mode & 1) mma_ss(dcol, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, 1u);
            else mma_ts(dcol, acol + (i & 15) * 8, bd + (uint64_t)(kk * 2), idesc, 1u);
          }
        }
        __syncwarp();
      }
      }
      if (elect_one()) tc_commit(smem_u32(&sm->bar));
      __syncwarp();
      mbar_wait(smem_u32(&sm->bar), ph);
      ph ^= 1;
    }
    long long t1 = clock64();
    if (lane == 0) {
      out_cycles[blockIdx.x] = t1 - t0;
      sm->stop = 1;
    }
  } else if (w >= 4 && w < 8 && (mode & 2)) {
    uint32_t v[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) v[c] = 0x01010101u * ((lane + c) & 1);
    const uint32_t lane_base = (uint32_t)((w - 4) * 32) << 16;
    int it = 0;
    while (!sm->stop) {
      tc_st32(acol + lane_base + (it & 3) * 32, v);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      ++it;
      long long t = clock64();
      while (clock64() - t < 300) {}
    }
  } else if (w < 4 && (mode & 4)) {
    uint32_t v[32], acc = 0;
    const uint32_t lane_base = (uint32_t)(w * 32) << 16;
    int it = 0;
    while (!sm->stop) {
      tc_ld32(tmem + lane_base + 192 + (it % 6) * 32, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int c = 0; c < 32; ++c) acc += v[c];
      ++it;
      long long t = clock64();
      while (clock64() - t < 200) {}
    }
    if (acc == 0x12345678u) sink[tid] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (w == 9) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main(int argc, char** argv) {
  const int mode = argc > 1 ? atoi(argv[1]) : 0, N = argc > 2 ? atoi(argv[2]) : 128;
  long long* d;
  uint32_t* sink;
  cudaMalloc(&d, 148 * sizeof(long long));
  cudaMalloc(&sink, 4096);
  const size_t smem = (256 + 128) * 128 + 1024 + 256;
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int batches = 50, per = 96;
  for (int grid : {1, 148}) {
    k_rate<<<grid, 320, smem>>>(N, mode, batches, per, d, sink);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("mode %d N %d grid %d: %s\n", mode, N, grid, cudaGetErrorString(e));
      return 1;
    }
    long long h[148];
    cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("mode %2d (A %s%s%s%s) N %3d grid %3d: %.1f clk per MMA (128*N/256 = %.0f)\n", mode, (mode & 1) ? "smem" : "tmem",
           (mode & 2) ? " +st" : "", (mode & 4) ? " +ld" : "", (mode & 8) ? " altD" : "", N, grid, mx / (double)(batches * per), 128.0 * N / 256.0);
  }
  return 0;
}
