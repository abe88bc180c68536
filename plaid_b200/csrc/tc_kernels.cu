// K0 — the dense-ish block of the gene-set score product on the 5th-generation tensor cores
// (reference R/plaid.R:100-123: Matrix::crossprod(G, X) -> CHOLMOD; here only the block of X rows
// that carry most of the adds, or every row of a dense bulk matrix).
//
//   D[set, (cell, slice)] = sum_g  A[set, g] * B[(cell, slice), g]          (tcgen05.mma kind::i8)
//
//   A = t(G != 0) restricted to the block's genes: 0/1, exact in int8.  It is never stored as
//       bytes: the plan keeps it as BIT masks (128 sets x 128 genes = 2 KB per tile) and four
//       "expander" warps widen one K-block at a time straight into TENSOR MEMORY
//       (tcgen05.st), which the MMA reads as its A operand — so A costs neither L2 nor
//       shared-memory bandwidth and the block may hold any number of genes (K is not bounded
//       by shared memory).
//   B = the block's rows of X as per-column FIXED POINT: q = rint(x * 2^e_j), |q| < 2^(8 SLICES - 2),
//       split into SLICES balanced signed base-256 digits, one int8 row per (cell, digit),
//       K-major ("Bd", written once per call by k_tc_prep_*).  Tiles of 192 rows x 128 genes are
//       fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle) through an 8-stage ring.
//   D = int32 accumulators in TMEM (2 x 192 columns, double buffered against the epilogue).
//       Integer accumulation is EXACT, so the only error of this path is the rounding of x to
//       fixed point: |err| <= 2^-(8 SLICES - 1) * max_g |x_gj| per term (4 slices: 4.7e-10 relative to the
//       column's largest block entry; the north star allows 1e-6).  The epilogue recombines the
//       digits in int64, converts to fp64, undoes the column scale and writes the partial set
//       sums (or, for dense X where there is no scatter pass, the final scores).
//
// Warp roles (448 threads, one CTA per SM): warps 0-3 epilogue (TMEM lanes 32w..32w+31), warps 4-11 expanders
// (A bits -> TMEM; two groups of four warps take alternate K blocks — one group alone could not keep up with the
// MMA: ncu showed the issuer waiting on A 9 polls per K block, tensor pipe 49 %), warp 12 TMA producer, warp 13
// MMA issuer + TMEM allocator.
#include <cuda.h>

#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace plaidgpu {

namespace {

constexpr int TC_M = 128;            // sets per CTA tile (TMEM lanes)
constexpr int TC_N = 192;            // accumulator columns per tile = cells x slices
constexpr int TC_KB = 128;           // genes (= bytes of a B row) per K block; one swizzle-128B row
constexpr int TC_BST = 5;            // B ring stages (shared memory)
constexpr int TC_TBOX = 16 * TC_M * 8;  // one tail box: 128 sets x 16 cells of int64 (128-byte swizzled rows)
constexpr int TC_TBUF = 3 * TC_TBOX;    // tail sums of one accumulator tile (48 cells), double buffered
constexpr int TC_AST = 2;            // A ring stages (tensor memory): a stage is a PAIR of K blocks, 64 columns
constexpr int TC_ACOL = 2 * TC_N;    // first TMEM column of the A ring
constexpr int TC_BSTAGE = TC_N * TC_KB;  // 24,576 bytes
#ifndef TC_EXP_GROUPS
#define TC_EXP_GROUPS 1              // expander groups of four warps (group g widens K blocks g, g + GROUPS, ...)
#endif
constexpr int TC_EPI_WARPS = 8;      // epilogue warps: two per TMEM lane quarter, alternate 32-column chunks
constexpr int TC_W_TMA = TC_EPI_WARPS + 4 * TC_EXP_GROUPS, TC_W_MMA = TC_W_TMA + 1;
constexpr int TC_W_TAIL = TC_W_MMA + 1;  // second TMA producer: the tail boxes (waits on the epilogue, so it must not hold up the B ring)
constexpr int TC_THREADS = 32 * (TC_W_TAIL + 1);

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait: a protocol error traps (the launch fails with an error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(x), "r"(y), "r"(bar)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem descriptor], int8 x int8 -> int32, M = 128, N = 192, K = 32
__device__ __forceinline__ void tc_mma_i8_ts(uint32_t d, uint32_t a, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d), "r"(a), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void tc_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
      "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]),
      "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]),
      "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
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
      : "r"(taddr)
      : "memory");
}

// shared-memory matrix descriptor of one B stage: K-major, 128-byte swizzle, rows of 128 bytes, groups of 8
// rows 1024 bytes apart (cute::UMMA::SmemDescriptor: start >> 4 | LBO << 16 | SBO << 32 | version 1 << 46 |
// SWIZZLE_128B (2) << 61)
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = S32 (2 << 4), A = B = INT8 (1 << 7, 1 << 10),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t TC_IDESC = (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

struct TcSmem {
  unsigned long long b_full[TC_BST], b_empty[TC_BST], a_full[TC_AST], a_empty[TC_AST], acc_full[2], acc_empty[2], tail_full[2], tail_empty[2];
  uint32_t tmem_base;
};

// Epilogue of one 32-column chunk (8 cells x 4 digit columns) of a tile whose cells and rows are all valid, final
// scores (the common case, instruction-minimal: one warp per scheduler cannot hide its own latencies, and the fp64
// pipe is narrow).  Per cell: digits + tail -> int64 (exact), ONE int -> fp64 conversion, the column's power-of-two
// scale applied by an integer add on the exponent field, one DMUL by 1 / (n_s + 1e-8); the sign class of the scores
// (any negative / any zero, all normalize_medians needs up front) is tracked with integer ops.
template <bool RANK, bool CS>
__device__ __forceinline__ void tc_epi_fast(const uint32_t (&v)[32], uint32_t trow, uint32_t tsw, bool has_tail, int ch,
                                            const TcParams& p, int64_t jc, double inv, double nsv, double* __restrict__ o,
                                            uint32_t& negbits, bool& anyzero, const int (&edh)[8]) {
  int ed[8];
  double fb[8], cs[8];
  long long tl[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    ed[c] = edh[c] - 0x3FF00000;  // (exponent of 2^-e_j) << 20; the high words were fetched one chunk ahead
    if (RANK) fb[c] = __ldg(p.colfb + jc + c);
    if (CS) cs[c] = __ldg(p.colscale + jc + c);
    tl[c] = 0;
  }
  if (has_tail) {
#pragma unroll
    for (int c = 0; c < 8; c += 2) {
      const int cell = ch * 8 + c;  // even: cells c, c + 1 share one 16-byte chunk
      const uint32_t a = trow + (uint32_t)(cell >> 4) * TC_TBOX + ((((uint32_t)(cell & 15) >> 1) ^ tsw) << 4);
      asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(tl[c]), "=l"(tl[c + 1]) : "r"(a));
    }
  }
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int lo = (int)v[4 * c] + ((int)v[4 * c + 1] << 8);
    const int hi = (int)v[4 * c + 2] + ((int)v[4 * c + 3] << 8);
    const long long tot = (long long)hi * 65536ll + tl[c] + (long long)lo;
    // exact int64 -> fp64 for |tot| < 2^51 (the fast path's precondition, TcParams::small_sums): tot added to the bit
    // pattern of 2^52 + 2^51 is that double plus tot ulps of 1; one DADD takes the bias off again.  I2F.F64.S64 is a
    // multi-pass instruction on the narrow fp64 pipe, which the 8 epilogue warps saturate (stall_math)
    double x = __longlong_as_double(tot + 0x4338000000000000ll) - 6755399441055744.0;
    int xh = __double2hiint(x);
    xh = tot != 0 ? xh + ed[c] : xh;  // x * 2^-e_j (prep keeps e_j far from the exponent limits)
    x = __hiloint2double(xh, __double2loint(x));
    if (RANK) x = fma(fb[c], nsv, x);
    x *= inv;
    if (CS) x *= cs[c];
    const uint32_t h = (uint32_t)__double2hiint(x);
    negbits |= h;
    anyzero = anyzero || (((h << 1) | (uint32_t)__double2loint(x)) == 0u);
    __stcs(o, x);
    o += p.ld;
  }
}

template <int SLICES>
__global__ void __launch_bounds__(TC_THREADS, 1) k_tc_score(const __grid_constant__ CUtensorMap tmapB,
                                                            const __grid_constant__ CUtensorMap tmapT, const TcParams p) {
  constexpr int CT = TC_N / SLICES;  // cells per tile
  if (p.skip_if && *p.skip_if != 0) return;  // non-finite block entries: the fp64 gather passes run instead
  extern __shared__ uint8_t tc_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sB = base;
  uint8_t* sT = base + (size_t)TC_BST * TC_BSTAGE;  // tail boxes (1024-byte aligned: the stage size is a multiple)
  TcSmem* sm = reinterpret_cast<TcSmem*>(sT + 2 * TC_TBUF);
  const bool has_tail = p.tail != 0;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;

  if (tid == TC_W_MMA * 32) {
    for (int i = 0; i < TC_BST; ++i) {
      mbar_init(smem_u32(&sm->b_full[i]), 1);
      mbar_init(smem_u32(&sm->b_empty[i]), 1);
    }
    for (int i = 0; i < TC_AST; ++i) {
      mbar_init(smem_u32(&sm->a_full[i]), 4);
      mbar_init(smem_u32(&sm->a_empty[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&sm->acc_full[i]), 1);
      mbar_init(smem_u32(&sm->acc_empty[i]), TC_EPI_WARPS);
      mbar_init(smem_u32(&sm->tail_full[i]), 1);
      mbar_init(smem_u32(&sm->tail_empty[i]), TC_EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (w == TC_W_MMA) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&sm->tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(&sm->tmem_base);

  const int m = blockIdx.x;
  const int nct = p.ncell_tiles;
  const int ct0 = (int)(((int64_t)nct * blockIdx.y) / gridDim.y), ct1 = (int)(((int64_t)nct * (blockIdx.y + 1)) / gridDim.y);
  const int KBN = p.kblocks;

  if (w == TC_W_TMA) {
    // ===== TMA producer: B tiles (192 rows x 128 bytes) of cell tile ct, K block kb =====
    uint32_t st = 0, ph = 0;
    for (int ct = ct0; ct < ct1; ++ct) {
      for (int kb = 0; kb < KBN; ++kb) {
        mbar_wait(smem_u32(&sm->b_empty[st]), ph ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&sm->b_full[st]), TC_BSTAGE);
          tma_load_2d(smem_u32(sB + (size_t)st * TC_BSTAGE), &tmapB, kb * TC_KB, ct * TC_N, smem_u32(&sm->b_full[st]));
        }
        __syncwarp();
        if (++st == TC_BST) { st = 0; ph ^= 1; }
      }
    }
  } else if (w == TC_W_TAIL) {
    // ===== second TMA producer: the tail sums (tail_kernels.cu) of each tile's 128 sets x 48 cells, three boxes of
    // 16 cells, double buffered against the epilogue =====
    if (has_tail)
      for (int ct = ct0, t = 0; ct < ct1; ++ct, ++t) {
        const uint32_t tb = t & 1;
        mbar_wait(smem_u32(&sm->tail_empty[tb]), ((t >> 1) & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(smem_u32(&sm->tail_full[tb]), TC_TBUF);
#pragma unroll
          for (int b = 0; b < 3; ++b)
            tma_load_2d(smem_u32(sT + (size_t)tb * TC_TBUF + (size_t)b * TC_TBOX), &tmapT, ct * 48 + b * 16, m * TC_M,
                        smem_u32(&sm->tail_full[tb]));
        }
        __syncwarp();
      }
  } else if (w == TC_W_MMA) {
    // ===== MMA issuer: the whole warp walks the barriers, one elected lane issues (elect.sync keeps the
    // tcgen05 operands in uniform registers; under `if (lane == 0)` ptxas wrapped every tcgen05 instruction in an
    // elect / branch loop and the ~100-instruction issue path, not the tensor pipe, set the pace) =====
    // One issue block per PAIR of K blocks: the A ring is filled in pairs (one tcgen05.st wait of the expanders per
    // 256 genes), the barrier polls of the next pair are issued right after the MMAs of the current one so that
    // their latency runs under the MMAs.
    uint32_t sa = 0, pa = 0, sb = 0, pb = 0;
    const uint64_t bd0 = b_desc(smem_u32(sB));
    auto adv_b = [&](uint32_t& s_, uint32_t& p_) { if (++s_ == TC_BST) { s_ = 0; p_ ^= 1; } };
    const int KP = (KBN + 1) >> 1;  // pairs per cell tile (the last one may hold a single K block)
    const int64_t total_pairs = (int64_t)(ct1 - ct0) * KP;
    uint32_t sb1 = sb, pb1 = pb;
    adv_b(sb1, pb1);
    bool ra = total_pairs > 0 && mbar_try(smem_u32(&sm->a_full[sa]), pa);
    bool rb0 = total_pairs > 0 && mbar_try(smem_u32(&sm->b_full[sb]), pb);
    bool rb1 = total_pairs > 0 && KBN > 1 && mbar_try(smem_u32(&sm->b_full[sb1]), pb1);
    int64_t done = 0;
    for (int ct = ct0, t = 0; ct < ct1; ++ct, ++t) {
      const uint32_t as = t & 1, pacc = (t >> 1) & 1;
      mbar_wait(smem_u32(&sm->acc_empty[as]), pacc ^ 1);
      tc_fence_after();
      const uint32_t dcol = tmem + as * TC_N;
      for (int kp = 0; kp < KP; ++kp) {
        const bool two = 2 * kp + 1 < KBN;
        if (!ra) mbar_wait(smem_u32(&sm->a_full[sa]), pa);
        if (!rb0) mbar_wait(smem_u32(&sm->b_full[sb]), pb);
        if (two && !rb1) mbar_wait(smem_u32(&sm->b_full[sb1]), pb1);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t acol = tmem + TC_ACOL + sa * 64;
          {
            const uint64_t bd = bd0 + (uint64_t)(sb * (TC_BSTAGE >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk)
              tc_mma_i8_ts(dcol, acol + kk * 8, bd + (uint64_t)(kk * 2), TC_IDESC, (kp | kk) != 0 ? 1u : 0u);
            tc_commit(smem_u32(&sm->b_empty[sb]));
          }
          if (two) {
            const uint64_t bd = bd0 + (uint64_t)(sb1 * (TC_BSTAGE >> 4));
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) tc_mma_i8_ts(dcol, acol + 32 + kk * 8, bd + (uint64_t)(kk * 2), TC_IDESC, 1u);
            tc_commit(smem_u32(&sm->b_empty[sb1]));
          }
          tc_commit(smem_u32(&sm->a_empty[sa]));
        }
        __syncwarp();
        ++done;
        if (++sa == TC_AST) { sa = 0; pa ^= 1; }
        adv_b(sb, pb);
        if (two) adv_b(sb, pb);
        sb1 = sb; pb1 = pb;
        adv_b(sb1, pb1);
        const bool more = done < total_pairs;
        const bool two_next = (kp + 1 < KP) ? (2 * (kp + 1) + 1 < KBN) : (KBN > 1);
        ra = more && mbar_try(smem_u32(&sm->a_full[sa]), pa);
        rb0 = more && mbar_try(smem_u32(&sm->b_full[sb]), pb);
        rb1 = more && two_next && mbar_try(smem_u32(&sm->b_full[sb1]), pb1);
      }
      if (elect_one()) tc_commit(smem_u32(&sm->acc_full[as]));
      __syncwarp();
    }
  } else if (w >= TC_EPI_WARPS && w < TC_W_TMA) {
    // ===== expanders: bit masks of A -> int8 0/1 in tensor memory, a PAIR of K blocks (2 x 32 columns) per stage:
    // the tcgen05.st round trip (store, wait::st, fence, arrive) is paid once per 256 genes =====
    const int wq = (w - TC_EPI_WARPS) & 3;
    const int row = wq * 32 + lane;
    const uint4* __restrict__ ab = p.abits + (size_t)m * KBN * TC_M + row;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const int KP = (KBN + 1) >> 1;
    const int64_t total = (int64_t)(ct1 - ct0) * KP;
    auto load_pair = [&](int kp, uint4& b0, uint4& b1) {
      b0 = __ldg(ab + (size_t)(2 * kp) * TC_M);
      b1 = (2 * kp + 1 < KBN) ? __ldg(ab + (size_t)(2 * kp + 1) * TC_M) : make_uint4(0, 0, 0, 0);
    };
    int kp = 0;
    uint4 n0 = make_uint4(0, 0, 0, 0), n1 = n0;
    if (total > 0) load_pair(0, n0, n1);
    for (int64_t q = 0; q < total; ++q) {
      const uint4 bits0 = n0, bits1 = n1;
      if (++kp == KP) kp = 0;
      if (q + 1 < total) load_pair(kp, n0, n1);
      uint32_t v0[32], v1[32];
      const uint32_t wd0[4] = {bits0.x, bits0.y, bits0.z, bits0.w}, wd1[4] = {bits1.x, bits1.y, bits1.z, bits1.w};
#pragma unroll
      for (int c = 0; c < 32; ++c) {  // bit 8b + k of a word = byte b of column k
        v0[c] = (wd0[c >> 3] >> (c & 7)) & 0x01010101u;
        v1[c] = (wd1[c >> 3] >> (c & 7)) & 0x01010101u;
      }
      const uint32_t sa = (uint32_t)(q & (TC_AST - 1)), pa = (uint32_t)((q / TC_AST) & 1);
      mbar_wait(smem_u32(&sm->a_empty[sa]), pa ^ 1);
      tc_fence_after();
      tc_st32(tmem + lane_base + TC_ACOL + sa * 64, v0);
      tc_st32(tmem + lane_base + TC_ACOL + sa * 64 + 32, v1);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&sm->a_full[sa]));
    }
  } else {
    // ===== epilogue: TMEM -> registers -> fp64 partial set sums (or final scores) -> global =====
    // Warps q and q + 4 share TMEM lane quarter q (set rows 32q .. 32q + 31) and take alternate 32-column chunks.
    // One warp per scheduler cannot hide its own latencies, so the per-cell instruction count is what sets the
    // pace: tiles with every cell and every row valid take a predicate-free path.
    const int wq = w & 3, wh = w >> 2;
    const int s = m * TC_M + wq * 32 + lane;
    const bool srow = s < p.S;
    const bool rows_full = m * TC_M + wq * 32 + 31 < p.S;  // warp-uniform
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    double inv = 1.0, nsv = 0.0;
    if (p.final && srow) {
      inv = p.inv[s];
      nsv = p.ns[s];
    }
    const bool rankmode = p.final && p.mode >= XF_SING;
    double vmin = INFINITY;
    uint32_t negbits = 0;   // fast path: OR of the high words of the scores (sign bit = some score is negative)
    bool anyzero = false;   //            some score is exactly zero
    const bool fastable = SLICES == 4 && p.final && rows_full && (!rankmode || p.colfb != nullptr) && !(p.dbg & 32) && p.small_sums;
    constexpr int CPC = 32 / SLICES;  // cells per 32-column chunk
    const uint32_t tsw = (uint32_t)(lane & 7);
    // high words of 2^-e_j for the chunk this warp handles next (fast path): fetched one chunk ahead, so the L2
    // round trip never sits between the accumulator load and the stores
    int edn[8];
    auto fetch_ed = [&](int ct_, int ch_) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int64_t j = (int64_t)ct_ * CT + ch_ * 8 + c;
        edn[c] = (ct_ < ct1 && j < p.N) ? __ldg(reinterpret_cast<const int*>(p.colinv + j) + 1) : 0x3FF00000;
      }
    };
    if (SLICES == 4 && fastable) fetch_ed(ct0, wh);
    for (int ct = ct0, t = 0; ct < ct1; ++ct, ++t) {
      const uint32_t as = t & 1, pacc = (t >> 1) & 1;
      mbar_wait(smem_u32(&sm->acc_full[as]), pacc);
      tc_fence_after();
      if (has_tail) mbar_wait(smem_u32(&sm->tail_full[as]), pacc);
      const int64_t j0 = (int64_t)ct * CT;
      const bool full = j0 + CT <= p.N;  // warp-uniform: only the last cell tile is ragged
      double* __restrict__ optr = p.out + j0 * p.ld + s;
      // this thread's row of the tail boxes: 128-byte rows, 16-byte chunks XOR-swizzled with (row mod 8)
      const uint32_t trow = smem_u32(sT) + as * TC_TBUF + (uint32_t)(wq * 32 + lane) * 128u;
#pragma unroll 1
      for (int ch = wh; ch < ((p.dbg & 16) ? 0 : TC_N / 32); ch += 2) {
        uint32_t v[32];
        tc_ld32(tmem + lane_base + as * TC_N + ch * 32, v);
        int edh[8];
        if (SLICES == 4 && fastable) {
#pragma unroll
          for (int c = 0; c < 8; ++c) edh[c] = edn[c];
          if (ch + 2 < TC_N / 32) fetch_ed(ct, ch + 2);
          else fetch_ed(ct + 1, wh);
        }
        if (SLICES == 4 && fastable && full) {
          double* o = optr + (int64_t)(ch * 8) * p.ld;
          const int64_t jc8 = j0 + ch * 8;
          if (rankmode) {
            if (p.colscale) tc_epi_fast<true, true>(v, trow, tsw, has_tail, ch, p, jc8, inv, nsv, o, negbits, anyzero, edh);
            else tc_epi_fast<true, false>(v, trow, tsw, has_tail, ch, p, jc8, inv, nsv, o, negbits, anyzero, edh);
          } else {
            if (p.colscale) tc_epi_fast<false, true>(v, trow, tsw, has_tail, ch, p, jc8, inv, nsv, o, negbits, anyzero, edh);
            else tc_epi_fast<false, false>(v, trow, tsw, has_tail, ch, p, jc8, inv, nsv, o, negbits, anyzero, edh);
          }
          continue;
        }
        double ci[CPC], fb[CPC], cs[CPC];
        long long tl[CPC];
        const int64_t jc = j0 + ch * CPC;
#pragma unroll
        for (int c = 0; c < CPC; ++c) {
          const bool okc = full || jc + c < p.N;
          ci[c] = okc ? __ldg(p.colinv + jc + c) : 0.0;
          fb[c] = (rankmode && okc) ? (p.colfb ? __ldg(p.colfb + jc + c) : xform_value(p.mode, p.r0 ? p.r0[jc + c] : 0.0, p.a0, p.a1)) : 0.0;
          cs[c] = (p.final && p.colscale && okc) ? __ldg(p.colscale + jc + c) : 1.0;
          tl[c] = 0;
        }
        if (SLICES == 4 && has_tail) {
#pragma unroll
          for (int c = 0; c < CPC; c += 2) {
            const int cell = ch * CPC + c;  // even: cells c, c + 1 share one 16-byte chunk
            const uint32_t a = trow + (uint32_t)(cell >> 4) * TC_TBOX + ((((uint32_t)(cell & 15) >> 1) ^ tsw) << 4);
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(tl[c]), "=l"(tl[c + 1]) : "r"(a));
          }
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        double val[CPC];
#pragma unroll
        for (int c = 0; c < CPC; ++c) {
          // digits -> value: pairs of digits recombine exactly in int32 (|digit sum| <= 128 Kp <= 2^22), then block +
          // tail in int64 (every term is an integer: no rounding before the one conversion to fp64)
          const int lo = (int)v[SLICES * c] + ((int)v[SLICES * c + 1] << 8);
          long long tot = (long long)lo + tl[c];
          if (SLICES == 4) {
            const int hi = (int)v[SLICES * c + 2] + ((int)v[SLICES * c + 3] << 8);
            tot += (long long)hi << 16;
          }
          double x = (double)tot * ci[c];
          if (p.final) {
            if (rankmode) x = fma(fb[c], nsv, x);
            x *= inv;
            if (p.colscale) x *= cs[c];
          }
          val[c] = x;
        }
        if (p.dbg & 8) {
          double a = 0.0;
#pragma unroll
          for (int c = 0; c < CPC; ++c) a += val[c];
          if (a == 1.2345e-300) optr[0] = a;
        } else if (full && rows_full) {
#pragma unroll
          for (int c = 0; c < CPC; ++c) {
            if (p.final) vmin = fmin(vmin, val[c]);
            __stcs(optr + (int64_t)(ch * CPC + c) * p.ld, val[c]);
          }
        } else if (srow) {
#pragma unroll
          for (int c = 0; c < CPC; ++c)
            if (full || jc + c < p.N) {
              if (p.final) vmin = fmin(vmin, val[c]);
              __stcs(optr + (int64_t)(ch * CPC + c) * p.ld, val[c]);
            }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&sm->acc_empty[as]));
        if (has_tail) mbar_arrive(smem_u32(&sm->tail_empty[as]));
      }
    }
    if (p.final && p.smin) {
      // the fast path tracked only the sign class of its scores; a representative stands in for the minimum (the
      // host looks at the sign of the result and nothing else, api.cu)
      if (negbits >> 31) vmin = fmin(vmin, -1.0);
      else if (anyzero) vmin = fmin(vmin, 0.0);
      else if (negbits) vmin = fmin(vmin, 1.0);
      unsigned long long k = vmin == INFINITY ? ~0ull : key_of(vmin);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(FULL, k, o);
        k = other < k ? other : k;
      }
      if (lane == 0 && k != ~0ull) atomicMin(p.smin, k);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (w == TC_W_MMA) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- B operand: fixed-point digit rows ------------------------------------------------------------
// Balanced base-256 digits of q (|q| < 2^(8 SLICES - 2)): d_k in [-128, 127], q = sum d_k 256^k.
template <int SLICES>
__device__ __forceinline__ void digits_of(long long q, signed char (&d)[SLICES]) {
#pragma unroll
  for (int k = 0; k < SLICES; ++k) {
    const int b = (int)(signed char)(q & 0xFF);
    d[k] = (signed char)b;
    q = (q - b) >> 8;
  }
}
// scale exponent: 2^e * maxabs < 2^(8 SLICES - 2)
__device__ __forceinline__ int scale_exp(double maxabs, int slices) {
  if (!(maxabs > 0.0)) return 0;
  int ex;
  frexp(maxabs, &ex);  // maxabs = f * 2^ex, f in [0.5, 1)
  return (8 * slices - 2) - ex;
}

// Sparse X: one warp per column.  Entries of the block's rows are quantised into the column's SLICES digit rows
// (assembled in shared memory, written out with 16-byte stores); every other entry is appended to the compacted
// column the scatter pass walks (as k_compact does).  flag[0] is raised when a block entry is not finite.
template <int SLICES>
__global__ void __launch_bounds__(256) k_tc_prep_csc(const int32_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                                     const double* __restrict__ xx, const double* __restrict__ r0,
                                                     const uint16_t* __restrict__ dmap, int64_t N, int mode, double a0,
                                                     double a1, int Kp, signed char* __restrict__ Bd,
                                                     double* __restrict__ colinv, int32_t* __restrict__ oi,
                                                     double* __restrict__ ox, int32_t* __restrict__ xe,
                                                     int* __restrict__ flag, const int32_t* __restrict__ tmap,
                                                     uint32_t* __restrict__ tcnt, int32_t Pt, int tileC,
                                                     double* __restrict__ colfb) {
  extern __shared__ uint8_t prep_sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, W = blockDim.x >> 5;
  signed char* rows = reinterpret_cast<signed char*>(prep_sm) + (size_t)w * SLICES * Kp;  // [SLICES][Kp]
  const unsigned lt = (1u << lane) - 1u;
  const int vecs = SLICES * Kp / 16;
  for (int64_t j = (int64_t)blockIdx.x * W + w; j < N; j += (int64_t)gridDim.x * W) {
    for (int i = lane; i < vecs; i += 32) reinterpret_cast<uint4*>(rows)[i] = make_uint4(0, 0, 0, 0);
    const int32_t c0 = xp[j], c1 = xp[j + 1];
    double fb = 0.0;
    if (mode >= XF_SING) fb = xform_value(mode, r0 ? r0[j] : 0.0, a0, a1);
    if (colfb && lane == 0) colfb[j] = fb;
    // pass 1: largest |value| among the block entries (and, with a tail pass, the tail entries) of this column
    double mx = 0.0;
    bool bad = false;
    // four iterations at a time: the index / value loads of all four are issued first, then the four map gathers —
    // one warp per column has nothing else to hide the index -> map dependency with
    for (int32_t e0 = c0; e0 < c1; e0 += 128) {
      int32_t rr[4];
      double xv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t e = e0 + 32 * u + lane;
        const bool in = e < c1;
        rr[u] = in ? __ldg(xi + e) : -1;
        xv[u] = in ? __ldg(xx + e) : 0.0;
      }
      bool use[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t r1 = rr[u] < 0 ? 0 : rr[u];
        const unsigned d = __ldg(dmap + r1);
        const int32_t g = tmap ? __ldg(tmap + r1) : -1;
        use[u] = rr[u] >= 0 && (d != 0xFFFFu || g >= 0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (use[u]) {
          double v = xform_value(mode, xv[u], a0, a1);
          if (mode >= XF_SING) v -= fb;
          const double a = fabs(v);
          if (!(a <= 1.0e300)) bad = true;  // NaN or Inf
          mx = fmax(mx, a);
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if (__any_sync(FULL, bad)) {
      if (lane == 0) atomicExch(flag, 1);
      mx = 0.0;
    }
    int ex = scale_exp(mx, SLICES);
    if (ex < -900 || ex > 900) {  // the epilogue scales by an integer add on the exponent field: keep clear of its limits
      if (lane == 0) atomicExch(flag, 1);
      ex = 0;
    }
    const double sc = ldexp(1.0, ex);
    if (lane == 0) colinv[j] = ldexp(1.0, -ex);
    __syncwarp();
    // pass 2: digits of the block entries, compaction of the rest
    int32_t o = c0;
    for (int32_t b0 = c0; b0 < c1; b0 += 128) {
      int32_t rr[4], gg[4];
      double xv[4];
      unsigned dd[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t e = b0 + 32 * u + lane;
        const bool in = e < c1;
        rr[u] = in ? __ldg(xi + e) : -1;
        xv[u] = in ? __ldg(xx + e) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t r1 = rr[u] < 0 ? 0 : rr[u];
        dd[u] = __ldg(dmap + r1);
        gg[u] = tmap ? __ldg(tmap + r1) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (b0 + 32 * u >= c1) break;  // warp-uniform
        const int32_t r = rr[u];
        const double x = xv[u];
        const bool in = r >= 0;
        const bool keep = in && dd[u] == 0xFFFFu;
        if (in && !keep) {
          double v = xform_value(mode, x, a0, a1);
          if (mode >= XF_SING) v -= fb;
          const long long q = (fabs(v) <= 1.0e300) ? __double2ll_rn(v * sc) : 0ll;
          if (q == 0 && v != 0.0) atomicExch(flag, 1);  // would become an exact zero (see k_tc_prep_dense): fp64 passes instead
          signed char dg[SLICES];
          digits_of<SLICES>(q, dg);
#pragma unroll
          for (int k = 0; k < SLICES; ++k) rows[k * Kp + dd[u]] = dg[k];
        }
        const unsigned mk = __ballot_sync(FULL, keep);
        if (keep) {
          const int32_t q = o + __popc(mk & lt);
          oi[q] = r;
          ox[q] = x;
          // entries per (cell tile, tail row): the gene-major regrouping of tail_kernels.cu
          if (gg[u] >= 0) atomicAdd(tcnt + (size_t)(j / tileC) * Pt + gg[u], 1u);
        }
        o += __popc(mk);
      }
    }
    if (lane == 0) xe[j] = o;
    __syncwarp();
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(Bd + (size_t)j * SLICES * Kp);
    for (int i = lane; i < vecs; i += 32) dst[i] = reinterpret_cast<const uint4*>(rows)[i];
    __syncwarp();
  }
}

// Dense X (column-major P x N, every row is in the block, local id = row): one CTA per column.
template <int SLICES>
__global__ void __launch_bounds__(256) k_tc_prep_dense(const double* __restrict__ x, int32_t P, int64_t N, int mode,
                                                       double a0, double a1, int Kp, signed char* __restrict__ Bd,
                                                       double* __restrict__ colinv, int* __restrict__ flag) {
  __shared__ double red[8];
  __shared__ int sbad;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int64_t j = blockIdx.x; j < N; j += gridDim.x) {
    const double* __restrict__ col = x + j * (int64_t)P;
    if (tid == 0) sbad = 0;
    __syncthreads();
    double mx = 0.0;
    bool bad = false;
    for (int r = tid; r < P; r += 256) {
      const double a = fabs(xform_value(mode, col[r], a0, a1));
      if (!(a <= 1.0e300)) bad = true;
      mx = fmax(mx, a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FULL, mx, o));
    if (lane == 0) red[w] = mx;
    if (bad) sbad = 1;
    __syncthreads();
    mx = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) mx = fmax(mx, red[i]);
    if (sbad) {
      if (tid == 0) atomicExch(flag, 1);
      mx = 0.0;
    }
    int ex = scale_exp(mx, SLICES);
    if (ex < -900 || ex > 900) {  // as in k_tc_prep_csc
      if (tid == 0) atomicExch(flag, 1);
      ex = 0;
    }
    const double sc = ldexp(1.0, ex);
    if (tid == 0) colinv[j] = ldexp(1.0, -ex);
    signed char* __restrict__ rows = Bd + (size_t)j * SLICES * Kp;
    bool lost = false;
    // 16 consecutive genes per thread: one 16-byte store per digit row
    for (int g0 = tid * 16; g0 < Kp; g0 += 256 * 16) {
      signed char dg[SLICES][16];
#pragma unroll
      for (int u = 0; u < 16; ++u) {
        const int r = g0 + u;
        long long q = 0;
        if (r < P) {
          const double v = xform_value(mode, col[r], a0, a1);
          q = (fabs(v) <= 1.0e300) ? __double2ll_rn(v * sc) : 0ll;
          // a non-zero entry more than 2^31 below the column's largest would become an exact zero: zeros carry meaning
          // on this path (normalize_medians drops them, R/plaid.R:557-566), so such a column takes the fp64 passes
          if (q == 0 && v != 0.0) lost = true;
        }
        signed char d1[SLICES];
        digits_of<SLICES>(q, d1);
#pragma unroll
        for (int k = 0; k < SLICES; ++k) dg[k][u] = d1[k];
      }
#pragma unroll
      for (int k = 0; k < SLICES; ++k) *reinterpret_cast<uint4*>(rows + (size_t)k * Kp + g0) = *reinterpret_cast<const uint4*>(dg[k]);
    }
    if (lost) atomicExch(flag, 1);
    __syncthreads();
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}

size_t tc_smem_bytes() { return (size_t)TC_BST * TC_BSTAGE + 2 * (size_t)TC_TBUF + sizeof(TcSmem) + 1024; }

}  // namespace

int tc_cells_per_tile(int slices) { return TC_N / slices; }

size_t tc_operand_bytes(int64_t N, int Kp, int slices) {
  const int ct = TC_N / slices;
  const int64_t tiles = (N + ct - 1) / ct;
  return (size_t)tiles * TC_N * (size_t)Kp;
}

cudaError_t launch_tc_prep_csc(const int32_t* xp, const int32_t* xi, const double* xx, const double* r0,
                               const uint16_t* dmap, int64_t N, int mode, double a0, double a1, int Kp, int slices,
                               signed char* Bd, double* colinv, int32_t* oi, double* ox, int32_t* xe, int* flag,
                               const int32_t* tmap, uint32_t* tcnt, int32_t Pt, int tileC, double* colfb, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  if (slices != 2 && slices != 4) return cudaErrorInvalidValue;
  const size_t per_warp = (size_t)slices * Kp;
  int warps = (int)std::min<size_t>(8, (200 * 1024) / per_warp);
  if (warps < 1) return cudaErrorInvalidValue;
  const size_t smem = per_warp * warps;
  const void* fn = slices == 2 ? (const void*)k_tc_prep_csc<2> : (const void*)k_tc_prep_csc<4>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int64_t grid = (N + warps - 1) / warps;
  if (grid > 148 * 8) grid = 148 * 8;
  if (slices == 2)
    k_tc_prep_csc<2><<<(unsigned)grid, warps * 32, smem, st>>>(xp, xi, xx, r0, dmap, N, mode, a0, a1, Kp, Bd, colinv, oi, ox, xe, flag, tmap, tcnt, Pt, tileC, colfb);
  else
    k_tc_prep_csc<4><<<(unsigned)grid, warps * 32, smem, st>>>(xp, xi, xx, r0, dmap, N, mode, a0, a1, Kp, Bd, colinv, oi, ox, xe, flag, tmap, tcnt, Pt, tileC, colfb);
  return cudaGetLastError();
}

cudaError_t launch_tc_prep_dense(const double* x, int32_t P, int64_t N, int mode, double a0, double a1, int Kp,
                                 int slices, signed char* Bd, double* colinv, int* flag, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  if (slices != 2 && slices != 4) return cudaErrorInvalidValue;
  const unsigned grid = (unsigned)std::min<int64_t>(N, 148 * 8);
  if (slices == 2) k_tc_prep_dense<2><<<grid, 256, 0, st>>>(x, P, N, mode, a0, a1, Kp, Bd, colinv, flag);
  else k_tc_prep_dense<4><<<grid, 256, 0, st>>>(x, P, N, mode, a0, a1, Kp, Bd, colinv, flag);
  return cudaGetLastError();
}

cudaError_t launch_tc_score(const TcParams& p0, const signed char* Bd, int Kp, int slices, cudaStream_t st) {
  if (p0.N <= 0) return cudaSuccess;
  if (slices != 2 && slices != 4) return cudaErrorInvalidValue;
  EncodeTiledFn enc = encode_fn();
  if (!enc) return cudaErrorNotSupported;
  TcParams p = p0;
  if (const char* e = getenv("PLAIDGPU_TC_DBG")) p.dbg = atoi(e);
  const int ct = TC_N / slices;
  p.ncell_tiles = (int)((p.N + ct - 1) / ct);
  p.kblocks = Kp / TC_KB;
  CUtensorMap map;
  const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)p.ncell_tiles * TC_N};
  const cuuint64_t gstr[1] = {(cuuint64_t)Kp};
  const cuuint32_t box[2] = {TC_KB, TC_N};
  const cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<signed char*>(Bd), gdim, gstr, box, estr,
          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  CUtensorMap mapT = map;  // unused without a tail
  if (p.tail) {
    if (slices != 4 || p.tail_ld < (int64_t)p.ncell_tiles * ct) return cudaErrorInvalidValue;
    const cuuint64_t tdim[2] = {(cuuint64_t)p.tail_ld, (cuuint64_t)p.S};
    const cuuint64_t tstr[1] = {(cuuint64_t)p.tail_ld * 8};
    const cuuint32_t tbox[2] = {16, TC_M};
    if (enc(&mapT, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<long long*>(p.tail), tdim, tstr, tbox, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
      return cudaErrorInvalidValue;
  }
  const size_t smem = tc_smem_bytes();
  const void* fn = slices == 2 ? (const void*)k_tc_score<2> : (const void*)k_tc_score<4>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int tiles_m = (p.S + TC_M - 1) / TC_M;
  // cell splits: fill whole waves of one CTA per SM (every CTA re-expands A for its own cell tiles, so fewer,
  // longer CTAs are better as long as the last wave is full)
  int best = 1;
  double best_eff = 0.0;
  for (int cs = 1; cs <= 64 && cs <= p.ncell_tiles; ++cs) {
    const int64_t ctas = (int64_t)tiles_m * cs;
    const int64_t waves = (ctas + sms - 1) / sms;
    const double eff = (double)ctas / (double)(waves * sms);
    if (eff > best_eff + 0.02) {
      best_eff = eff;
      best = cs;
    }
  }
  if (const char* e2 = getenv("PLAIDGPU_TC_CSPLIT")) best = std::max(1, std::min(atoi(e2), p.ncell_tiles));
  dim3 grid((unsigned)tiles_m, (unsigned)best);
  if (slices == 2) k_tc_score<2><<<grid, TC_THREADS, smem, st>>>(map, mapT, p);
  else k_tc_score<4><<<grid, TC_THREADS, smem, st>>>(map, mapT, p);
  return cudaGetLastError();
}

}  // namespace plaidgpu
