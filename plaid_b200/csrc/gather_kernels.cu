// K2 — gather pass of the gene-set score product over one BLOCK of genes whose rows of X are
// (nearly) dense: the ubiquitous genes of a single-cell matrix (a few hundred genes carry most of
// the adds because they are expressed in most cells AND sit in thousands of sets), or every gene
// of a dense bulk / proteomics matrix (reference R/plaid.R:100-123 with a dense y -> cholmod_sdmult).
//
// For such rows the scatter form (score_kernels.cu) is bound by shared-memory read-modify-write
// (16 B per add, bank conflicts, one gene in flight per accumulator tile).  Here the block's X rows
// are staged DENSE in shared memory for 32 columns at a time ([K+1][32] fp64, row K = zeros), one
// warp owns one set at a time, lanes = columns: every member of the set costs ONE conflict-free
// LDS.64 (256 contiguous bytes) and a register DADD.  The set -> member lists of the block are
// 16-bit local gene ids, padded to multiples of 4 so a warp fetches 4 members with one 8-byte
// broadcast load.  32 finished sets x 32 columns are transposed through a padded shared-memory
// tile so that the global stores are 256-byte contiguous runs along the set axis of the
// column-major output.
#include "common.cuh"

#include <stdlib.h>

#include <algorithm>

namespace plaidgpu {

namespace {

constexpr int GC = 32;   // columns per CTA batch = lanes
constexpr int GW = 20;   // warps per CTA (measured 16 / 20 / 24: 138.4 / 136.1 / 137.4 ms score product per 125k cells)
constexpr int GS = 8;    // sets per staging tile
constexpr int GPAD = 9;

__global__ void __launch_bounds__(GW * 32, 1) k_gather(const GatherParams p) {
  extern __shared__ double gsm[];
  if (p.run_if && *p.run_if == 0) return;  // fallback pass of the tensor-core path: nothing to redo
  double* __restrict__ Xs = gsm;                                   // [(K + 1)][GC]
  double* __restrict__ st = gsm + (size_t)(p.K + 1) * GC;          // [GW][GC][GPAD]
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  double* __restrict__ stw = st + (size_t)w * GC * GPAD;
  const int K = p.K;
  double vmin = INFINITY;  // smallest final score this thread wrote (p.smin)
  const int64_t nbatch = (p.N + GC - 1) / GC;
  // few columns (bulk data): several CTAs share one batch and split the set range between them
  const int nsplit = p.nsplit, split = blockIdx.x % nsplit;
  const int sgroups = (p.S + GS - 1) / GS;                       // groups of GS sets
  const int g_lo = (int)(((int64_t)sgroups * split) / nsplit), g_hi = (int)(((int64_t)sgroups * (split + 1)) / nsplit);

  for (int64_t bt = blockIdx.x / nsplit; bt < nbatch; bt += gridDim.x / nsplit) {
    const int64_t j0 = bt * GC;
    const int nc = (int)min((int64_t)GC, p.N - j0);
    // ---- stage the block's rows of X for these columns ---------------------------------
    if (p.xp) {
      for (int i = tid; i < (K + 1) * GC / 2; i += GW * 32) reinterpret_cast<double2*>(Xs)[i] = make_double2(0.0, 0.0);
      __syncthreads();
      for (int c = w; c < nc; c += GW) {
        const int64_t j = j0 + c;
        const int64_t c0 = p.xp[j], c1 = p.xp[j + 1];
        double fb = 0.0;
        if (p.mode >= XF_SING) fb = xform_value(p.mode, p.r0 ? p.r0[j] : 0.0, p.a0, p.a1);
        for (int64_t e = c0 + lane; e < c1; e += 32) {
          const unsigned d = (unsigned)p.dmap[p.xi[e]] - (unsigned)p.dlo;  // local id inside this block
          if (d < (unsigned)K) {
            double v = xform_value(p.mode, p.xx[e], p.a0, p.a1);
            if (p.mode >= XF_SING) v -= fb;
            Xs[d * GC + c] = v;
          }
        }
      }
    } else {
      // dense mode: rows [g0, g0 + K) of a column-major P x N matrix; rows beyond P and row K are zero
      for (int c = w; c < GC; c += GW) {
        const int64_t j = j0 + c;
        const double* __restrict__ col = p.xx + j * (int64_t)p.P + p.g0;
        const int rows = (c < nc) ? min(K, p.P - p.g0) : 0;
        for (int r = lane; r <= K; r += 32) Xs[r * GC + c] = (r < rows) ? xform_value(p.mode, col[r], p.a0, p.a1) : 0.0;
      }
    }
    __syncthreads();

    // ---- gather: one warp per group of GS consecutive sets, lanes = columns ---------------------
    // The member lists of consecutive sets are contiguous, so a group is ONE stream of 4-member
    // chunks; chunks are fetched 4..8 ahead of their use (the L2 latency of the list loads was the
    // top stall), set boundaries are multiples of a chunk and are held one per lane.
    const char* __restrict__ xlane = reinterpret_cast<const char*>(Xs) + lane * 8;
    const uint4* __restrict__ ids = reinterpret_cast<const uint4*>(p.didx);
    const uint4 zrow = make_uint4((unsigned)K * 256u, (unsigned)K * 256u, (unsigned)K * 256u, (unsigned)K * 256u);
    for (int s0 = (g_lo + w) * GS; s0 < g_hi * GS && s0 < p.S; s0 += GW * GS) {
      const int ns = min(GS, p.S - s0);
      const uint32_t bnd = (lane <= ns) ? (p.dptr[s0 + lane] >> 2) : 0u;  // chunk index where set lane starts
      const uint32_t cbeg = __shfl_sync(FULL, bnd, 0), cend = __shfl_sync(FULL, bnd, ns);
      int i = 0;
      uint32_t nextb = __shfl_sync(FULL, bnd, 1);
      double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
      uint4 q[4], nq[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) q[c] = (cbeg + c < cend) ? __ldg(ids + cbeg + c) : zrow;
      for (uint32_t m = cbeg; m < cend || i < ns; m += 4) {
#pragma unroll
        for (int c = 0; c < 4; ++c) nq[c] = (m + 4 + c < cend) ? __ldg(ids + m + 4 + c) : zrow;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          while (i < ns && m + c == nextb) {  // set i is complete (possibly empty): stage its 32 column sums
            stw[lane * GPAD + i] = (a0 + a1) + (a2 + a3);
            a0 = a1 = a2 = a3 = 0.0;
            ++i;
            nextb = __shfl_sync(FULL, bnd, min(i + 1, 31));
          }
          const uint4 qq = q[c];
          a0 += *reinterpret_cast<const double*>(xlane + qq.x);
          a1 += *reinterpret_cast<const double*>(xlane + qq.y);
          a2 += *reinterpret_cast<const double*>(xlane + qq.z);
          a3 += *reinterpret_cast<const double*>(xlane + qq.w);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) q[c] = nq[c];
      }
      __syncwarp();
      // transposed write-out: 4 columns x 8 consecutive sets (64 contiguous bytes each) per store
      const int si = lane & 7, cg = lane >> 3;
      if (si < ns) {
        const int s = s0 + si;
        for (int c = cg; c < nc; c += 4) {
          const int64_t j = j0 + c;
          double* __restrict__ o = p.out + j * p.ld + s;
          double v = stw[c * GPAD + si];
          if (p.accumulate) v += __ldcs(o);
          if (p.final) {
            double fb = 0.0;
            if (p.mode >= XF_SING) fb = xform_value(p.mode, p.r0 ? p.r0[j] : 0.0, p.a0, p.a1);
            v = score_epilogue(v, s, j, fb, p.mode, p.inv, p.ns, p.colscale);
            vmin = fmin(vmin, v);
          }
          __stcs(o, v);
        }
      }
      __syncwarp();
    }
    __syncthreads();  // Xs is rebuilt by the next batch
  }
  if (p.final && p.smin) {  // smallest final score of this warp -> one atomicMin
    unsigned long long k = vmin == INFINITY ? ~0ull : key_of(vmin);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(FULL, k, o);
      k = other < k ? other : k;
    }
    if (lane == 0 && k != ~0ull) atomicMin(p.smin, k);
  }
}

// one warp per column: sum of |f(x)| over the column's stored entries (sparse) or P rows (dense)
__global__ void __launch_bounds__(256) k_colabs(const int32_t* __restrict__ xp, const double* __restrict__ xx,
                                                int32_t P, int64_t N, int mode, double a0, double a1, int kind,
                                                double* __restrict__ colscale) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = w0; j < N; j += nw) {
    const int64_t c0 = xp ? xp[j] : j * (int64_t)P;
    const int64_t c1 = xp ? xp[j + 1] : c0 + P;
    double a = 0.0;
    for (int64_t e = c0 + lane; e < c1; e += 32) a += fabs(xform_value(mode, xx[e], a0, a1));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(FULL, a, o);
    if (lane == 0) colscale[j] = (kind == 1) ? 100.0 / (a + 1e-8) : 1.0 / (a / (double)P + 1e-8);
  }
}

size_t gather_smem(int K) { return ((size_t)(K + 1) * GC + (size_t)GW * GC * GPAD) * sizeof(double); }

}  // namespace

int gather_max_block(int device) {
  int optin = 0;
  if (cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) return 0;
  const size_t fixed = (size_t)GW * GC * GPAD * sizeof(double) + 1024;
  if ((size_t)optin <= fixed) return 0;
  int K = (int)(((size_t)optin - fixed) / (GC * sizeof(double))) - 1;
  K = (K / 32) * 32;
  if (const char* e = getenv("PLAIDGPU_GATHER_K")) {  // tuning knob (bench / profiling only)
    const int v = atoi(e);
    if (v >= 0 && v < K) K = (v / 32) * 32;
  }
  return K > 0xFFF0 ? 0xFFF0 : K;
}

cudaError_t launch_gather(const GatherParams& p, cudaStream_t st) {
  if (p.N <= 0 || p.K <= 0) return cudaSuccess;
  const size_t smem = gather_smem(p.K);
  cudaError_t e = cudaFuncSetAttribute(k_gather, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int per_sm = 1;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gather, GW * 32, smem);
  if (per_sm < 1) per_sm = 1;
  int64_t grid = (int64_t)sms * per_sm;
  const int64_t nbatch = (p.N + GC - 1) / GC;
  GatherParams q = p;
  q.nsplit = 1;
  if (nbatch < grid) {
    q.nsplit = (int)std::max<int64_t>(1, std::min<int64_t>(grid / nbatch, 16));
    grid = nbatch * q.nsplit;
  }
  k_gather<<<(unsigned)grid, GW * 32, smem, st>>>(q);
  return cudaGetLastError();
}

cudaError_t launch_colabs(const int32_t* xp, const double* xx, int32_t P, int64_t N, int mode, double a0,
                          double a1, int kind, double* colscale, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t grid = (N + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  k_colabs<<<(unsigned)grid, 256, 0, st>>>(xp, xx, P, N, mode, a0, a1, kind, colscale);
  return cudaGetLastError();
}

}  // namespace plaidgpu
