// K5 — normalize_medians() (reference R/plaid.R:554-575): per-column medians of the S x N score
// matrix (matrixStats::colMedians(na.rm=TRUE), with and without the zeros), the column minima
// that decide the global `ignore.zero` switch, and the fused fix-up
//   out = alpha * (x - med_j + mean(med)) + beta_s.
//
// Medians are exact order statistics: an MSD radix select on order-preserving 64-bit keys
// (11-bit digits, shared-memory histograms), which stops as soon as the bucket holding the
// wanted rank is small and finishes by ranking the few candidates directly.  Up to four ranks
// are selected together (two middles x {all values, non-zero values}); removing the zeros only
// shifts ranks because the zeros are one contiguous block of the sorted column.
#include "common.cuh"

#include <math.h>

namespace plaidgpu {

namespace {

constexpr int NT = 256;     // threads per column CTA
constexpr int NBIN = 2048;  // 11-bit digits
constexpr int NTGT = 4;     // simultaneous order statistics
constexpr int CAND = 256;   // finish by direct ranking below this bucket size
constexpr int NDIG = 6;     // 11+11+11+11+11+9 bits
constexpr unsigned long long ZERO_KEY = 0x8000000000000000ull;

__device__ __forceinline__ int digit_shift(int d) { return d < 5 ? 53 - 11 * d : 0; }
__device__ __forceinline__ int digit_bits(int d) { return d < 5 ? 11 : 9; }

struct StatsSmem {
  unsigned hist[NTGT][NBIN];
  unsigned long long cand[NTGT][CAND];
  unsigned long long prefix[NTGT];  // key bits decided so far (high digits)
  unsigned long long result[NTGT];
  unsigned k[NTGT];      // remaining 0-based rank inside the current bucket
  unsigned cnt[NTGT];    // size of the current bucket
  unsigned ncand[NTGT];
  int active[NTGT];
  int depth[NTGT];       // digits decided
  unsigned red_u[NT / 32][3];
  unsigned long long red_k[NT / 32];
  unsigned nnan, nzero, nneg;
  unsigned long long minkey;
};

// One warp finds, for target g, the bin of hist[g] (digit d) that holds rank k[g].
__device__ void pick_bin(StatsSmem& s, int g, int d) {
  const int lane = threadIdx.x & 31;
  const int nb = 1 << digit_bits(d);
  const unsigned k = s.k[g];
  unsigned base = 0;
  int found = -1;
  unsigned before = 0, inbin = 0;
  for (int b0 = 0; b0 < nb && found < 0; b0 += 32) {
    const unsigned c = s.hist[g][b0 + lane];
    unsigned incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned tot = __shfl_sync(FULL, incl, 31);
    if (k < base + tot) {
      const unsigned hit = __ballot_sync(FULL, k < base + incl);
      const int l = __ffs(hit) - 1;
      found = b0 + l;
      before = base + __shfl_sync(FULL, incl - c, l);
      inbin = __shfl_sync(FULL, c, l);
    }
    base += tot;
  }
  __syncwarp();
  if (lane == 0) {
    s.prefix[g] |= (unsigned long long)found << digit_shift(d);
    s.k[g] = k - before;
    s.cnt[g] = inbin;
    s.depth[g] = d + 1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(NT) k_colstats(const double* __restrict__ x, int64_t ld, int32_t S,
                                                 int64_t N, double* __restrict__ med_all,
                                                 double* __restrict__ med_nz,
                                                 double* __restrict__ colmin) {
  extern __shared__ unsigned char smem_raw[];
  StatsSmem& s = *reinterpret_cast<StatsSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;

  for (int64_t j = blockIdx.x; j < N; j += gridDim.x) {
    const double* __restrict__ c = x + j * ld;

    // ---- pass 1: top-digit histogram + NaN / zero counts + min -------------------------
    for (int i = tid; i < NBIN; i += NT) s.hist[0][i] = 0;
    __syncthreads();
    unsigned nnan = 0, nzero = 0;
    unsigned long long mink = ~0ull;
    int runbin = -1;
    unsigned runcnt = 0;
    for (int l = tid; l < S; l += NT) {
      const double v = c[l];
      if (v != v) {
        ++nnan;
        continue;
      }
      const unsigned long long key = key_of(v);
      nzero += (key == ZERO_KEY);
      mink = key < mink ? key : mink;
      const int bin = (int)(key >> 53);
      if (bin != runbin) {  // run-length aggregation: the scores of a column share 1-3 binades
        if (runcnt) atomicAdd(&s.hist[0][runbin], runcnt);
        runbin = bin;
        runcnt = 0;
      }
      ++runcnt;
    }
    if (runcnt) atomicAdd(&s.hist[0][runbin], runcnt);
    __syncthreads();
    // negatives = keys with the top bit clear = bins [0, NBIN/2)
    unsigned nneg = 0;
    for (int i = tid; i < NBIN / 2; i += NT) nneg += s.hist[0][i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      nnan += __shfl_xor_sync(FULL, nnan, o);
      nzero += __shfl_xor_sync(FULL, nzero, o);
      nneg += __shfl_xor_sync(FULL, nneg, o);
      const unsigned long long mk = __shfl_xor_sync(FULL, mink, o);
      mink = mk < mink ? mk : mink;
    }
    if (lane == 0) {
      s.red_u[wid][0] = nnan;
      s.red_u[wid][1] = nzero;
      s.red_u[wid][2] = nneg;
      s.red_k[wid] = mink;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned a = 0, z = 0, ng = 0;
      unsigned long long mk = ~0ull;
      for (int i = 0; i < NT / 32; ++i) {
        a += s.red_u[i][0];
        z += s.red_u[i][1];
        ng += s.red_u[i][2];
        mk = s.red_k[i] < mk ? s.red_k[i] : mk;
      }
      s.nnan = a;
      s.nzero = z;
      s.nneg = ng;
      s.minkey = mk;
      const unsigned nvalid = (unsigned)S - a;
      for (int g = 0; g < NTGT; ++g) {
        s.active[g] = 0;
        s.prefix[g] = 0;
        s.ncand[g] = 0;
        s.cnt[g] = nvalid;
        s.depth[g] = 0;
        s.k[g] = 0;
        s.result[g] = 0;
      }
      // targets: overall 0-based ranks among the sorted non-NaN values
      if (nvalid > 0) {
        s.active[0] = 1;
        s.k[0] = (nvalid - 1) / 2;
        s.active[1] = 1;
        s.k[1] = nvalid / 2;  // == k[0] when the count is odd
      }
      const unsigned m2 = nvalid - z;
      if (m2 > 0) {
        const unsigned r0 = (m2 - 1) / 2, r1 = m2 / 2;
        s.active[2] = 1;
        s.k[2] = r0 < ng ? r0 : r0 + z;
        s.active[3] = 1;
        s.k[3] = r1 < ng ? r1 : r1 + z;
      }
    }
    __syncthreads();
    // digit 0 was histogrammed once for all targets
    for (int g = 1; g < NTGT; ++g)
      if (s.active[g])
        for (int i = tid; i < NBIN; i += NT) s.hist[g][i] = s.hist[0][i];
    __syncthreads();
    if (wid < NTGT && s.active[wid]) pick_bin(s, wid, 0);
    __syncthreads();

    // ---- refine digit by digit while a bucket is large (a target that stops never resumes) --
    for (int d = 1; d < NDIG; ++d) {
      bool need = false;
#pragma unroll
      for (int g = 0; g < NTGT; ++g) need |= (s.active[g] && s.cnt[g] > CAND);
      if (!need) break;
      const int sh = digit_shift(d), nb = 1 << digit_bits(d);
      const int psh = digit_shift(d - 1);  // bits >= psh are decided for refining targets
      bool ref[NTGT];
      unsigned long long pf[NTGT];
#pragma unroll
      for (int g = 0; g < NTGT; ++g) {
        ref[g] = s.active[g] && s.cnt[g] > CAND;
        pf[g] = s.prefix[g] >> psh;
        if (ref[g])
          for (int i = tid; i < nb; i += NT) s.hist[g][i] = 0;
      }
      __syncthreads();
      for (int l = tid; l < S; l += NT) {
        const double v = c[l];
        if (v != v) continue;
        const unsigned long long key = key_of(v);
        const unsigned long long hi = key >> psh;
        const unsigned bin = (unsigned)(key >> sh) & (unsigned)(nb - 1);
#pragma unroll
        for (int g = 0; g < NTGT; ++g)
          if (ref[g] && hi == pf[g]) atomicAdd(&s.hist[g][bin], 1u);
      }
      __syncthreads();
      if (wid < NTGT && ref[wid]) pick_bin(s, wid, d);
      __syncthreads();
    }

    // ---- collect the candidates of every unfinished target and rank them directly ---------
    bool collect = false;
    bool col[NTGT];
    unsigned long long pf[NTGT];
    int psh_g[NTGT];
#pragma unroll
    for (int g = 0; g < NTGT; ++g) {
      col[g] = s.active[g] && s.depth[g] < NDIG;
      psh_g[g] = digit_shift(s.depth[g] > 0 ? s.depth[g] - 1 : 0);
      pf[g] = s.prefix[g] >> psh_g[g];
      collect |= col[g];
      if (s.active[g] && s.depth[g] == NDIG && tid == 0) s.result[g] = s.prefix[g];
    }
    if (collect) {
      for (int l = tid; l < S; l += NT) {
        const double v = c[l];
        if (v != v) continue;
        const unsigned long long key = key_of(v);
#pragma unroll
        for (int g = 0; g < NTGT; ++g)
          if (col[g] && (key >> psh_g[g]) == pf[g]) {
            const unsigned pos = atomicAdd(&s.ncand[g], 1u);
            if (pos < CAND) s.cand[g][pos] = key;
          }
      }
      __syncthreads();
#pragma unroll
      for (int g = 0; g < NTGT; ++g) {
        if (!col[g]) continue;
        const unsigned n = min(s.ncand[g], (unsigned)CAND);
        if ((unsigned)tid < n) {
          const unsigned long long mine = s.cand[g][tid];
          unsigned r = 0;
          for (unsigned u = 0; u < n; ++u) {
            const unsigned long long o = s.cand[g][u];
            r += (o < mine) || (o == mine && u < (unsigned)tid);
          }
          if (r == s.k[g]) s.result[g] = mine;
        }
      }
    }
    __syncthreads();
    if (tid == 0) {
      const unsigned nvalid = (unsigned)S - s.nnan;
      double ma = nan(""), mz = 0.0;
      if (nvalid > 0) {
        const double a = value_of(s.result[0]), b = value_of(s.result[1]);
        ma = (nvalid & 1u) ? a : (a + b) / 2.0;
      }
      const unsigned m2 = nvalid - s.nzero;
      if (m2 > 0) {
        const double a = value_of(s.result[2]), b = value_of(s.result[3]);
        mz = (m2 & 1u) ? a : (a + b) / 2.0;
      }
      med_all[j] = ma;
      med_nz[j] = mz;
      colmin[j] = nvalid > 0 ? value_of(s.minkey) : INFINITY;
    }
    __syncthreads();
  }
}

// out[s, j] = alpha * (x[s, j] - med[j] + c) + beta[s]   for columns j0 <= j < j1
__global__ void __launch_bounds__(256) k_fixup(const double* __restrict__ x, double* __restrict__ out,
                                               int64_t ld, int32_t S, int64_t j0, int64_t j1,
                                               const double* __restrict__ med, double c, double alpha,
                                               const double* __restrict__ beta) {
  for (int64_t j = j0 + blockIdx.x; j < j1; j += gridDim.x) {
    const double shift = (med ? -med[j] : 0.0) + c;
    const double* __restrict__ xi = x + j * ld;
    double* __restrict__ oi = out + j * ld;
    const bool vec = ((((uintptr_t)xi) | ((uintptr_t)oi)) & 15) == 0;
    if (vec) {
      const int S2 = S >> 1;
      const double2* __restrict__ x2 = reinterpret_cast<const double2*>(xi);
      double2* __restrict__ o2 = reinterpret_cast<double2*>(oi);
      for (int l = threadIdx.x; l < S2; l += blockDim.x) {
        double2 v = __ldcs(x2 + l);
        v.x = alpha * (v.x + shift);
        v.y = alpha * (v.y + shift);
        if (beta) {
          v.x += beta[2 * l];
          v.y += beta[2 * l + 1];
        }
        __stcs(o2 + l, v);
      }
      if ((S & 1) && threadIdx.x == 0) {
        double v = alpha * (xi[S - 1] + shift);
        if (beta) v += beta[S - 1];
        oi[S - 1] = v;
      }
    } else {
      for (int l = threadIdx.x; l < S; l += blockDim.x) {
        double v = alpha * (xi[l] + shift);
        if (beta) v += beta[l];
        oi[l] = v;
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_minmax(const double* __restrict__ x, int64_t n,
                                                unsigned long long* __restrict__ res) {
  unsigned long long lo = ~0ull, hi = 0ull;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    const double v = x[i];
    if (v != v) continue;
    const unsigned long long k = key_of(v);
    lo = k < lo ? k : lo;
    hi = k > hi ? k : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long a = __shfl_xor_sync(FULL, lo, o), b = __shfl_xor_sync(FULL, hi, o);
    lo = a < lo ? a : lo;
    hi = b > hi ? b : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(res, lo);
    atomicMax(res + 1, hi);
  }
}

__global__ void k_minmax_init(unsigned long long* res) {
  res[0] = ~0ull;
  res[1] = 0ull;
}
__global__ void k_minmax_fin(const unsigned long long* res, double* out) {
  out[0] = res[0] == ~0ull ? INFINITY : value_of(res[0]);
  out[1] = res[1] == 0ull ? -INFINITY : value_of(res[1]);
}

}  // namespace

cudaError_t launch_colstats(const double* x, int64_t ld, int32_t S, int64_t N, double* med_all,
                            double* med_nz, double* colmin, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k_colstats, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(StatsSmem));
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)sms * 4;  // 4 x 40 KB of shared memory per SM
  if (grid > N) grid = N;
  k_colstats<<<(unsigned)grid, NT, sizeof(StatsSmem), st>>>(x, ld, S, N, med_all, med_nz, colmin);
  return cudaGetLastError();
}

cudaError_t launch_fixup(const double* x, double* out, int64_t ld, int32_t S, int64_t j0, int64_t j1,
                         const double* med, double c, double alpha, const double* beta,
                         cudaStream_t st) {
  if (j1 <= j0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t grid = (int64_t)sms * 8;
  if (grid > j1 - j0) grid = j1 - j0;
  k_fixup<<<(unsigned)grid, 256, 0, st>>>(x, out, ld, S, j0, j1, med, c, alpha, beta);
  return cudaGetLastError();
}

// res2: device double[2]; uses 16 bytes right after it as scratch -> caller passes double[4]
cudaError_t launch_minmax(const double* x, int64_t n, double* res4, cudaStream_t st) {
  unsigned long long* scratch = reinterpret_cast<unsigned long long*>(res4 + 2);
  k_minmax_init<<<1, 1, 0, st>>>(scratch);
  if (n > 0) {
    int64_t grid = (n + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    k_minmax<<<(unsigned)grid, 256, 0, st>>>(x, n, scratch);
  }
  k_minmax_fin<<<1, 1, 0, st>>>(scratch, res4);
  return cudaGetLastError();
}

}  // namespace plaidgpu
