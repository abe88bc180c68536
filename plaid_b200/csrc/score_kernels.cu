// K1/K2 — the gene-set score product  out = t(G) %*% X  (scaled per set), the hot loop of
// plaid() / chunked_crossprod() (reference R/plaid.R:60-87, 100-123; Matrix::crossprod ->
// CHOLMOD ssmult/sdmult).
//
// Formulation (B200-first, not a translation of Gustavson-on-CPU):
//   * scatter form: for every stored x(g, j) add it to all sets containing gene g.  Work is
//     nnz(X) x avg-degree adds, the work-optimal count for sparse X sparse G -> dense out.
//   * accumulators are fp64 and live in SHARED MEMORY; one warp owns one tile of Ts sets of
//     one column at a time, so no atomics are needed: the sets of one gene are distinct, and
//     genes are processed one after the other inside the warp.
//   * the set axis is tiled (T tiles of Ts sets) because one fp64 column of 30k sets (240 KB)
//     exceeds the 227 KB of shared memory; the gene -> sets adjacency is stored per X row with
//     per-tile offsets (ptr[row][tile]) and 16-bit tile-local set ids, and is read through L1/L2
//     (it is ~6 MB for 2.7M memberships: L2 resident).
//   * persistent grid (a multiple of the SM count), work items (column, tile) dealt round-robin
//     to the warps of a CTA so all warps of a CTA walk the same column at about the same time
//     (X column and ptr sectors hit L1).
//   * the epilogue (set scaling 1/(n_s+1e-8), scse column scaling, rank zero-group term) is
//     fused into the flush of the tile; output is written once with coalesced streaming stores.
#include "common.cuh"

#include <stdlib.h>

namespace plaidgpu {

template <bool DENSE, bool GENERAL>
__global__ void __launch_bounds__(256, 1) k_score(const ScoreParams p) {
  extern __shared__ double sacc[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int W = blockDim.x >> 5;
  double* __restrict__ acc = sacc + (size_t)w * p.Ts;
  uint2* gq = reinterpret_cast<uint2*>(sacc + (size_t)W * p.Ts) + w * 32;
  double* gx = sacc + (size_t)W * p.Ts + W * 32 + w * 32;
  for (int l = lane; l < p.Ts; l += 32) acc[l] = 0.0;
  __syncwarp();

  const int T = p.T;
  const int64_t ncols_cta = (p.N - blockIdx.x + gridDim.x - 1) / gridDim.x;  // columns of this CTA
  const int64_t nitems = ncols_cta * T;

  for (int64_t q = w; q < nitems; q += W) {
    const int64_t k = q / T;
    const int t = (int)(q - k * T);
    const int64_t j = blockIdx.x + k * (int64_t)gridDim.x;

    int64_t c0, c1;
    if (DENSE) {
      c0 = j * (int64_t)p.P;
      c1 = c0 + p.P;
    } else {
      c0 = p.xp[j];
      c1 = p.xp[j + 1];
    }
    double fb = 0.0;  // f(rank of the zero group): contribution of every implicit zero
    if (GENERAL && p.mode >= XF_SING) fb = xform_value(p.mode, p.r0 ? p.r0[j] : 0.0, p.a0, p.a1);
    const uint32_t* __restrict__ ptr_t = p.ptr + t;
    const int stride = T + 1;
    // Software pipeline over batches of 32 stored entries:
    //   stage 0 (two batches ahead): row index + value of the entry        (coalesced)
    //   stage 1 (one batch ahead)  : ptr[row][t], ptr[row][t+1]            (gather, L1/L2)
    //   stage 2 (current)          : set lists, 8 genes at a time, the next 8 lists' first
    //                                chunks already in flight while 8 are accumulated
    int gi_n = 0;            // batch b+1 after the rotate below: row indices
    double xv_n = 0.0;
    uint32_t q0_c = 0, len_c = 0;
    double xv_c = 0.0;
    auto load_entry = [&](int64_t e, int& gi, double& xv) {
      gi = -1;
      xv = 0.0;
      if (e < c1) {
        gi = DENSE ? (int)(e - c0) : p.xi[e];
        xv = p.xx[e];
      }
    };
    auto load_ptr = [&](int gi, uint32_t& q0, uint32_t& len) {
      q0 = 0;
      len = 0;
      if (gi >= 0) {
        const uint32_t* pp = ptr_t + (size_t)gi * stride;
        q0 = pp[0];
        len = pp[1] - q0;
      }
    };
    {
      int gi0;
      load_entry(c0 + lane, gi0, xv_c);
      load_ptr(gi0, q0_c, len_c);
      load_entry(c0 + 32 + lane, gi_n, xv_n);
    }
    for (int64_t b = c0; b < c1; b += 32) {
      // rotate the pipeline: issue the loads of the next batches before touching this one
      uint32_t q0_n, len_n;
      load_ptr(gi_n, q0_n, len_n);
      int gi_nn;
      double xv_nn;
      load_entry(b + 64 + lane, gi_nn, xv_nn);

      double xv = xv_c;
      if (GENERAL) {
        if (p.mode >= XF_SING) {
          xv = (b + lane < c1) ? xform_value(p.mode, xv, p.a0, p.a1) - fb : 0.0;
        } else {
          xv = (b + lane < c1) ? xform_value(p.mode, xv, p.a0, p.a1) : 0.0;
        }
      }
      if (__ballot_sync(FULL, len_c != 0) != 0) {  // else: no gene of this batch is in a set of this tile
        gq[lane] = make_uint2(q0_c, len_c);
        gx[lane] = xv;
        __syncwarp();
        int lk[8], ln[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint2 g = gq[k];
          lk[k] = (lane < (int)g.y) ? (int)p.idx[g.x + lane] : -1;
        }
#pragma unroll 1
        for (int k0 = 0; k0 < 32; k0 += 8) {
          if (k0 + 8 < 32) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const uint2 g = gq[k0 + 8 + k];
              ln[k] = (lane < (int)g.y) ? (int)p.idx[g.x + lane] : -1;
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const double xk = gx[k0 + k];
            if (lk[k] >= 0) acc[lk[k]] += xk;
            __syncwarp();  // the next gene may hit a set this one just updated
            const uint2 g = gq[k0 + k];
            if (g.y > 32) {  // warp-uniform: long list, remaining chunks 4 at a time (distinct sets)
              const uint32_t s1 = g.x + g.y;
              for (uint32_t eb = g.x + 32; eb < s1; eb += 128) {
                const uint32_t ee = eb + lane;
                const int l0 = (ee < s1) ? (int)p.idx[ee] : -1;
                const int l1 = (ee + 32 < s1) ? (int)p.idx[ee + 32] : -1;
                const int l2 = (ee + 64 < s1) ? (int)p.idx[ee + 64] : -1;
                const int l3 = (ee + 96 < s1) ? (int)p.idx[ee + 96] : -1;
                double v0 = 0, v1 = 0, v2 = 0, v3 = 0;
                if (l0 >= 0) v0 = acc[l0];
                if (l1 >= 0) v1 = acc[l1];
                if (l2 >= 0) v2 = acc[l2];
                if (l3 >= 0) v3 = acc[l3];
                if (l0 >= 0) acc[l0] = v0 + xk;
                if (l1 >= 0) acc[l1] = v1 + xk;
                if (l2 >= 0) acc[l2] = v2 + xk;
                if (l3 >= 0) acc[l3] = v3 + xk;
              }
              __syncwarp();
            }
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) lk[k] = ln[k];
        }
        __syncwarp();  // gq / gx are rewritten by the next batch
      }
      q0_c = q0_n;
      len_c = len_n;
      xv_c = xv_n;
      gi_n = gi_nn;
      xv_n = xv_nn;
    }

    // ---- flush tile t of column j: fused epilogue, coalesced streaming store, re-zero ----
    const int sbase = t * p.Ts;
    const int tl = min(p.Ts, p.S - sbase);
    double* __restrict__ o = p.out + j * p.ld + sbase;
    for (int l = lane; l < tl; l += 32) {
      double v = acc[l];
      acc[l] = 0.0;
      if (p.accumulate) v += __ldcs(o + l);  // partial sums of the gather pass
      if (p.final) v = score_epilogue(v, sbase + l, j, fb, GENERAL ? p.mode : XF_IDENT, p.inv, p.ns, p.colscale);
      __stcs(o + l, v);
    }
    __syncwarp();
  }
}

cudaError_t score_configure(int device, int32_t S, int32_t tile_hint, int32_t* Ts_out, int32_t* T_out,
                            LaunchCfg* cfg) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return e;
  int warps = 8;
  if (const char* w = getenv("PLAIDGPU_WARPS")) {  // tuning knob (bench / profiling only)
    const int v = atoi(w);
    if (v == 1 || v == 2 || v == 4 || v == 8) warps = v;
  }
  const size_t smem_max = (size_t)prop.sharedMemPerBlockOptin - 1024;  // leave the 1 KB reserve
  int32_t ts_max = (int32_t)((smem_max - (size_t)warps * 32 * 16) / (8 * warps));
  ts_max = (ts_max / 32) * 32;
  if (ts_max > 65536) ts_max = 65536;
  int32_t Ts, T;
  if (tile_hint > 0) {
    Ts = ((tile_hint + 31) / 32) * 32;
    if (Ts > ts_max) Ts = ts_max;
    T = (S + Ts - 1) / Ts;
  } else {
    T = (S + ts_max - 1) / ts_max;
    if (T < 1) T = 1;
    Ts = (((S + T - 1) / T) + 31) / 32 * 32;  // even out the tiles
    if (Ts < 32) Ts = 32;
  }
  cfg->warps = warps;
  cfg->smem = (size_t)warps * Ts * sizeof(double) + (size_t)warps * 32 * 16;  // + per-warp staging
  // persistent grid: SM count x resident CTAs per SM
  int per_sm = 0;
  const void* fns[4] = {(const void*)k_score<false, false>, (const void*)k_score<false, true>,
                        (const void*)k_score<true, false>, (const void*)k_score<true, true>};
  for (int i = 0; i < 4; ++i) {
    e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return e;
  }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_score<false, true>, warps * 32, cfg->smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 8) per_sm = 8;
  cfg->ctas = prop.multiProcessorCount * per_sm;
  *Ts_out = Ts;
  *T_out = T;
  return cudaSuccess;
}

cudaError_t launch_score(const ScoreParams& p, bool dense, const LaunchCfg& cfg, cudaStream_t st) {
  if (p.N <= 0) return cudaSuccess;
  const bool general = (p.mode != XF_IDENT);
  int64_t grid = cfg.ctas;
  if (grid > p.N) grid = p.N;
  dim3 g((unsigned)grid), b((unsigned)cfg.warps * 32);
  if (dense) {
    if (general) k_score<true, true><<<g, b, cfg.smem, st>>>(p);
    else k_score<true, false><<<g, b, cfg.smem, st>>>(p);
  } else {
    if (general) k_score<false, true><<<g, b, cfg.smem, st>>>(p);
    else k_score<false, false><<<g, b, cfg.smem, st>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace plaidgpu
