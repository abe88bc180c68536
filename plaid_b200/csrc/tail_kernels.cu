// K3 — the sparse remainder of the gene-set score product (reference R/plaid.R:100-123, Matrix::crossprod ->
// CHOLMOD): every row of a sparse X that is NOT in the tensor-core block (tc_kernels.cu), i.e. the many genes
// that are expressed in a few per cent of the cells and sit in ~100 sets each.
//
// Formulation (replaces the lanes = genes scatter of score_kernels.cu on the fixed-point path):
//   * the cells are cut into tiles of TAIL_C columns; per tile the tail entries of X are regrouped GENE-major
//     (k_tile_scan / k_tile_place: counting sort by tail row -> `rowptr` and 8-byte records `ent` = {value in the
//     column's fixed point (the same 2^e_j the tensor-core pass uses, int32), cell inside the tile});
//   * one warp owns one (tile, set) item at a time: it walks the set's tail members and adds each member's
//     sparse row of the tile into TAIL_C accumulators in shared memory, LANES = the row's ENTRIES.  The cells
//     of one row are distinct, and rows are handled one after the other, so there are no duplicate addresses
//     inside a step: no tags, no retries, no atomics — a step is one 32-bit LDS / IADD / STS per lane;
//   * accumulators are 64-bit integers split into a u32 low word (touched by every add) and an i16 high word
//     (touched only on carry / borrow): random 32-bit accesses cost ~3 shared-memory wavefronts per warp
//     instead of the ~6 of a 64-bit access, and integer sums are exact and order-independent;
//   * the finished row (TAIL_C sums of one set) is written as int64 to `tmp[set][cell]` with coalesced stores;
//     the tensor-core kernel fetches its 128 sets x 48 cells by TMA and adds it to its own integer sums
//     before the one conversion to fp64 (so tail + block is exact, then scaled by 2^-e_j).
// Items are dealt dynamically (one atomic per item) in decreasing order of the set's tail size.
#include "common.cuh"

namespace plaidgpu {

namespace {

constexpr int TAIL_WARPS = 32;

// block-wide exclusive scan of the per-(tile, tail row) entry counts -> tile-local row pointers; the counters
// are zeroed again (k_tile_place re-uses them as cursors).  grid = tiles, 1024 threads.
__global__ void __launch_bounds__(1024) k_tile_scan(uint32_t* __restrict__ cnt, int32_t Pt, uint32_t* __restrict__ rowptr,
                                                    uint32_t* __restrict__ total) {
  __shared__ uint32_t wsum[32];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t* __restrict__ c = cnt + (size_t)t * Pt;
  uint32_t* __restrict__ rp = rowptr + (size_t)t * (Pt + 1);
  const int per = (Pt + 1023) / 1024;
  const int lo = min(Pt, tid * per), hi = min(Pt, lo + per);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  uint32_t incl = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(FULL, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) wsum[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t v = wsum[lane], a = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t u = __shfl_up_sync(FULL, a, o);
      if (lane >= o) a += u;
    }
    wsum[lane] = a - v;  // exclusive
    if (lane == 31) {
      total[t] = a;
      rp[Pt] = a;
    }
  }
  __syncthreads();
  uint32_t run = wsum[w] + incl - s;
  for (int i = lo; i < hi; ++i) {
    const uint32_t v = c[i];
    rp[i] = run;
    run += v;
    c[i] = 0;
  }
}

// one warp per column: tail entries of the (compacted) column -> their slot in the tile's gene-major arrays
__global__ void __launch_bounds__(256) k_tile_place(const int32_t* __restrict__ xp, const int32_t* __restrict__ xe,
                                                    const int32_t* __restrict__ oi, const double* __restrict__ ox,
                                                    const double* __restrict__ r0, const int32_t* __restrict__ tmap,
                                                    const double* __restrict__ colinv, int64_t N, int mode, double a0,
                                                    double a1, int32_t Pt, int C, const uint32_t* __restrict__ rowptr,
                                                    const uint32_t* __restrict__ total, uint32_t* __restrict__ cnt,
                                                    uint2* __restrict__ ent, const int* __restrict__ skip_if) {
  if (skip_if && *skip_if != 0) return;
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = w0; j < N; j += nw) {
    const int t = (int)(j / C);
    const uint32_t cell = (uint32_t)(j - (int64_t)t * C);
    uint32_t base = 0;  // entries of the tiles before this one
    for (int u = lane; u < t; u += 32) base += total[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(FULL, base, o);
    const uint32_t* __restrict__ rp = rowptr + (size_t)t * (Pt + 1);
    uint32_t* __restrict__ cur = cnt + (size_t)t * Pt;
    double fb = 0.0;
    if (mode >= XF_SING) fb = xform_value(mode, r0 ? r0[j] : 0.0, a0, a1);
    const double sc = 1.0 / colinv[j];  // 2^e_j, exact
    // four iterations at a time, each dependent level (index -> tail row -> row pointer + cursor) issued for all four
    // before the next: one warp per column has no other way of overlapping the chain
    const int32_t e1 = xe[j];
    for (int32_t e0 = xp[j]; e0 < e1; e0 += 128) {
      int32_t rr[4], gg[4];
      double xv[4];
      uint32_t pos[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int32_t e = e0 + 32 * u + lane;
        const bool in = e < e1;
        rr[u] = in ? __ldg(oi + e) : -1;
        xv[u] = in ? __ldg(ox + e) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) gg[u] = rr[u] >= 0 ? __ldg(tmap + rr[u]) : -1;
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (gg[u] >= 0) pos[u] = base + rp[gg[u]] + atomicAdd(cur + gg[u], 1u);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (gg[u] < 0) continue;
        double v = xform_value(mode, xv[u], a0, a1);
        if (mode >= XF_SING) v -= fb;
        const int32_t q = (int32_t)__double2ll_rn(v * sc);
        // a non-zero tail entry that the column's fixed point would flush to zero: the call is redone in fp64
        // (tc_kernels.cu, k_tc_prep_dense); every kernel after this one looks at the flag before it starts
        if (q == 0 && v != 0.0 && skip_if) atomicExch(const_cast<int*>(skip_if), 1);
        ent[pos[u]] = make_uint2((uint32_t)q, cell);
      }
    }
  }
}

// fills by kernel, not cudaMemsetAsync: a memset may be scheduled on a copy engine and then waits behind the
// gigabyte result blocks that are leaving for the host on another stream
__global__ void __launch_bounds__(256) k_fill_u32(uint32_t* __restrict__ p, uint32_t v, int64_t n) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

struct TailParams {
  const uint32_t* tptr;    // [S + 1] set -> its tail members
  const uint16_t* tidx;    // tail ids, ascending inside a set
  const int32_t* sorder;   // [S] sets in decreasing order of their tail size
  const uint32_t* rowptr;  // [tiles][Pt + 1]
  const uint32_t* total;   // [tiles]
  const uint2* ent;        // {fixed-point value, cell inside the tile}, gene-major per tile
  int32_t S, Pt, C, tiles;
  long long* tmp;          // [S][ld] int64 sums
  int64_t ld;              // = tiles * C
  unsigned int* counter;   // work counter (zeroed before the launch)
  const int* skip_if;
};

__global__ void __launch_bounds__(TAIL_WARPS * 32, 1) k_tail(const TailParams p) {
  if (p.skip_if && *p.skip_if != 0) return;
  extern __shared__ uint32_t tail_sm[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int C = p.C;
  uint32_t* __restrict__ lo = tail_sm + (size_t)w * C;
  short* __restrict__ hi = reinterpret_cast<short*>(tail_sm + (size_t)TAIL_WARPS * C) + (size_t)w * C;
  for (int c = lane; c < C; c += 32) {
    lo[c] = 0u;
    hi[c] = 0;
  }
  __syncwarp();
  const uint32_t lo_sa = (uint32_t)__cvta_generic_to_shared(lo), hi_sa = (uint32_t)__cvta_generic_to_shared(hi);
  const unsigned nitems = (unsigned)p.tiles * (unsigned)p.S;
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(p.counter, 1u);
    item = __shfl_sync(FULL, item, 0);
    if (item >= nitems) break;
    const int t = (int)(item / (unsigned)p.S);
    const int s = p.sorder[item - (unsigned)t * (unsigned)p.S];
    const uint32_t* __restrict__ rp = p.rowptr + (size_t)t * (p.Pt + 1);
    uint32_t base = 0;
    for (int u = lane; u < t; u += 32) base += p.total[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) base += __shfl_xor_sync(FULL, base, o);
    const uint2* __restrict__ ent = p.ent + base;
    const uint32_t m0 = p.tptr[s], m1 = p.tptr[s + 1];
    for (uint32_t mb = m0; mb < m1; mb += 32) {
      // lane i looks up the row of member mb + i in this tile
      uint32_t r0 = 0, rn = 0;
      if (mb + lane < m1) {
        const uint32_t g = p.tidx[mb + lane];
        r0 = rp[g];
        rn = rp[g + 1] - r0;
      }
      const int cnt = (int)min(32u, m1 - mb);
      // member k's row: segments of 32 entries, lanes = entries.  The first segments of the next four members are
      // in flight while a member is added (entries come from L2, ~350 clocks away; most rows are one segment)
      auto add = [&](const uint2 rec) {
        if (rec.y != 0xFFFFu) {
          const uint32_t a = lo_sa + rec.y * 4u;
          uint32_t old;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(old) : "r"(a) : "memory");
          const uint32_t nv = old + rec.x;
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(nv) : "memory");
          // carry out of the low word vs. sign extension of the addend: equal -> the high word keeps its value
          if ((nv < rec.x) != ((int32_t)rec.x < 0)) {
            const uint32_t ah = hi_sa + rec.y * 2u;
            short hv;
            asm volatile("ld.shared.s16 %0, [%1];" : "=h"(hv) : "r"(ah) : "memory");
            hv = (short)(hv + ((nv < rec.x) ? 1 : -1));
            asm volatile("st.shared.s16 [%0], %1;" ::"r"(ah), "h"(hv) : "memory");
          }
        }
        __syncwarp();
      };
      auto prefetch = [&](int kk, uint2& rec, uint32_t& nk, uint32_t& pk) {
        nk = kk < cnt ? __shfl_sync(FULL, rn, kk & 31) : 0u;
        pk = __shfl_sync(FULL, r0, kk & 31);
        rec = make_uint2(0u, 0xFFFFu);
        if ((uint32_t)lane < nk) rec = __ldg(ent + pk + lane);
      };
      auto seg = [&](uint32_t off, uint32_t nk, uint32_t pk) {
        uint2 r = make_uint2(0u, 0xFFFFu);
        if (off + lane < nk) r = __ldg(ent + pk + off + lane);
        return r;
      };
      auto member = [&](const uint2 rec, uint32_t nk, uint32_t pk) {
        if (nk <= 32u) {  // warp-uniform
          add(rec);
          return;
        }
        // long row (a highly expressed gene in few sets): its entries are one contiguous stream — three segments in
        // flight while three are added, so the L2 latency is paid once per 96 entries instead of once per 32
        uint2 q1 = seg(32u, nk, pk), q2 = seg(64u, nk, pk), q3 = seg(96u, nk, pk);
        add(rec);
        for (uint32_t off = 32; off < nk; off += 96) {
          const uint2 a = q1, b = q2, c = q3;
          q1 = seg(off + 96u, nk, pk);
          q2 = seg(off + 128u, nk, pk);
          q3 = seg(off + 160u, nk, pk);
          add(a);
          if (off + 32u < nk) add(b);
          if (off + 64u < nk) add(c);
        }
      };
      uint2 e0, e1, e2, e3;
      uint32_t n0, n1, n2, n3, p0, p1, p2, p3;
      prefetch(0, e0, n0, p0);
      prefetch(1, e1, n1, p1);
      prefetch(2, e2, n2, p2);
      prefetch(3, e3, n3, p3);
      for (int k = 0; k < cnt; k += 4) {
        member(e0, n0, p0);
        prefetch(k + 4, e0, n0, p0);
        member(e1, n1, p1);
        prefetch(k + 5, e1, n1, p1);
        member(e2, n2, p2);
        prefetch(k + 6, e2, n2, p2);
        member(e3, n3, p3);
        prefetch(k + 7, e3, n3, p3);
      }
    }
    // flush: int64 sums of set s over the tile's cells, coalesced; re-zero
    long long* __restrict__ o = p.tmp + (size_t)s * p.ld + (size_t)t * C;
    for (int c = lane; c < C; c += 32) {
      const long long v = ((long long)hi[c] << 32) | (long long)lo[c];
      lo[c] = 0u;
      hi[c] = 0;
      __stcs(o + c, v);
    }
    __syncwarp();
  }
}

}  // namespace

cudaError_t launch_fill_u32(void* p, uint32_t v, int64_t n, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int64_t grid = (n + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  k_fill_u32<<<(unsigned)grid, 256, 0, st>>>(static_cast<uint32_t*>(p), v, n);
  return cudaGetLastError();
}

int tail_tile_cells() { return 1056; }  // 22 tensor-core cell tiles of 48

size_t tail_smem_bytes() { return (size_t)TAIL_WARPS * tail_tile_cells() * 6; }

cudaError_t launch_tile_scan(uint32_t* cnt, int32_t Pt, int tiles, uint32_t* rowptr, uint32_t* total, cudaStream_t st) {
  if (tiles <= 0) return cudaSuccess;
  k_tile_scan<<<(unsigned)tiles, 1024, 0, st>>>(cnt, Pt, rowptr, total);
  return cudaGetLastError();
}

cudaError_t launch_tile_place(const int32_t* xp, const int32_t* xe, const int32_t* oi, const double* ox, const double* r0,
                              const int32_t* tmap, const double* colinv, int64_t N, int mode, double a0, double a1,
                              int32_t Pt, const uint32_t* rowptr, const uint32_t* total, uint32_t* cnt, uint2* ent,
                              const int* skip_if, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t grid = (N + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  k_tile_place<<<(unsigned)grid, 256, 0, st>>>(xp, xe, oi, ox, r0, tmap, colinv, N, mode, a0, a1, Pt, tail_tile_cells(),
                                               rowptr, total, cnt, ent, skip_if);
  return cudaGetLastError();
}

cudaError_t launch_tail(const uint32_t* tptr, const uint16_t* tidx, const int32_t* sorder, const uint32_t* rowptr,
                        const uint32_t* total, const uint2* ent, int32_t S, int32_t Pt, int tiles,
                        long long* tmp, unsigned int* counter, const int* skip_if, cudaStream_t st) {
  if (tiles <= 0 || S <= 0) return cudaSuccess;
  TailParams p{};
  p.tptr = tptr;
  p.tidx = tidx;
  p.sorder = sorder;
  p.rowptr = rowptr;
  p.total = total;
  p.ent = ent;
  p.S = S;
  p.Pt = Pt;
  p.C = tail_tile_cells();
  p.tiles = tiles;
  p.tmp = tmp;
  p.ld = (int64_t)tiles * p.C;
  p.counter = counter;
  p.skip_if = skip_if;
  const size_t smem = tail_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(k_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = launch_fill_u32(counter, 0u, 1, st);
  if (e != cudaSuccess) return e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  k_tail<<<(unsigned)sms, TAIL_WARPS * 32, smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace plaidgpu
