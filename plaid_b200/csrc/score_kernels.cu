// K1/K2 — the gene-set score product  out = t(G) %*% X  (scaled per set), the hot loop of
// plaid() / chunked_crossprod() (reference R/plaid.R:60-87, 100-123; Matrix::crossprod ->
// CHOLMOD ssmult/sdmult).
//
// Formulation (B200-first, not a translation of Gustavson-on-CPU):
//   * scatter form: for every stored x(g, j) add it to all sets containing gene g.  Work is
//     nnz(X) x avg-degree adds, the work-optimal count for sparse X sparse G -> dense out.
//   * accumulators are fp64 and live in SHARED MEMORY; one warp owns one tile of Ts sets of
//     one column at a time (no other warp touches it), lanes = genes (see k_scatter).
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

// One warp owns one accumulator tile (Ts sets of one column) at a time.  LANES = GENES: every lane
// streams the tile-local set list of ONE stored entry (gene) of the column, one set per step, so 32
// read-modify-writes of 32 different genes are in flight per step no matter how short the individual
// lists are.  Two lanes may hit the same set in the same step (two genes of one set): each lane
// first writes its lane id into a per-set byte tag and only the lane that reads its own id back does
// the read-modify-write; the others retry (about one step in seven needs a second round).
//
// Work distribution inside the warp is a queue: a batch of 32 entries of the column is staged in
// shared memory (list range, value, first 8-byte chunk of the list; empty lists are dropped), a lane
// that finished its list takes the next record.  Batches are prepared three rotations ahead in
// registers (entry -> row pointers -> first chunk), so no global-load latency sits on the refill
// path.  Lists are padded to multiples of 4 entries (0xFFFF): all lanes cross chunk boundaries
// together and refills happen only there.

constexpr int STAGE_BYTES = 1280;  // per warp: 32 staged genes x {hdr, x, 3 chunks} of 8 bytes

// one 32-byte tile record with a single 256-bit load (LDG.E.256 on sm_100a)
__device__ __forceinline__ void ld_rec32(const void* p, unsigned long long& a, unsigned long long& b,
                                         unsigned long long& c, unsigned long long& d) {
  asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}

template <bool GENERAL>
__global__ void __launch_bounds__(768, 1) k_scatter(const ScoreParams p) {
  if (p.run_if && *p.run_if == 0) return;
  extern __shared__ double sacc[];
  const int lane = threadIdx.x & 31;
  const int w = threadIdx.x >> 5;
  const int W = blockDim.x >> 5;
  const int tagw = (p.Ts + 15) & ~15;
  double* __restrict__ acc = sacc + (size_t)w * p.Ts;
  char* __restrict__ accb = reinterpret_cast<char*>(acc);
  unsigned char* __restrict__ tag = reinterpret_cast<unsigned char*>(sacc + (size_t)W * p.Ts) + (size_t)w * tagw;
  // staged batch, structure-of-arrays (8-byte columns: conflict-free stores, ~1 wavefront per gathered read)
  unsigned char* __restrict__ recb = reinterpret_cast<unsigned char*>(sacc + (size_t)W * p.Ts) + (size_t)W * tagw +
                                     (size_t)w * STAGE_BYTES;
  uint2* __restrict__ rec_hdr = reinterpret_cast<uint2*>(recb);          // {overflow offset, chunks}
  double* __restrict__ rec_x = reinterpret_cast<double*>(recb + 256);
  uint2* __restrict__ rec_c0 = reinterpret_cast<uint2*>(recb + 512);
  uint2* __restrict__ rec_c1 = reinterpret_cast<uint2*>(recb + 768);
  uint2* __restrict__ rec_c2 = reinterpret_cast<uint2*>(recb + 1024);
  for (int l = lane; l < p.Ts; l += 32) acc[l] = 0.0;
  __syncwarp();
  const unsigned lt = (1u << lane) - 1u;
  double vmin = INFINITY;  // smallest final score this thread wrote (p.smin)

  const int T = p.T;
  const int64_t ncols_cta = (p.N - blockIdx.x + gridDim.x - 1) / gridDim.x;  // columns of this CTA
  const int64_t nitems = ncols_cta * T;

  for (int64_t q = w; q < nitems; q += W) {
    const int64_t k = q / T;
    const int t = (int)(q - k * T);
    const int64_t j = blockIdx.x + k * (int64_t)gridDim.x;
    const int64_t c0 = p.xp[j], c1 = p.xe ? p.xe[j] : p.xp[j + 1];
    double fb = 0.0;  // f(rank of the zero group): contribution of every implicit zero
    if (GENERAL && p.mode >= XF_SING) fb = xform_value(p.mode, p.r0 ? p.r0[j] : 0.0, p.a0, p.a1);
    // ---- batch pipeline (lane i prepares entry i of a batch) ---------------------------------
    // adjacency format: one 32-byte record (= one DRAM/L2 sector) per (X row, tile) = {overflow offset,
    // list length, 12 inline entries}; longer lists continue in the overflow array in chunks of 4
    // (padded with 0xFFFF).  One scattered 256-bit load yields the pointer AND the first three chunks of
    // a (gene, tile) list - most lists are complete with it.
    const unsigned char* __restrict__ trec = reinterpret_cast<const unsigned char*>(p.ptr) + (size_t)t * 32;
    int64_t ebase = c0;          // first entry of the batch whose (row, value) load is issued next
    int gi1 = -1;                // R1: row index + value loaded
    double x1 = 0.0;
    uint32_t rest2 = 0, nch2 = 0;  // R2: tile record loaded -> ready to be staged
    uint2 c02 = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu), c12 = c02, c22 = c02;
    double x2 = 0.0;
    int qhead = 0, qcount = 0;   // staged batch: records [qhead, qcount) are unassigned (warp-uniform)

    auto rotate = [&]() {
      const bool have = nch2 > 0;
      const unsigned m = __ballot_sync(FULL, have);
      if (have) {
        const int slot = __popc(m & lt);
        rec_hdr[slot] = make_uint2(rest2, nch2);
        rec_x[slot] = x2;
        rec_c0[slot] = c02;
        rec_c1[slot] = c12;
        rec_c2[slot] = c22;
      }
      qhead = 0;
      qcount = __popc(m);
      __syncwarp();
      nch2 = 0;
      if (gi1 >= 0) {
        unsigned long long w0, w1, w2, w3;
        ld_rec32(trec + (size_t)gi1 * T * 32, w0, w1, w2, w3);
        rest2 = (uint32_t)w0;
        nch2 = ((uint32_t)(w0 >> 32) & 0xFFFFu) + 3u >> 2;
        c02 = make_uint2((uint32_t)w1, (uint32_t)(w1 >> 32));
        c12 = make_uint2((uint32_t)w2, (uint32_t)(w2 >> 32));
        c22 = make_uint2((uint32_t)w3, (uint32_t)(w3 >> 32));
        x2 = x1;
        if (GENERAL) x2 = xform_value(p.mode, x1, p.a0, p.a1) - (p.mode >= XF_SING ? fb : 0.0);
      }
      gi1 = -1;
      if (ebase + lane < c1) {
        gi1 = p.xi[ebase + lane];
        x1 = p.xx[ebase + lane];
      }
      ebase += 32;
    };
    // prime: after two rotations the first batch is in R2; the third stages it
    rotate();
    rotate();
    int remaining = (int)((c1 - c0 + 31) >> 5);  // batches still to be staged (one per further rotation)
    uint32_t ck = 0, nck = 0, rest = 0;  // this lane's active list: chunk ck of nck
    double x = 0.0;
    uint2 buf = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu), nbuf = buf, nnbuf = buf;

    for (;;) {
      // ---- chunk boundary (every 4 steps, all lanes together): advance chunk / take a new gene ----
      ++ck;
      bool need = ck >= nck;
      bool fresh = false;
      if (!need) {
        buf = nbuf;
        nbuf = nnbuf;
      }
      for (;;) {
        const unsigned nm = __ballot_sync(FULL, need);
        if (nm == 0) break;
        if (qhead >= qcount) {
          if (remaining == 0) break;
          rotate();
          --remaining;
          continue;
        }
        const int pos = qhead + __popc(nm & lt);
        if (need && pos < qcount) {
          const uint2 hd = rec_hdr[pos];
          rest = hd.x; nck = hd.y; ck = 0; x = rec_x[pos]; buf = rec_c0[pos]; nbuf = rec_c1[pos]; nnbuf = rec_c2[pos];
          need = false;
          fresh = true;
        }
        qhead += __popc(nm);
        __syncwarp();
      }
      const bool alive = ck < nck;
      if (!__any_sync(FULL, alive)) break;  // queue and pipeline are empty too (loop above ran dry)
      // chunks 3.. live in the overflow array and are fetched two boundaries ahead of their use
      if (alive && !fresh && ck + 2 < nck) nnbuf = *reinterpret_cast<const uint2*>(p.idx + rest + 4 * (ck - 1));
#pragma unroll
      for (int s4 = 0; s4 < 4; ++s4) {
        // branch-free fast path: inactive lanes are pointed at set 0 and simply never win
        const unsigned raw = ((s4 < 2 ? buf.x : buf.y) >> (16 * (s4 & 1))) & 0xFFFFu;
        bool pend = alive && raw != 0xFFFFu;
        const unsigned off = pend ? raw : 0u;
        unsigned char* __restrict__ tg = tag + (off >> 3);
        volatile double* a = reinterpret_cast<volatile double*>(accb + off);
        if (pend) *tg = (unsigned char)lane;
        __syncwarp();  // every lane's tag store is performed before any lane's tag load
        const unsigned tv = *reinterpret_cast<volatile unsigned char*>(tg);
        const double v0 = *a;
        const bool win = pend && tv == (unsigned)lane;
        if (win) *a = v0 + x;
        pend = pend && !win;
        while (__any_sync(FULL, pend)) {  // rare: two genes of this step share a set
          if (pend) *tg = (unsigned char)lane;
          __syncwarp();
          const bool w2 = pend && (*reinterpret_cast<volatile unsigned char*>(tg) == (unsigned char)lane);
          if (w2) *a = *a + x;
          pend = pend && !w2;
          __syncwarp();
        }
      }
    }

    // ---- flush tile t of column j: fused epilogue, coalesced streaming store, re-zero ----
    // (4 rows per lane at a time with the global loads issued first: the partial sums of the gather
    //  pass and the per-set scales would otherwise each expose a DRAM / L2 round trip)
    const int sbase = t * p.Ts;
    const int tl = min(p.Ts, p.S - sbase);
    double* __restrict__ o = p.out + j * p.ld + sbase;
    const double* __restrict__ invt = p.inv + sbase;
    const double* __restrict__ nst = p.ns + sbase;
    const bool rankmode = GENERAL && p.mode >= XF_SING;
    const double csj = (p.final && p.colscale) ? p.colscale[j] : 1.0;
    for (int l0 = lane; l0 < tl; l0 += 128) {
      double part[4], iv[4], nv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = l0 + 32 * u;
        part[u] = 0.0; iv[u] = 1.0; nv[u] = 0.0;
        if (l < tl) {
          if (p.accumulate) part[u] = __ldcs(o + l);
          if (p.final) {
            iv[u] = invt[l];
            if (rankmode) nv[u] = nst[l];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int l = l0 + 32 * u;
        if (l < tl) {
          double v = acc[l] + part[u];
          acc[l] = 0.0;
          if (p.final) {
            if (rankmode) v += fb * nv[u];
            v = v * iv[u] * csj;
            vmin = fmin(vmin, v);
          }
          __stcs(o + l, v);
        }
      }
    }
    __syncwarp();
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

// Drop the entries of the gather block from every column once (they are a third of a single-cell column
// and would otherwise be fetched and discarded by each of the T tile walks).  Compacted entries keep
// their column's start offset; xe[j] = new end of column j.
__global__ void __launch_bounds__(256) k_compact(const int32_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                                 const double* __restrict__ xx, const uint16_t* __restrict__ dmap,
                                                 int64_t N, int32_t* __restrict__ oi, double* __restrict__ ox,
                                                 int32_t* __restrict__ xe) {
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lt = (1u << lane) - 1u;
  for (int64_t j = w0; j < N; j += nw) {
    const int32_t c0 = xp[j], c1 = xp[j + 1];
    int32_t o = c0;
    for (int32_t b = c0; b < c1; b += 32) {
      const int32_t e = b + lane;
      int32_t r = 0;
      double v = 0.0;
      bool keep = false;
      if (e < c1) {
        r = xi[e];
        v = xx[e];
        keep = dmap[r] == 0xFFFFu;
      }
      const unsigned m = __ballot_sync(FULL, keep);
      if (keep) {
        const int32_t q = o + __popc(m & lt);
        oi[q] = r;
        ox[q] = v;
      }
      o += __popc(m);
    }
    if (lane == 0) xe[j] = o;
  }
}

cudaError_t launch_compact(const int32_t* xp, const int32_t* xi, const double* xx, const uint16_t* dmap, int64_t N,
                           int32_t* oi, double* ox, int32_t* xe, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t grid = (N + 7) / 8;
  if (grid > 148 * 16) grid = 148 * 16;
  k_compact<<<(unsigned)grid, 256, 0, st>>>(xp, xi, xx, dmap, N, oi, ox, xe);
  return cudaGetLastError();
}

cudaError_t score_configure(int device, int32_t S, int32_t tile_hint, int32_t* Ts_out, int32_t* T_out,
                            LaunchCfg* cfg) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return e;
  int warps = 20;  // measured: 16 -> 20.07 ms, 20 -> 19.52, 24 -> 20.56 (16,384 cells, C4-shaped)
  if (const char* w = getenv("PLAIDGPU_WARPS")) {  // tuning knob (bench / profiling only)
    const int v = atoi(w);
    if (v >= 1 && v <= 24) warps = v;
  }
  const size_t smem_max = (size_t)prop.sharedMemPerBlockOptin - 1024;  // leave the 1 KB reserve
  // per warp: Ts fp64 accumulators + Ts byte tags (16-byte aligned) + 32 staged genes of 40 B
  int32_t ts_max = (int32_t)((smem_max / warps - STAGE_BYTES - 16) / 9);
  if (ts_max > 8160) ts_max = 8160;  // tile-local byte offsets are 16-bit (0xFFFF = padding)
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
  cfg->smem = (size_t)warps * ((size_t)Ts * sizeof(double) + (size_t)((Ts + 15) & ~15) + STAGE_BYTES);
  // persistent grid: SM count x resident CTAs per SM
  int per_sm = 0;
  const void* fns[2] = {(const void*)k_scatter<false>, (const void*)k_scatter<true>};
  for (int i = 0; i < 2; ++i) {
    e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return e;
  }
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_scatter<true>, warps * 32, cfg->smem);
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
  if (dense) return cudaErrorInvalidValue;  // dense X goes through the gather passes only
  int64_t grid = cfg.ctas;
  if (grid > p.N) grid = p.N;
  dim3 g((unsigned)grid), b((unsigned)cfg.warps * 32);
  if (p.mode != XF_IDENT) k_scatter<true><<<g, b, cfg.smem, st>>>(p);
  else k_scatter<false><<<g, b, cfg.smem, st>>>(p);
  return cudaGetLastError();
}

}  // namespace plaidgpu
