// K3/K4 — column ranks with ties, the ranking path of colranks() / sparse_colranks()
// (reference R/plaid.R:589-650; base::rank, sparseMatrixStats::colRanks, matrixStats::colRanks).
//
// One CTA sorts one column: order-preserving 64-bit keys + original positions live in shared
// memory (global workspace only when a column does not fit), sorted by an all-ascending bitonic
// network over the next power of two with VIRTUAL +inf padding (pairs whose upper index is >= n
// are skipped, which is exact because every compare-exchange moves the minimum down).  Ranks
// come from the tie run [first, last] of each key found by binary search in the sorted column:
// average = (first + last + 1) / 2, min = first + 1, max = last  (exact multiples of 0.5).
// The implicit zeros of a CSC column never get materialised: they form one tie group whose
// size is P - nnz, which only shifts the ranks of the positive entries.
#include "common.cuh"

#include <stdlib.h>

#include <math.h>

namespace plaidgpu {

namespace {

constexpr int RT_SHORT = 256;  // threads per column CTA, short columns
constexpr int RT_MAX = 1024;  // long columns (one CTA per SM by shared memory): more warps to hide the latencies
constexpr unsigned long long ZERO_KEY = 0x8000000000000000ull;
constexpr unsigned long long NAN_KEY = ~0ull;

// stable: equal keys are ordered by position (ties.method first / last / dense need the order of appearance)
template <typename PosT>
__device__ __forceinline__ void cas(unsigned long long* k, PosT* p, int i, int l, bool stable) {
  const unsigned long long a = k[i], b = k[l];
  if (a > b || (stable && a == b && p[i] > p[l])) {
    k[i] = b;
    k[l] = a;
    const PosT t = p[i];
    p[i] = p[l];
    p[l] = t;
  }
}

template <typename PosT>
__device__ void bitonic_sort(unsigned long long* k, PosT* p, int n, bool stable = false) {
  int np2 = 1;
  while (np2 < n) np2 <<= 1;
  const int half = np2 >> 1;
  for (int size = 2; size <= np2; size <<= 1) {
    const int hs = size >> 1;
    for (int q = threadIdx.x; q < half; q += blockDim.x) {  // flip step
      const int blk = q / hs, o = q - blk * hs;
      const int i = blk * size + o, l = blk * size + size - 1 - o;
      if (l < n) cas(k, p, i, l, stable);
    }
    __syncthreads();
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      for (int q = threadIdx.x; q < half; q += blockDim.x) {
        const int i = 2 * stride * (q / stride) + (q % stride), l = i + stride;
        if (l < n) cas(k, p, i, l, stable);
      }
      __syncthreads();
    }
  }
}

__device__ __forceinline__ int lower_bound(const unsigned long long* k, int n, unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (k[mid] < v) lo = mid + 1; else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int upper_bound(const unsigned long long* k, int n, unsigned long long v) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (k[mid] <= v) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// ---- tie-aware fast path -------------------------------------------------------------------
// Single-cell columns are made of a few dozen distinct values (log-normalised counts: 95.8 % of the
// non-zeros of the reference fixture are tied).  Ranks then follow from COUNTING instead of sorting:
// the distinct keys of the column are collected in a shared-memory hash table with their
// multiplicities, only those (<= RANK_DCAP) are sorted, a prefix sum over the multiplicities gives
// every value its tie run [first, last), and every entry looks its rank up.  Columns with more
// distinct values fall through to the bitonic sort below; both paths produce the same integers.
// hash slots: every thread may insert one new key after the distinct-value limit was last seen below the cap, so the
// table must hold RANK_DCAP + (threads per CTA) keys with room to probe: 1,024 slots for 256 threads, 2,048 for 1,024
template <bool WIDE> struct RankHt { static constexpr int value = WIDE ? 2048 : 1024; };
constexpr int RANK_DCAP = 448;   // distinct values the fast path accepts (table at most ~70 % full while racing)
constexpr int RANK_DMAX = 512;
template <int RANK_HT>
struct RankFast {
  unsigned long long tab[RANK_HT];
  unsigned cnt[RANK_HT];
  double rk[RANK_HT];
  unsigned long long dk[RANK_DMAX];
  unsigned short ds[RANK_DMAX];
  int first[RANK_DMAX];
  int ndist, m;
};
template <int RANK_HT>
__device__ __forceinline__ unsigned rank_hash(unsigned long long key) {
  return (unsigned)((key * 0x9E3779B97F4A7C15ull) >> (RANK_HT == 2048 ? 53 : 54));  // top 11 / 10 bits
}

struct RankParams {
  const int32_t* xp;  // nullptr: dense column-major input
  const double* xx;
  int32_t P;
  int64_t N;
  int ties, is_signed, dense_sem;
  double* rank;
  double* r0;
  double* colmax;
  unsigned long long* ws_keys;  // global workspace (GLOBAL_WS only)
  void* ws_pos;
  int cap;  // shared-memory capacity in elements (!GLOBAL_WS)
  int ss_max;       // largest splitter count the bucket region holds (0: bucket path off), power of two
  size_t bucket_off;  // byte offset of the bucket region behind keys + positions
};

// ---- bucket path (columns with many distinct values) ----------------------------------------------------------
// Ranks need "how many keys are smaller / equal", not a sorted array.  A strided sample of SS keys is sorted and
// used as splitters; every key falls either ON a splitter value (an "equal class": all its members are tied, the
// rank follows from counts alone) or strictly BETWEEN two splitters (an "interval bucket" of ~n / SS keys).  One
// counting pass + an exclusive scan give each class its first position; the members of every interval bucket are
// listed (unordered: atomics) and each key counts the smaller / equal keys inside its own bucket — a segmented
// MSD pass with data-dependent splitters followed by direct ranking, ~3 binary searches + ~n / SS compares per key
// instead of the log^2(n) / 2 compare-exchange stages of a bitonic network.  Large tie classes are caught by the
// sample (a class holding 1 % of the column is missed by 1,024 samples with probability 3e-5); a column whose
// largest interval bucket still exceeds BUCKET_MAX is sorted by the network instead.
constexpr int BUCKET_MIN_N = 256;   // shorter columns: the network is cheap
constexpr int BUCKET_MAX = 1024;    // largest interval bucket ranked directly

__device__ void sort_keys_pow2(unsigned long long* k, int n) {  // all threads; n = power of two
  for (int size = 2; size <= n; size <<= 1) {
    const int hs = size >> 1;
    for (int q = threadIdx.x; q < n / 2; q += blockDim.x) {
      const int o = q & (hs - 1), base = (q & ~(hs - 1)) << 1;
      const int i = base + o, l = base + size - 1 - o;
      const unsigned long long a = k[i], b = k[l];
      if (a > b) { k[i] = b; k[l] = a; }
    }
    __syncthreads();
    for (int stride = size >> 2; stride >= 1; stride >>= 1) {
      for (int q = threadIdx.x; q < n / 2; q += blockDim.x) {
        const int i = ((q & ~(stride - 1)) << 1) | (q & (stride - 1)), l = i + stride;
        const unsigned long long a = k[i], b = k[l];
        if (a > b) { k[i] = b; k[l] = a; }
      }
      __syncthreads();
    }
  }
}

template <typename PosT, bool GLOBAL_WS, bool WIDE>
__global__ void __launch_bounds__(WIDE ? RT_MAX : RT_SHORT) k_rank(const RankParams p) {
  constexpr int RT = WIDE ? RT_MAX : RT_SHORT;  // 256 (several short columns per SM) or 1024 (one long column per SM)
  extern __shared__ unsigned long long rsm[];
  __shared__ int s_nnan, s_bmax, s_zneg, s_zs;
  __shared__ double s_max[RT / 32];
  // the fast path's tables share the dynamic buffer with the sort path's keys / positions (it is done, or has
  // given up, before those are written)
  constexpr int RANK_HT = RankHt<WIDE>::value;
  RankFast<RANK_HT>& sf = *reinterpret_cast<RankFast<RANK_HT>*>(rsm);
  const int tid = threadIdx.x;

  for (int64_t j = blockIdx.x; j < p.N; j += gridDim.x) {
    int64_t c0;
    int n;
    if (p.xp) {
      c0 = p.xp[j];
      n = p.xp[j + 1] - p.xp[j];
    } else {
      c0 = j * (int64_t)p.P;
      n = p.P;
    }
    unsigned long long* keys;
    PosT* pos;
    if (GLOBAL_WS) {
      keys = p.ws_keys + c0;
      pos = reinterpret_cast<PosT*>(p.ws_pos) + c0;
    } else {
      keys = rsm;
      pos = reinterpret_cast<PosT*>(rsm + p.cap);
    }
    if (tid == 0) s_nnan = 0;
    __syncthreads();
    // ---- fast path: count the distinct values instead of sorting the column ----
    bool fast_done = false;
    double mymax = 0.0;
    int nneg_f = 0, z_f = 0;
    const bool ordered = p.ties > PLAIDGPU_TIES_MAX;  // first / last / dense: stable network, order of appearance
    if (!GLOBAL_WS && !ordered) {
      for (int i = tid; i < RANK_HT; i += RT) {
        sf.tab[i] = NAN_KEY;  // empty (NaN entries never enter the table)
        sf.cnt[i] = 0u;
      }
      if (tid == 0) sf.ndist = sf.m = 0;
      __syncthreads();
      int my_nan = 0;
      for (int l = tid; l < n; l += RT) {
        const double v = p.xx[c0 + l];
        if (v != v) {
          ++my_nan;
          continue;
        }
        if (*reinterpret_cast<volatile int*>(&sf.ndist) > RANK_DCAP) break;  // too many distinct values: give up early
        const unsigned long long key = key_of(p.is_signed ? fabs(v) : v);
        unsigned h = rank_hash<RANK_HT>(key);
        for (;;) {
          const unsigned long long old = atomicCAS(&sf.tab[h], NAN_KEY, key);
          if (old == NAN_KEY) atomicAdd(&sf.ndist, 1);
          if (old == NAN_KEY || old == key) {
            atomicAdd(&sf.cnt[h], 1u);
            break;
          }
          h = (h + 1) & (RANK_HT - 1);
        }
      }
      if (my_nan) atomicAdd(&s_nnan, my_nan);
      __syncthreads();
      if (sf.ndist <= RANK_DCAP) {  // uniform: read after the barrier
        for (int i = tid; i < RANK_HT; i += RT)
          if (sf.tab[i] != NAN_KEY) {
            const int q = atomicAdd(&sf.m, 1);
            sf.dk[q] = sf.tab[i];
            sf.ds[q] = (unsigned short)i;
          }
        __syncthreads();
        const int m = sf.m;
        bitonic_sort<unsigned short>(sf.dk, sf.ds, m);
        if (tid < 32) {  // exclusive prefix sum of the multiplicities in value order (one warp)
          const int per = (m + 31) / 32, a = tid * per, b = min(m, a + per);
          int sum = 0;
          for (int i = a; i < b; ++i) sum += (int)sf.cnt[sf.ds[i]];
          int incl = sum;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(FULL, incl, o);
            if (tid >= o) incl += t;
          }
          int run = incl - sum;
          for (int i = a; i < b; ++i) {
            sf.first[i] = run;
            run += (int)sf.cnt[sf.ds[i]];
          }
        }
        __syncthreads();
        const int nvf = n - s_nnan;
        int zs = 0, zimp = 0;
        if (p.dense_sem) {
          const int iz = lower_bound(sf.dk, m, ZERO_KEY);
          nneg_f = iz < m ? sf.first[iz] : nvf;
          zs = (iz < m && sf.dk[iz] == ZERO_KEY) ? (int)sf.cnt[sf.ds[iz]] : 0;
          zimp = p.P - n;
        }
        z_f = zs + zimp;
        for (int i = tid; i < m; i += RT) {
          const unsigned long long key = sf.dk[i];
          int first = sf.first[i], last = first + (int)sf.cnt[sf.ds[i]];
          if (p.dense_sem && key == ZERO_KEY) {
            first = nneg_f;
            last = nneg_f + z_f;
          } else if (p.dense_sem && key > ZERO_KEY) {
            first += zimp;
            last += zimp;
          }
          sf.rk[sf.ds[i]] = p.ties == PLAIDGPU_TIES_AVERAGE ? 0.5 * (double)(first + 1 + last)
                            : p.ties == PLAIDGPU_TIES_MIN   ? (double)(first + 1)
                                                            : (double)last;
        }
        __syncthreads();
        for (int l = tid; l < n; l += RT) {
          const double v = p.xx[c0 + l];
          double r;
          if (v != v) {
            r = nan("");
          } else {
            const unsigned long long key = key_of(p.is_signed ? fabs(v) : v);
            unsigned h = rank_hash<RANK_HT>(key);
            while (sf.tab[h] != key) h = (h + 1) & (RANK_HT - 1);
            r = sf.rk[h];
            if (p.is_signed) r = v > 0.0 ? r : (v < 0.0 ? -r : 0.0);
            mymax = fmax(mymax, fabs(r));
          }
          p.rank[c0 + l] = r;
        }
        fast_done = true;
      }
      __syncthreads();
      if (!fast_done && tid == 0) s_nnan = 0;  // the sort path counts the NaNs again
      __syncthreads();
    }
    int nneg = nneg_f, z = z_f;
    if (!fast_done) {
    int my_nan = 0;
    for (int l = tid; l < n; l += RT) {
      double v = p.xx[c0 + l];
      unsigned long long key;
      if (v != v) {
        key = NAN_KEY;
        ++my_nan;
      } else {
        key = key_of(p.is_signed ? fabs(v) : v);
      }
      keys[l] = key;
      pos[l] = (PosT)l;
    }
    if (my_nan) atomicAdd(&s_nnan, my_nan);
    __syncthreads();
    bool bucket_done = false;
    if (!GLOBAL_WS && !ordered && p.ss_max >= 32 && n >= BUCKET_MIN_N) {
      // ---- bucket path: splitters from a sorted sample, counting, direct ranking inside the interval buckets ----
      int SS = 32;
      // keys per interval bucket: 8-16 where the CTA owns the SM anyway, 32-64 for short columns (a larger bucket
      // region would cost the common tied-column path its occupancy)
      while (SS * (WIDE ? 8 : 32) <= n && SS < p.ss_max) SS <<= 1;
      unsigned long long* const sp = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(rsm) + p.bucket_off);
      unsigned* const cnt = reinterpret_cast<unsigned*>(sp + SS);  // code = 2 * b + eq, b = #splitters below the key
      unsigned* const base = cnt + (2 * SS + 2);
      unsigned* const part = base + (2 * SS + 2);                   // RT partial sums
      const int ncode = 2 * SS + 1;
      for (int i = tid; i < SS; i += RT) sp[i] = keys[(int)(((int64_t)i * n) / SS)];
      for (int i = tid; i < ncode + 1; i += RT) cnt[i] = 0u;
      if (tid == 0) s_bmax = 0;
      __syncthreads();
      sort_keys_pow2(sp, SS);
      auto code_of = [&](unsigned long long key) {
        const int b = lower_bound(sp, SS, key);
        return 2 * b + ((b < SS && sp[b] == key) ? 1 : 0);
      };
      for (int l = tid; l < n; l += RT) {
        const unsigned long long key = keys[l];
        if (key != NAN_KEY) atomicAdd(&cnt[code_of(key)], 1u);
      }
      __syncthreads();
      {  // exclusive scan of the class sizes in value order; cnt becomes the cursors of the interval buckets
        const int per = (ncode + RT - 1) / RT, a = min(ncode, tid * per), b = min(ncode, a + per);
        unsigned sum = 0, bm = 0;
        for (int i = a; i < b; ++i) {
          sum += cnt[i];
          if (!(i & 1)) bm = max(bm, cnt[i]);
        }
        part[tid] = sum;
        if (bm) atomicMax(&s_bmax, (int)bm);
        __syncthreads();
        if (tid < 32) {
          unsigned loc = 0;
          for (int i = 0; i < RT / 32; ++i) loc += part[tid * (RT / 32) + i];
          unsigned incl = loc;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(FULL, incl, o);
            if (tid >= o) incl += t;
          }
          unsigned run = incl - loc;
          for (int i = 0; i < RT / 32; ++i) {
            const unsigned t = part[tid * (RT / 32) + i];
            part[tid * (RT / 32) + i] = run;
            run += t;
          }
        }
        __syncthreads();
        unsigned run = part[tid];
        for (int i = a; i < b; ++i) {
          const unsigned t = cnt[i];
          base[i] = run;
          cnt[i] = 0u;
          run += t;
        }
        if (tid == RT - 1) base[ncode] = run;  // == number of valid keys (threads past the end carry the total)
      }
      __syncthreads();
      if (s_bmax <= BUCKET_MAX) {  // block-uniform
        for (int l = tid; l < n; l += RT) {  // member lists of all classes (order inside a class is irrelevant)
          const unsigned long long key = keys[l];
          if (key == NAN_KEY) {
            p.rank[c0 + l] = nan("");
            continue;
          }
          const int c = code_of(key);
          pos[base[c] + atomicAdd(&cnt[c], 1u)] = (PosT)l;
        }
        const int nv = n - s_nnan;
        int zimp = 0;
        if (p.dense_sem) {  // zero group (dense semantics): stored zeros + implicit zeros
          if (tid == 0) {
            const int c = code_of(ZERO_KEY);
            if (c & 1) {
              s_zneg = (int)base[c];
              s_zs = (int)(base[c + 1] - base[c]);
            } else {
              s_zneg = -1 - c;  // no stored zero on a splitter: counted inside interval bucket c below
              s_zs = 0;
            }
          }
          zimp = p.P - n;
        }
        __syncthreads();
        if (p.dense_sem && s_zneg < 0) {
          const int c = -1 - s_zneg;
          __syncthreads();
          if (tid < 32) {
            int less = 0, eq = 0;
            for (unsigned q = base[c] + tid; q < base[c + 1]; q += 32) {
              const unsigned long long k2 = keys[pos[q]];
              less += k2 < ZERO_KEY;
              eq += k2 == ZERO_KEY;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              less += __shfl_xor_sync(FULL, less, o);
              eq += __shfl_xor_sync(FULL, eq, o);
            }
            if (tid == 0) {
              s_zneg = (int)base[c] + less;
              s_zs = eq;
            }
          }
          __syncthreads();
        }
        if (p.dense_sem) {
          nneg = s_zneg;
          z = s_zs + zimp;
        } else {
          nneg = 0;
          z = 0;
        }
        // one thread per SLOT of the class-ordered list: the lanes of a warp sit in the same one or two buckets, so
        // their member loops have the same length and read the same addresses (broadcast)
        for (int q = tid; q < nv; q += RT) {
          const int l = (int)pos[q];
          const unsigned long long key = keys[l];
          int first, last;
          if (p.dense_sem && key == ZERO_KEY) {
            first = nneg;
            last = nneg + z;
          } else {
            int lo = 0, hi = ncode + 1;  // class of slot q: base[c] <= q < base[c + 1]
            while (lo < hi) {
              const int mid = (lo + hi) >> 1;
              if (base[mid] <= (unsigned)q) lo = mid + 1; else hi = mid;
            }
            const int c = lo - 1;
            if (c & 1) {
              first = (int)base[c];
              last = (int)base[c + 1];
            } else {
              int less = 0, eq = 0;
              const unsigned q1 = base[c + 1];
#pragma unroll 4
              for (unsigned q2 = base[c]; q2 < q1; ++q2) {
                const unsigned long long k2 = keys[pos[q2]];
                less += k2 < key;
                eq += k2 == key;
              }
              first = (int)base[c] + less;
              last = first + eq;
            }
            if (p.dense_sem && key > ZERO_KEY) {
              first += zimp;
              last += zimp;
            }
          }
          double r = p.ties == PLAIDGPU_TIES_AVERAGE ? 0.5 * (double)(first + 1 + last)
                     : p.ties == PLAIDGPU_TIES_MIN   ? (double)(first + 1)
                                                     : (double)last;
          if (p.is_signed) {
            const double v = p.xx[c0 + l];
            r = v > 0.0 ? r : (v < 0.0 ? -r : 0.0);
          }
          mymax = fmax(mymax, fabs(r));
          p.rank[c0 + l] = r;
        }
        bucket_done = true;
      } else {
        for (int l = tid; l < n; l += RT) pos[l] = (PosT)l;  // the network needs the identity positions back
      }
      __syncthreads();
    }
    if (!bucket_done) {
    bitonic_sort<PosT>(keys, pos, n, ordered);
    const int nv = n - s_nnan;  // NaN sorted last
    if (p.ties == PLAIDGPU_TIES_DENSE) {
      // RT partial sums: the (idle) bucket region behind keys + positions; global workspace variant: the dynamic buffer
      unsigned* const s_part = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(rsm) + (GLOBAL_WS ? 0 : p.bucket_off));
      // rank = number of distinct values <= the key: run starts counted per contiguous chunk of sorted slots, scanned
      const int per = (nv + RT - 1) / RT, a = min(nv, tid * per), b = min(nv, a + per);
      unsigned cntd = 0;
      for (int q = a; q < b; ++q) cntd += (q == 0 || keys[q] != keys[q - 1]);
      s_part[tid] = cntd;
      __syncthreads();
      if (tid < 32) {
        unsigned loc = 0;
        for (int i = 0; i < RT / 32; ++i) loc += s_part[tid * (RT / 32) + i];
        unsigned incl = loc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned t = __shfl_up_sync(FULL, incl, o);
          if (tid >= o) incl += t;
        }
        unsigned run = incl - loc;
        for (int i = 0; i < RT / 32; ++i) {
          const unsigned t = s_part[tid * (RT / 32) + i];
          s_part[tid * (RT / 32) + i] = run;
          run += t;
        }
      }
      __syncthreads();
      unsigned d = s_part[tid];
      for (int q = a; q < b; ++q) {
        d += (q == 0 || keys[q] != keys[q - 1]);
        double r = (double)d;
        if (p.is_signed) {
          const double v = p.xx[c0 + pos[q]];
          r = v > 0.0 ? r : (v < 0.0 ? -r : 0.0);
        }
        mymax = fmax(mymax, fabs(r));
        p.rank[c0 + pos[q]] = r;
      }
      for (int q = nv + tid; q < n; q += RT) p.rank[c0 + pos[q]] = nan("");
    } else {
    // zero group (dense semantics): stored zeros + implicit zeros
    int zs = 0, zimp = 0;
    nneg = 0;
    if (p.dense_sem) {
      nneg = lower_bound(keys, nv, ZERO_KEY);
      zs = upper_bound(keys, nv, ZERO_KEY) - nneg;
      zimp = p.P - n;
    }
    z = zs + zimp;
    for (int q = tid; q < n; q += RT) {
      const unsigned long long key = keys[q];
      double r;
      if (q >= nv) {
        r = nan("");
      } else {
        int first, last;  // 0-based inclusive-exclusive tie run in the full column order
        if (p.dense_sem && key == ZERO_KEY) {
          first = nneg;
          last = nneg + z;
        } else {
          first = (q > 0 && keys[q - 1] == key) ? lower_bound(keys, q, key) : q;
          last = (q + 1 < nv && keys[q + 1] == key) ? q + 1 + upper_bound(keys + q + 1, nv - q - 1, key) : q + 1;
          if (p.dense_sem && key > ZERO_KEY) {
            first += zimp;
            last += zimp;
          }
        }
        r = p.ties == PLAIDGPU_TIES_AVERAGE ? 0.5 * (double)(first + 1 + last)
            : p.ties == PLAIDGPU_TIES_MIN   ? (double)(first + 1)
            : p.ties == PLAIDGPU_TIES_MAX   ? (double)last
            : p.ties == PLAIDGPU_TIES_FIRST ? (double)(q + 1)            // the stable order IS the order of appearance
                                            : (double)(first + last - q);  // last: the run in reverse
        if (p.is_signed) {
          const double v = p.xx[c0 + pos[q]];
          r = v > 0.0 ? r : (v < 0.0 ? -r : 0.0);
        }
        mymax = fmax(mymax, fabs(r));
      }
      p.rank[c0 + pos[q]] = r;
    }
    }  // ties != dense
    }  // !bucket_done
    }  // !fast_done
    double rz = 0.0;
    if (p.dense_sem && z > 0 && !p.is_signed) {
      rz = p.ties == PLAIDGPU_TIES_AVERAGE ? (double)nneg + 0.5 * (double)(z + 1)
           : p.ties == PLAIDGPU_TIES_MIN   ? (double)(nneg + 1)
                                           : (double)(nneg + z);
      mymax = fmax(mymax, rz);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mymax = fmax(mymax, __shfl_xor_sync(FULL, mymax, o));
    if ((tid & 31) == 0) s_max[tid >> 5] = mymax;
    __syncthreads();
    if (tid == 0) {
      double m = 0.0;
      for (int i = 0; i < RT / 32; ++i) m = fmax(m, s_max[i]);
      if (p.colmax) p.colmax[j] = m;
      if (p.r0) p.r0[j] = rz;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_expand(const int32_t* __restrict__ xp, const int32_t* __restrict__ xi,
                                                const double* __restrict__ rank,
                                                const double* __restrict__ r0, int32_t P, int64_t N,
                                                double* __restrict__ dense) {
  for (int64_t j = blockIdx.x; j < N; j += gridDim.x) {
    double* __restrict__ d = dense + j * (int64_t)P;
    const double fill = r0 ? r0[j] : 0.0;
    for (int l = threadIdx.x; l < P; l += blockDim.x) d[l] = fill;
    __syncthreads();
    const int64_t c0 = xp[j], c1 = xp[j + 1];
    for (int64_t e = c0 + threadIdx.x; e < c1; e += blockDim.x) d[xi[e]] = rank[e];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_xform_dense(const double* __restrict__ in, double* __restrict__ out,
                                                     int64_t n, int mode, double a0, double a1) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = xform_value(mode, in[i], a0, a1);
}

__global__ void __launch_bounds__(256) k_max_col_nnz(const int32_t* __restrict__ xp, int64_t N, int32_t* res) {
  int m = 0;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < N;
       j += (int64_t)gridDim.x * blockDim.x)
    m = max(m, xp[j + 1] - xp[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(res, m);
}

// one thread per row, columns in order: deterministic; Neumaier-compensated so the fp64 result is the
// correctly rounded sum in all but pathological cases (R sums rows in long double)
__global__ void __launch_bounds__(128) k_row_moments(const double* __restrict__ x, int32_t P, int64_t N,
                                                     const double* __restrict__ mean, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= P) return;
  const double mu = mean ? mean[r] : 0.0;
  double s = 0.0, comp = 0.0;
  for (int64_t j = 0; j < N; ++j) {
    double v = x[j * (int64_t)P + r];
    if (mean) {
      v -= mu;
      v *= v;
    }
    const double t = s + v;
    comp += (fabs(s) >= fabs(v)) ? (s - t) + v : (v - t) + s;
    s = t;
  }
  out[r] = s + comp;
}

__global__ void __launch_bounds__(256) k_ztransform(const double* __restrict__ x, int32_t P, int64_t N,
                                                    const double* __restrict__ mean, const double* __restrict__ sd,
                                                    double* __restrict__ z) {
  for (int64_t j = blockIdx.x; j < N; j += gridDim.x)
    for (int r = threadIdx.x; r < P; r += blockDim.x)
      z[j * (int64_t)P + r] = (x[j * (int64_t)P + r] - mean[r]) / (1e-8 + sd[r]);
}

// out[c * R + r] = in[r * C + c] * scale : (R x C column-major ... ) generic tiled transpose of a column-major
// matrix with `rows` rows and `cols` columns into its transpose (cols x rows, column-major)
__global__ void __launch_bounds__(256) k_transpose(const double* __restrict__ in, int64_t rows, int64_t cols,
                                                   double scale, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.x * 32, c0 = (int64_t)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int64_t r = r0 + tx, c = c0 + k;
    tile[k][tx] = (r < rows && c < cols) ? in[c * rows + r] : 0.0;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int64_t c = c0 + tx, r = r0 + k;
    if (r < rows && c < cols) out[r * cols + c] = tile[tx][k] * scale;
  }
}

int sm_count() {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms;
}

template <typename PosT>
cudaError_t launch_rank_impl(RankParams p, int max_n, int64_t total, cudaStream_t st) {
  int dev = 0, smem_optin = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  const size_t per = sizeof(unsigned long long) + sizeof(PosT);
  int cap = (max_n + 3) & ~3;  // keep the pos array and the bucket region 8-byte aligned behind the keys
  if (cap < 4) cap = 4;
  size_t need = (size_t)cap * per;
  // bucket region behind keys + positions: SS splitters, 2 x (2 SS + 2) counters, RT partial sums
  const bool wide = max_n >= 8192;  // long columns: <= 2 CTAs per SM by shared memory anyway, so wide CTAs
  const int threads = wide ? RT_MAX : RT_SHORT;
  auto bucket_bytes = [threads](int ss) { return (size_t)ss * 8 + 2 * (size_t)(2 * ss + 2) * 4 + (size_t)threads * 4; };
  int ss = 0;
  if (max_n >= BUCKET_MIN_N && !getenv("PLAIDGPU_RANK_NETWORK")) {
    ss = 32;
    while (ss * (wide ? 8 : 32) <= max_n && ss < 1024) ss <<= 1;
    while (ss >= 32 && need + bucket_bytes(ss) + 2048 > (size_t)smem_optin) ss >>= 1;  // a coarser split still beats the network
    if (ss < 32) ss = 0;
  }
  p.ss_max = ss;
  p.bucket_off = need;
  need += ss ? bucket_bytes(ss) : (size_t)threads * 4;  // at least the partial sums of the dense-rank scan
  if (need < sizeof(RankFast<1024>)) need = sizeof(RankFast<1024>);
  if (wide && need < sizeof(RankFast<2048>)) need = sizeof(RankFast<2048>);
  cudaError_t e;
  if (need + 2048 <= (size_t)smem_optin) {
    p.cap = cap;
    const void* fn = wide ? (const void*)k_rank<PosT, false, true> : (const void*)k_rank<PosT, false, false>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
    if (e != cudaSuccess) return e;
    int per_sm = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, threads, need);
    if (per_sm < 1) per_sm = 1;
    int64_t grid = (int64_t)sm_count() * per_sm;
    if (grid > p.N) grid = p.N;
    void* args[] = {(void*)&p};
    e = cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(threads), args, need, st);
    if (e != cudaSuccess) return e;
    return cudaGetLastError();
  }
  // column does not fit in shared memory: sort in a global workspace (slow path)
  unsigned long long* wk = nullptr;
  PosT* wp = nullptr;
  e = cudaMallocAsync(&wk, (size_t)total * sizeof(unsigned long long), st);
  if (e != cudaSuccess) return e;
  e = cudaMallocAsync(&wp, (size_t)total * sizeof(PosT), st);
  if (e != cudaSuccess) return e;
  p.ws_keys = wk;
  p.ws_pos = wp;
  int64_t grid = (int64_t)sm_count() * 4;
  if (grid > p.N) grid = p.N;
  k_rank<PosT, true, false><<<(unsigned)grid, RT_SHORT, RT_SHORT * sizeof(unsigned), st>>>(p);
  e = cudaGetLastError();
  cudaFreeAsync(wk, st);
  cudaFreeAsync(wp, st);
  return e;
}

}  // namespace

cudaError_t launch_rank_csc(const int32_t* xp, const double* xx, int32_t P, int64_t N, int ties,
                            int is_signed, int dense_semantics, double* rank, double* r0,
                            double* colmax, int32_t max_col_nnz, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  RankParams p{};
  p.xp = xp; p.xx = xx; p.P = P; p.N = N; p.ties = ties; p.is_signed = is_signed;
  p.dense_sem = dense_semantics; p.rank = rank; p.r0 = r0; p.colmax = colmax;
  int32_t last = 0;
  cudaError_t e = cudaMemcpyAsync(&last, xp + N, sizeof(int32_t), cudaMemcpyDeviceToHost, st);
  if (e != cudaSuccess) return e;
  e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return e;
  if (max_col_nnz <= 65536) return launch_rank_impl<uint16_t>(p, max_col_nnz, last, st);
  return launch_rank_impl<uint32_t>(p, max_col_nnz, last, st);
}

cudaError_t launch_rank_dense(const double* x, int32_t P, int64_t N, int ties, int is_signed,
                              double* rank, double* colmax, cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  RankParams p{};
  p.xp = nullptr; p.xx = x; p.P = P; p.N = N; p.ties = ties; p.is_signed = is_signed;
  p.dense_sem = 0; p.rank = rank; p.r0 = nullptr; p.colmax = colmax;
  if (P <= 65536) return launch_rank_impl<uint16_t>(p, P, (int64_t)P * N, st);
  return launch_rank_impl<uint32_t>(p, P, (int64_t)P * N, st);
}

cudaError_t launch_expand_ranks(const int32_t* xp, const int32_t* xi, const double* rank,
                                const double* r0, int32_t P, int64_t N, double* dense,
                                cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid > N) grid = N;
  k_expand<<<(unsigned)grid, 256, 0, st>>>(xp, xi, rank, r0, P, N, dense);
  return cudaGetLastError();
}

cudaError_t launch_xform_dense(const double* in, double* out, int64_t n, int mode, double a0,
                               double a1, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int64_t grid = (n + 255) / 256;
  if (grid > (int64_t)sm_count() * 16) grid = (int64_t)sm_count() * 16;
  k_xform_dense<<<(unsigned)grid, 256, 0, st>>>(in, out, n, mode, a0, a1);
  return cudaGetLastError();
}

cudaError_t launch_row_moments(const double* x, int32_t P, int64_t N, const double* mean, double* out, cudaStream_t st) {
  if (P <= 0) return cudaSuccess;
  k_row_moments<<<(unsigned)((P + 127) / 128), 128, 0, st>>>(x, P, N, mean, out);
  return cudaGetLastError();
}

cudaError_t launch_ztransform(const double* x, int32_t P, int64_t N, const double* mean, const double* sd, double* z,
                              cudaStream_t st) {
  if (N <= 0) return cudaSuccess;
  int64_t grid = (int64_t)sm_count() * 8;
  if (grid > N) grid = N;
  k_ztransform<<<(unsigned)grid, 256, 0, st>>>(x, P, N, mean, sd, z);
  return cudaGetLastError();
}

cudaError_t launch_densify(const int32_t* xp, const int32_t* xi, const double* xx, int32_t P, int64_t N, double* dense,
                           cudaStream_t st) {
  return launch_expand_ranks(xp, xi, xx, nullptr, P, N, dense, st);  // same scatter: zeros filled, stored entries placed
}

cudaError_t launch_transpose(const double* in, int64_t rows, int64_t cols, double scale, double* out, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return cudaSuccess;
  dim3 g((unsigned)((rows + 31) / 32), (unsigned)((cols + 31) / 32));
  k_transpose<<<g, 256, 0, st>>>(in, rows, cols, scale, out);
  return cudaGetLastError();
}

cudaError_t launch_max_col_nnz(const int32_t* xp, int64_t N, int32_t* d_res, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(d_res, 0, sizeof(int32_t), st);
  if (e != cudaSuccess || N <= 0) return e;
  int64_t grid = (N + 255) / 256;
  if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
  k_max_col_nnz<<<(unsigned)grid, 256, 0, st>>>(xp, N, d_res);
  return cudaGetLastError();
}

}  // namespace plaidgpu
