// Time-domain regression losses (MSE / log-MSE / log1p-MSE / SDR / SI-SDR / SA-SDR) and their PIT
// wrapper, computed from one streaming pass of double-precision pair statistics.
// Reference: padertorch/ops/losses/regression.py:4-376 and the loop of TasNet.loss
// (padertorch/contrib/examples/source_separation/tasnet/model.py:154-176: 3 loss fns x K! permutations
// x B examples of ~12 small ATen reductions each).
//
// Every loss on this path is a function of {<e_i,t_j>, |e_i|^2, |t_j|^2, sum e_i, sum t_j}: products of
// fp32 inputs are exact in fp64, so the expansions |e - a t|^2 = Ee - 2 a D + a^2 Tt do not suffer the
// cancellation they would in fp32, and one read of (estimate, target) serves all K^2 pairs, all K!
// permutations and all loss kinds.  The gradient is a per-row affine map a e_i + b t_k + c.
#include <algorithm>
#include <math.h>
#include <stdlib.h>
#include <cooperative_groups.h>

#include "common.cuh"
#include "tma.cuh"
#include "perm.cuh"

using namespace b2s;

namespace {

constexpr int kStatsThreads = 256;
constexpr double kLn10 = 2.302585092994045684;

// A set of PIT losses evaluated from the statistics (b2s_pair_loss_set), and the same set evaluated INSIDE the
// statistics kernel by the CTA that folds an example's chunks (b2s_pair_stats_loss_set: one launch for TasNet.loss).
struct LossSet { int n; int kind[B2S_MAX_LOSS_SET]; int reduction[B2S_MAX_LOSS_SET]; };
struct FusedLossSet {
  LossSet set;            // set.n == 0: statistics only
  int flags;
  double tau;
  float* loss;            // [n][examples]
  int32_t* perm;          // [n][examples][K]
  float* mean;            // [n]
  int* done;              // ticket of finished examples (zero between calls)
  int64_t examples;
};
// Value of one pair loss and its partial derivatives w.r.t. the estimate-side statistics.
struct PairEval { double value, dEe, dD, dSe; };

__device__ inline PairEval eval_pair(int kind, int flags, double tau, double T, double Ee, double D,
                                     double Tt, double Se, double St) {
  PairEval r;
  r.dSe = 0.0;
  const bool offset = (flags & B2S_FLAG_OFFSET_INVARIANT) != 0;
  const double Se0 = Se, St0 = St;
  if (offset) {  // statistics of the mean-removed signals
    Ee -= Se * Se / T;
    D -= Se * St / T;
    Tt -= St * St / T;
  }
  switch (kind) {
    case B2S_LOSS_MSE: {
      r.value = (Ee - 2.0 * D + Tt) / T;
      r.dEe = 1.0 / T; r.dD = -2.0 / T;
    } break;
    case B2S_LOSS_LOG_MSE: {
      double m = (Ee - 2.0 * D + Tt) / T;
      if (tau >= 0.0) m += tau * Tt / T;
      r.value = log10(m);
      const double dm = 1.0 / (m * kLn10);
      r.dEe = dm / T; r.dD = -2.0 * dm / T;
    } break;
    case B2S_LOSS_LOG1P_MSE: {
      const double m = (Ee - 2.0 * D + Tt) / T;
      r.value = log10(1.0 + m);
      const double dm = 1.0 / ((1.0 + m) * kLn10);
      r.dEe = dm / T; r.dD = -2.0 * dm / T;
    } break;
    case B2S_LOSS_SDR: {
      double den = Ee - 2.0 * D + Tt;
      if (tau >= 0.0) den += tau * Tt;
      r.value = -10.0 * log10(Tt / den);
      const double dden = 10.0 / (kLn10 * den);
      r.dEe = dden; r.dD = -2.0 * dden;
    } break;
    default: {  // B2S_LOSS_SI_SDR
      const double alpha = D / Tt;
      const double sig = alpha * alpha * Tt;
      const double noise = Ee - 2.0 * alpha * D + alpha * alpha * Tt;
      const double den = tau >= 0.0 ? noise + tau * sig : noise;
      r.value = -10.0 * log10(sig / den);
      const double dnoise = 10.0 / (kLn10 * den);
      r.dEe = dnoise;
      if (flags & B2S_FLAG_GRAD_STOP) {
        r.dD = -2.0 * alpha * dnoise;
      } else {
        // sig = D^2/Tt, noise = Ee - D^2/Tt
        const double dsig = -10.0 / kLn10 * (1.0 / sig - (tau >= 0.0 ? tau / den : 0.0));
        r.dD = 2.0 * alpha * (dsig - dnoise);
      }
    } break;
  }
  if (offset) r.dSe = r.dEe * (-2.0 * Se0 / T) + r.dD * (-St0 / T);
  return r;
}


__host__ __device__ inline int stats_per_group(int K) { return K * K + 4 * K; }

int pair_chunks(int64_t groups, int64_t max_length) {
  // 2 CTAs of 256 threads per SM, each thread with 4K 16-byte loads in flight, cover the latency-bandwidth
  // product; more CTAs only shorten the per-thread loops and multiply the reduction epilogues (measured: 2 / 4 / 8
  // CTAs per SM -> 27.0 / 28.6 / 35.8 us for the TasNet forward at batch 64 x 2 x 4 s)
  static const int per_sm = [] { const char* e = getenv("B2S_PAIR_CTAS"); return e ? std::max(1, atoi(e)) : 2; }();
  const int64_t capacity = (int64_t)kNumSMs * per_sm;
  int64_t c = capacity / std::max<int64_t>(1, groups);
  const int64_t most = std::max<int64_t>(1, max_length / (kStatsThreads * 8));
  return (int)std::max<int64_t>(1, std::min<int64_t>(c, most));
}

// The K! assignments of one K x K cost matrix in itertools order; the first minimum wins (source_separation.py:112-119).
template <int K>
__device__ __forceinline__ double best_assignment(const double* cost, int (&bp)[K]) {
  int p[K];
#pragma unroll
  for (int k = 0; k < K; ++k) { p[k] = k; bp[k] = k; }
  double best = 0.0;
  int idx = 0;
  do {
    double v = 0.0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int i = 0; i < K; ++i) if (p[k] == i) v += cost[i * K + k];
    }
    if (idx == 0 || candidate_better(v, idx, best, 0)) {   // strictly better only
      best = v;
#pragma unroll
      for (int k = 0; k < K; ++k) bp[k] = p[k];
    }
    ++idx;
  } while (next_permutation(p, K));
  return best;
}

// CLUSTER: the chunks of a group are the CTAs of one thread-block cluster (cluster dims (1, nchunks), nchunks <= 8);
// their partial statistics meet in the shared memory of the cluster's first CTA, which folds them in chunk order --
// no global partials, no fence, no ticket: four dependent global round trips less in the tail of a ~10 us launch.
template <int K, bool PIPE, bool CLUSTER>
__global__ void __launch_bounds__(kStatsThreads, K <= 4 ? 2 : 1)
pair_stats_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                  const int64_t* __restrict__ meta, int nchunks, int64_t est_stride, int64_t tgt_stride,
                  double* __restrict__ partial, int* __restrict__ counters, double* __restrict__ stats,
                  const FusedLossSet f) {
  constexpr int NV = K * K + 4 * K;
  __shared__ double sm[NV * (kStatsThreads / 32)];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // a dependent (the loss-set kernel) may be scheduled
  const int g = blockIdx.x, chunk = blockIdx.y;
  const int64_t T = meta[g * B2S_PAIR_META + 0];
  const float* e_ = est + meta[g * B2S_PAIR_META + 1];
  const float* t_ = tgt + meta[g * B2S_PAIR_META + 2];
  const int64_t n0 = T * chunk / nchunks, n1 = T * (chunk + 1) / nchunks;
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  auto fold = [&](const float (&e)[K], const float (&t)[K]) {
    double ed[K], td[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { ed[i] = (double)e[i]; td[i] = (double)t[i]; }
#pragma unroll
    for (int i = 0; i < K; ++i) {
#pragma unroll
      for (int j = 0; j < K; ++j) acc[i * K + j] = fma(ed[i], td[j], acc[i * K + j]);
      acc[K * K + i] = fma(ed[i], ed[i], acc[K * K + i]);
      acc[K * K + K + i] = fma(td[i], td[i], acc[K * K + K + i]);
      acc[K * K + 2 * K + i] += ed[i];
      acc[K * K + 3 * K + i] += td[i];
    }
  };
  // rows that start at 16-byte aligned addresses are read with 16-byte loads, two per row in flight; whole
  // float4 units are split over the chunks, the T % 4 tail samples belong to the last chunk
  const bool vec = ((reinterpret_cast<uintptr_t>(e_) | reinterpret_cast<uintptr_t>(t_)) & 15) == 0 &&
                   (est_stride & 3) == 0 && (tgt_stride & 3) == 0;
  if (vec) {
    const int64_t nv = T >> 2;
    const int64_t v0 = nv * chunk / nchunks, v1 = nv * (chunk + 1) / nchunks;
    const float4* e4 = reinterpret_cast<const float4*>(e_);
    const float4* t4 = reinterpret_cast<const float4*>(t_);
    const int64_t es4 = est_stride >> 2, ts4 = tgt_stride >> 2;
    auto fold4 = [&](const float4 (&e)[K], const float4 (&t)[K]) {
      float a[K], b[K];
#pragma unroll
      for (int i = 0; i < K; ++i) { a[i] = e[i].x; b[i] = t[i].x; }
      fold(a, b);
#pragma unroll
      for (int i = 0; i < K; ++i) { a[i] = e[i].y; b[i] = t[i].y; }
      fold(a, b);
#pragma unroll
      for (int i = 0; i < K; ++i) { a[i] = e[i].z; b[i] = t[i].z; }
      fold(a, b);
#pragma unroll
      for (int i = 0; i < K; ++i) { a[i] = e[i].w; b[i] = t[i].w; }
      fold(a, b);
    };
    int64_t v = v0 + threadIdx.x;
    if constexpr (PIPE && K <= 2) {
      // software pipeline over units of two 16-byte loads per row: the NEXT unit's 4K loads are issued before the
      // current unit is folded, so every thread keeps 4K loads in flight through the fp64 arithmetic as well
      float4 ea[K], ta[K], eb[K], tb[K];
      bool have = v + kStatsThreads < v1;
      if (have) {
#pragma unroll
        for (int i = 0; i < K; ++i) {
          ea[i] = __ldg(e4 + i * es4 + v);
          ta[i] = __ldg(t4 + i * ts4 + v);
          eb[i] = __ldg(e4 + i * es4 + v + kStatsThreads);
          tb[i] = __ldg(t4 + i * ts4 + v + kStatsThreads);
        }
      }
      while (have) {
        const int64_t vn = v + 2 * kStatsThreads;
        const bool more = vn + kStatsThreads < v1;
        float4 na[K], nta[K], nb[K], ntb[K];
        if (more) {
#pragma unroll
          for (int i = 0; i < K; ++i) {
            na[i] = __ldg(e4 + i * es4 + vn);
            nta[i] = __ldg(t4 + i * ts4 + vn);
            nb[i] = __ldg(e4 + i * es4 + vn + kStatsThreads);
            ntb[i] = __ldg(t4 + i * ts4 + vn + kStatsThreads);
          }
        }
        fold4(ea, ta);
        fold4(eb, tb);
        if (more) {
#pragma unroll
          for (int i = 0; i < K; ++i) { ea[i] = na[i]; ta[i] = nta[i]; eb[i] = nb[i]; tb[i] = ntb[i]; }
        }
        v = vn;
        have = more;
      }
    }
    for (; K <= 3 && v + kStatsThreads < v1; v += 2 * kStatsThreads) {   // (register budget: K <= 3 only)
      float4 ea[K], ta[K], eb[K], tb[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        ea[i] = __ldg(e4 + i * es4 + v);
        ta[i] = __ldg(t4 + i * ts4 + v);
        eb[i] = __ldg(e4 + i * es4 + v + kStatsThreads);
        tb[i] = __ldg(t4 + i * ts4 + v + kStatsThreads);
      }
      fold4(ea, ta);
      fold4(eb, tb);
    }
    for (; v < v1; v += kStatsThreads) {
      float4 ea[K], ta[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        ea[i] = __ldg(e4 + i * es4 + v);
        ta[i] = __ldg(t4 + i * ts4 + v);
      }
      fold4(ea, ta);
    }
    if (chunk == nchunks - 1 && (nv << 2) + threadIdx.x < T) {
      const int64_t n = (nv << 2) + threadIdx.x;
      float e0[K], t0[K];
#pragma unroll
      for (int i = 0; i < K; ++i) {
        e0[i] = __ldg(e_ + i * est_stride + n);
        t0[i] = __ldg(t_ + i * tgt_stride + n);
      }
      fold(e0, t0);
    }
  }
  int64_t n = vec ? n1 : n0 + threadIdx.x;
  for (; n + kStatsThreads < n1; n += 2 * kStatsThreads) {  // two samples in flight per row
    float e0[K], t0[K], e1[K], t1[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      e0[i] = __ldg(e_ + i * est_stride + n);
      t0[i] = __ldg(t_ + i * tgt_stride + n);
      e1[i] = __ldg(e_ + i * est_stride + n + kStatsThreads);
      t1[i] = __ldg(t_ + i * tgt_stride + n + kStatsThreads);
    }
    fold(e0, t0);
    fold(e1, t1);
  }
  if (n < n1) {
    float e0[K], t0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      e0[i] = __ldg(e_ + i * est_stride + n);
      t0[i] = __ldg(t_ + i * tgt_stride + n);
    }
    fold(e0, t0);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kStatsThreads / 32;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const double s = warp_sum(acc[i]);
    if (lane == 0) sm[i * nwarps + warp] = s;
  }
  __syncthreads();
  __shared__ int s_last;
  if constexpr (CLUSTER) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double mine_sm[NV];
    for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s += sm[i * nwarps + w];
      mine_sm[i] = s;
    }
    cluster.sync();   // every chunk's sums are in place (and the reads of sm above are done)
    if (chunk == 0) {
      for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
        double s = 0.0;
        for (int c = 0; c < nchunks; ++c) s += *cluster.map_shared_rank(&mine_sm[i], c);   // chunk order
        stats[(int64_t)g * NV + i] = s;
        sm[i] = s;
      }
    }
    cluster.sync();   // the other CTAs' shared memory stays alive until the first CTA has read it
    if (chunk != 0) return;
  } else {
  double* mine = partial + ((int64_t)g * nchunks + chunk) * NV;
  for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sm[i * nwarps + w];
    mine[i] = s;
  }
  if (threadIdx.x < NV) __threadfence();   // the writers of the partial sums
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counters + g, 1) == nchunks - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
    const double s = ordered_sum(partial + (int64_t)g * nchunks * NV + i, nchunks, NV);
    stats[(int64_t)g * NV + i] = s;
    sm[i] = s;
  }
  if (threadIdx.x == 0) counters[g] = 0;
  }
  if constexpr (K <= 4) {
    // b2s_pair_stats_loss_set (inner == 1): the CTA that folded the example evaluates the loss set from the
    // statistics it holds; the CTA that finishes the LAST example folds the batch means -- one launch in all
    if (f.set.n == 0) return;
    constexpr int KK = K * K;
    __shared__ double cost_sm[B2S_MAX_LOSS_SET * KK];
    __syncthreads();
    if ((int)threadIdx.x < f.set.n * KK) {
      const int which = threadIdx.x / KK, ij = threadIdx.x - which * KK;
      const int ci = ij / K, cj = ij - ci * K;
      cost_sm[threadIdx.x] = eval_pair(f.set.kind[which], f.flags, f.tau, (double)T, sm[KK + ci], sm[ci * K + cj],
                                       sm[KK + K + cj], sm[KK + 2 * K + ci], sm[KK + 3 * K + cj]).value;
    }
    __syncthreads();
    if ((int)threadIdx.x < f.set.n) {
      const int which = threadIdx.x;
      int bp[K];
      double best = best_assignment<K>(cost_sm + which * KK, bp);
      if (f.set.reduction[which] == B2S_REDUCE_MEAN) best /= (double)K;
      f.loss[which * f.examples + g] = (float)best;
#pragma unroll
      for (int k = 0; k < K; ++k) f.perm[(which * f.examples + g) * K + k] = bp[k];
      __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(f.done, 1) == (int)f.examples - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (warp < f.set.n) {   // batch means, fixed order: lane-strided partial sums folded by the warp tree
      const volatile float* row = f.loss + warp * f.examples;
      double local = 0.0;
      for (int64_t ex = lane; ex < f.examples; ex += 32) local += (double)row[ex];
      local = warp_sum(local);
      if (lane == 0) f.mean[warp] = (float)(local / (double)f.examples);
    }
    if (threadIdx.x == 0) *f.done = 0;
  }
}

// ------------------------------------------------------------------------------------------- segment-staged statistics
// The (group, segment of S samples) units of the whole call form one list; CTA i of a persistent grid owns the
// contiguous range [i total / P, (i + 1) total / P) -- equal work per CTA whatever the number of groups (the
// (group, chunk) grid above puts 64 x 4 = 256 CTAs on 148 SMs and leaves 27 % of the SM-time idle).  A unit's 2K
// row segments arrive by TMA bulk copies (enclosing 16-byte aligned range, data at the source's misalignment:
// any row alignment is served at full rate), double buffered; threads read them with conflict-free LDS.32.
// At a group boundary the CTA flushes its partial statistics; the last contributor of a group (ticket) folds
// them in slot order.
// resident CTAs per SM: the K * K + 4K double accumulators set the register budget
__host__ __device__ constexpr int seg_ctas(int K) { return K <= 2 ? 3 : (K <= 4 ? 2 : 1); }
__host__ __device__ constexpr int seg_samples(int K) { return K <= 2 ? 2048 : (K <= 4 ? 1024 : 512); }
__host__ __device__ constexpr int seg_row_floats(int K) { return seg_samples(K) + 8; }

struct SegGrid { int64_t units_per_group, total; int ctas, slots; };
SegGrid seg_grid(int64_t groups, int64_t max_length, int K) {
  SegGrid g;
  g.units_per_group = std::max<int64_t>(1, ceil_div(max_length, (int64_t)seg_samples(K)));
  g.total = groups * g.units_per_group;
  g.ctas = (int)std::max<int64_t>(1, std::min<int64_t>(g.total, (int64_t)kNumSMs * seg_ctas(K)));
  g.slots = (int)(ceil_div(g.units_per_group * g.ctas, g.total) + 2);
  return g;
}

template <int K>
__global__ void __launch_bounds__(kStatsThreads, seg_ctas(K))
pair_stats_seg_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                      const int64_t* __restrict__ meta, int64_t units_per_group, int64_t total, int slots,
                      int64_t est_stride, int64_t tgt_stride, double* __restrict__ partial,
                      int* __restrict__ counters, double* __restrict__ stats) {
  constexpr int NV = K * K + 4 * K;
  constexpr int S = seg_samples(K), RF = seg_row_floats(K);
  extern __shared__ __align__(16) float seg_sm[];   // [2][2K][RF]
  __shared__ __align__(8) uint64_t full[2];
  __shared__ double red[NV * (kStatsThreads / 32)];
  __shared__ int s_last;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int64_t P = gridDim.x, me = blockIdx.x;
  const int64_t q0 = me * total / P, q1 = (me + 1) * total / P;
  if (threadIdx.x == 0) {
    tma::mbar_init(&full[0], 1);
    tma::mbar_init(&full[1], 1);
    tma::fence_mbar_init();
  }
  __syncthreads();
  // row r < K: estimate row r, else target row r - K; the unit's segment of it
  auto issue = [&](int64_t q, int s) {   // thread 0
    const int64_t g = q / units_per_group, u = q - g * units_per_group;
    const int64_t T = meta[g * B2S_PAIR_META + 0];
    const int64_t n0 = u * S;
    const int len = (int)max((int64_t)0, min((int64_t)S, T - n0));
    const float* e_ = est + meta[g * B2S_PAIR_META + 1] + n0;
    const float* t_ = tgt + meta[g * B2S_PAIR_META + 2] + n0;
    unsigned total_bytes = 0;
    if (len > 0) {
#pragma unroll
      for (int r = 0; r < 2 * K; ++r) {
        const float* src = r < K ? e_ + r * est_stride : t_ + (r - K) * tgt_stride;
        total_bytes += (unsigned)(((reinterpret_cast<uintptr_t>(src) & 15) + (size_t)len * 4 + 15) & ~(size_t)15);
      }
    }
    tma::fence_proxy_async();   // the buffer was last read through the generic proxy
    tma::mbar_expect_tx(&full[s], total_bytes);
    if (len > 0) {
#pragma unroll
      for (int r = 0; r < 2 * K; ++r) {
        const float* src = r < K ? e_ + r * est_stride : t_ + (r - K) * tgt_stride;
        const uintptr_t a = reinterpret_cast<uintptr_t>(src);
        tma::bulk_g2s(seg_sm + (s * 2 * K + r) * RF, reinterpret_cast<const void*>(a & ~(uintptr_t)15),
                      (unsigned)(((a & 15) + (size_t)len * 4 + 15) & ~(size_t)15), &full[s]);
      }
    }
  };
  if (threadIdx.x == 0) {
    if (q0 < q1) issue(q0, 0);
    if (q0 + 1 < q1) issue(q0 + 1, 1);
  }
  double acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int nwarps = kStatsThreads / 32;

  // flush the CTA's partial statistics of group g
  auto flush = [&](int64_t g) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const double v = warp_sum(acc[i]);
      if (lane == 0) red[i * nwarps + warp] = v;
      acc[i] = 0.0;
    }
    __syncthreads();
    const int64_t first_owner = ((g * units_per_group + 1) * P + total - 1) / total - 1;
    const int64_t last_owner = (((g + 1) * units_per_group) * P + total - 1) / total - 1;
    const int slot = (int)(me - first_owner), nparts = (int)(last_owner - first_owner + 1);
    double* mine = partial + (g * slots + slot) * NV;
    for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
      double v = 0.0;
      for (int w = 0; w < nwarps; ++w) v += red[i * nwarps + w];
      mine[i] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(counters + g, 1) == nparts - 1;
    __syncthreads();
    if (s_last) {   // block-uniform
      __threadfence();
      for (int i = threadIdx.x; i < NV; i += kStatsThreads) {
        const double v = ordered_sum(partial + g * slots * NV + i, nparts, NV);
        stats[g * NV + i] = v;
      }
      if (threadIdx.x == 0) counters[g] = 0;
    }
    __syncthreads();   // red / s_last may be reused
  };

  int64_t g_cur = q0 < q1 ? q0 / units_per_group : -1;
  for (int64_t q = q0; q < q1; ++q) {
    const int s = (int)((q - q0) & 1);
    const int64_t g = q / units_per_group, u = q - g * units_per_group;
    if (g != g_cur) {   // block-uniform
      flush(g_cur);
      g_cur = g;
    }
    const int64_t T = meta[g * B2S_PAIR_META + 0];
    const int64_t n0 = u * S;
    const int len = (int)max((int64_t)0, min((int64_t)S, T - n0));
    tma::mbar_wait(&full[s], (unsigned)(((q - q0) >> 1) & 1));
    if (len > 0) {
      const float* rows[2 * K];
#pragma unroll
      for (int r = 0; r < 2 * K; ++r) {
        const float* src = (r < K ? est + meta[g * B2S_PAIR_META + 1] + r * est_stride
                                  : tgt + meta[g * B2S_PAIR_META + 2] + (r - K) * tgt_stride) + n0;
        rows[r] = seg_sm + (s * 2 * K + r) * RF + (int)((reinterpret_cast<uintptr_t>(src) & 15) >> 2) + threadIdx.x;
      }
#pragma unroll 2
      for (int j = 0; j < S / kStatsThreads; ++j) {
        if (threadIdx.x + j * kStatsThreads < len) {
          double ed[K], td[K];
#pragma unroll
          for (int i = 0; i < K; ++i) {
            ed[i] = (double)rows[i][j * kStatsThreads];
            td[i] = (double)rows[K + i][j * kStatsThreads];
          }
#pragma unroll
          for (int i = 0; i < K; ++i) {
#pragma unroll
            for (int jj = 0; jj < K; ++jj) acc[i * K + jj] = fma(ed[i], td[jj], acc[i * K + jj]);
            acc[K * K + i] = fma(ed[i], ed[i], acc[K * K + i]);
            acc[K * K + K + i] = fma(td[i], td[i], acc[K * K + K + i]);
            acc[K * K + 2 * K + i] += ed[i];
            acc[K * K + 3 * K + i] += td[i];
          }
        }
      }
    }
    __syncthreads();   // every thread is done with buffer s
    if (threadIdx.x == 0 && q + 2 < q1) issue(q + 2, s);
  }
  if (g_cur >= 0) flush(g_cur);
}


// one CTA per example (= `inner` consecutive groups)
__global__ void __launch_bounds__(128)
pair_loss_kernel(const double* __restrict__ stats, const int64_t* __restrict__ meta, int64_t inner, int K,
                 int kind, int flags, double tau, int reduction, int pit, float* __restrict__ loss,
                 int32_t* __restrict__ perm) {
  __shared__ double cost[B2S_MAX_SOURCES * B2S_MAX_SOURCES];
  const int ex = blockIdx.x;
  const int NV = stats_per_group(K);
  if (kind == B2S_LOSS_SA_SDR) {
    // aggregated over every row of the example: -10 log10(sum Tt / (sum |e-t|^2 [+ tau sum Tt]))
    if (threadIdx.x == 0) {
      double st = 0.0, sn = 0.0;
      for (int64_t c = 0; c < inner; ++c) {
        const double* s = stats + (ex * inner + c) * NV;
        for (int k = 0; k < K; ++k) {
          const double Ee = s[K * K + k], Tt = s[K * K + K + k], D = s[k * K + k];
          st += Tt; sn += Ee - 2.0 * D + Tt;
        }
      }
      if (tau >= 0.0) sn += tau * st;
      loss[ex] = (float)(-10.0 * log10(st / sn));
    }
    return;
  }
  if (!pit) {
    for (int64_t idx = threadIdx.x; idx < inner * K; idx += blockDim.x) {
      const int64_t c = idx / K; const int k = (int)(idx - c * K);
      const int64_t g = ex * inner + c;
      const double* s = stats + g * NV;
      const double T = (double)meta[g * B2S_PAIR_META];
      const PairEval r = eval_pair(kind, flags, tau, T, s[K * K + k], s[k * K + k], s[K * K + K + k],
                                   s[K * K + 2 * K + k], s[K * K + 3 * K + k]);
      loss[g * K + k] = (float)r.value;
    }
    return;
  }
  for (int ij = threadIdx.x; ij < K * K; ij += blockDim.x) {
    const int i = ij / K, j = ij - i * K;
    double v = 0.0;
    for (int64_t c = 0; c < inner; ++c) {  // fixed order
      const int64_t g = ex * inner + c;
      const double* s = stats + g * NV;
      const double T = (double)meta[g * B2S_PAIR_META];
      v += eval_pair(kind, flags, tau, T, s[K * K + i], s[i * K + j], s[K * K + K + j],
                     s[K * K + 2 * K + i], s[K * K + 3 * K + j]).value;
    }
    cost[ij] = v;
  }
  __syncthreads();
  double best;
  int bp[B2S_MAX_SOURCES];
  search_permutations(cost, K, best, bp);
  if (threadIdx.x == 0) {
    if (reduction == B2S_REDUCE_MEAN) best /= (double)(K * inner);
    loss[ex] = (float)best;
    for (int k = 0; k < K; ++k) perm[(int64_t)ex * K + k] = bp[k];
  }
}

// Several PIT losses of the same statistics in ONE launch, plus their batch means (TasNet.loss evaluates
// three loss functions per step and averages each over the batch: 3 x (loss kernel + mean kernel) otherwise).
// CTA = loss kind; thread = example (K <= 4: the K! <= 24 assignments are walked serially in itertools order).
template <int K>
__global__ void __launch_bounds__(256)
pair_loss_set_kernel(const double* __restrict__ stats, const int64_t* __restrict__ meta, int64_t examples,
                     int64_t inner, LossSet set, int flags, double tau, float* __restrict__ loss,
                     int32_t* __restrict__ perm, float* __restrict__ mean) {
  constexpr int NV = K * K + 4 * K;
  __shared__ double red[256];
  const int which = blockIdx.x, kind = set.kind[which], reduction = set.reduction[which];
  // launched with programmatic stream serialization: everything above overlaps the tail of the statistics kernel
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // thread = (example of the tile, pair (i, j)): the K * K double-precision loss evaluations of an example (log10
  // chains, latency bound) run side by side; the thread of pair 0 then walks the K! assignments
  constexpr int KK = K * K, TE = 256 / KK;   // examples per tile
  __shared__ double cost_sm[TE * KK];
  const int te = threadIdx.x / KK, ij = threadIdx.x - te * KK;
  const int ci = ij / K, cj = ij - ci * K;
  double local = 0.0;
  for (int64_t base = 0; base < examples; base += TE) {
    const int64_t ex = base + te;
    const bool on = te < TE && ex < examples;
    if (on) {
      double v = 0.0;
      for (int64_t c = 0; c < inner; ++c) {  // fixed order
        const int64_t g = ex * inner + c;
        const double* s = stats + g * NV;
        const double T = (double)meta[g * B2S_PAIR_META];
        v += eval_pair(kind, flags, tau, T, s[K * K + ci], s[ci * K + cj], s[K * K + K + cj],
                       s[K * K + 2 * K + ci], s[K * K + 3 * K + cj]).value;
      }
      cost_sm[te * KK + ij] = v;
    }
    __syncthreads();
    if (on && ij == 0) {
      int bp[K];
      double best = best_assignment<K>(cost_sm + te * KK, bp);
      if (reduction == B2S_REDUCE_MEAN) best /= (double)(K * inner);
      const float out = (float)best;
      loss[which * examples + ex] = out;
#pragma unroll
      for (int k = 0; k < K; ++k) perm[(which * examples + ex) * K + k] = bp[k];
      local += (double)out;
    }
    __syncthreads();   // cost_sm is rewritten by the next tile
  }
  // batch mean, fixed order: thread partials (one per example slot of the tiles) folded in thread order
  red[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s8 = 0.0;
    for (int j = 0; j < 8; ++j) s8 += red[threadIdx.x * 8 + j];
    s8 = warp_sum(s8);
    if (threadIdx.x == 0) mean[which] = (float)(s8 / (double)examples);
  }
}

// fallback for K > 4: mean of each row of loss [n][examples], one CTA per row, fixed order
__global__ void __launch_bounds__(256)
row_mean_kernel(const float* __restrict__ loss, int64_t examples, float* __restrict__ mean) {
  __shared__ double red[256];
  double local = 0.0;
  for (int64_t ex = threadIdx.x; ex < examples; ex += blockDim.x) local += (double)loss[blockIdx.x * examples + ex];
  red[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x < 32) {
    double s8 = 0.0;
    for (int j = 0; j < 8; ++j) s8 += red[threadIdx.x * 8 + j];
    s8 = warp_sum(s8);
    if (threadIdx.x == 0) mean[blockIdx.x] = (float)(s8 / (double)examples);
  }
}

template <int K>
__global__ void __launch_bounds__(kStatsThreads)
pair_backward_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                     const int64_t* __restrict__ meta, int nchunks, int64_t inner, int64_t est_stride,
                     int64_t tgt_stride, const double* __restrict__ stats, int kind, int flags, double tau,
                     int reduction, int pit, const int32_t* __restrict__ perm,
                     const float* __restrict__ grad_loss, int64_t grad_stride, double grad_scale,
                     float* __restrict__ grad_est) {
  constexpr int NV = K * K + 4 * K;
  __shared__ float ca[K], cb[K], cc[K];
  __shared__ int match[K];  // target index matched to estimate row i
  const int g = blockIdx.x, chunk = blockIdx.y;
  const int64_t ex = g / inner;
  const int64_t T = meta[g * B2S_PAIR_META + 0];
  const int64_t eoff = meta[g * B2S_PAIR_META + 1];
  const float* e_ = est + eoff;
  const float* t_ = tgt + meta[g * B2S_PAIR_META + 2];
  float* g_ = grad_est + eoff;
  if (threadIdx.x < K) {
    const int i = threadIdx.x;
    int k = i;
    if (pit && perm) {
      for (int kk = 0; kk < K; ++kk) if (perm[ex * K + kk] == i) k = kk;
    }
    const double* s = stats + (int64_t)g * NV;
    double dEe, dD, dSe = 0.0, up;
    if (kind == B2S_LOSS_SA_SDR) {
      double st = 0.0, sn = 0.0;
      for (int64_t c = 0; c < inner; ++c) {
        const double* sc = stats + (ex * inner + c) * NV;
        for (int kk = 0; kk < K; ++kk) {
          st += sc[K * K + K + kk];
          sn += sc[K * K + kk] - 2.0 * sc[kk * K + kk] + sc[K * K + K + kk];
        }
      }
      if (tau >= 0.0) sn += tau * st;
      dEe = 10.0 / (kLn10 * sn); dD = -2.0 * dEe;
      up = grad_scale * (double)grad_loss[ex * grad_stride];
    } else {
      const PairEval r = eval_pair(kind, flags, tau, (double)T, s[K * K + i], s[i * K + k], s[K * K + K + k],
                                   s[K * K + 2 * K + i], s[K * K + 3 * K + k]);
      dEe = r.dEe; dD = r.dD; dSe = r.dSe;
      if (pit) {
        up = grad_scale * (double)grad_loss[ex * grad_stride];
        if (reduction == B2S_REDUCE_MEAN) up /= (double)(K * inner);
      } else {
        up = grad_scale * (double)grad_loss[((int64_t)g * K + i) * grad_stride];
      }
    }
    ca[i] = (float)(up * 2.0 * dEe);
    cb[i] = (float)(up * dD);
    cc[i] = (float)(up * dSe);
    match[i] = k;
  }
  __syncthreads();
  const bool vec = ((reinterpret_cast<uintptr_t>(e_) | reinterpret_cast<uintptr_t>(t_) |
                     reinterpret_cast<uintptr_t>(g_)) & 15) == 0 && (est_stride & 3) == 0 && (tgt_stride & 3) == 0;
  if (vec) {   // 16-byte loads and stores; the T % 4 tail samples belong to the last chunk
    const int64_t nv = T >> 2;
    const int64_t v0 = nv * chunk / nchunks, v1 = nv * (chunk + 1) / nchunks;
    const float4* e4 = reinterpret_cast<const float4*>(e_);
    const float4* t4 = reinterpret_cast<const float4*>(t_);
    float4* g4 = reinterpret_cast<float4*>(g_);
    const int64_t es4 = est_stride >> 2, ts4 = tgt_stride >> 2;
    float a[K], bm[K], c[K];
    int mt[K];
#pragma unroll
    for (int i = 0; i < K; ++i) { a[i] = ca[i]; bm[i] = cb[i]; c[i] = cc[i]; mt[i] = match[i]; }
    auto emit = [&](int64_t v, const float4 (&t)[K], const float4 (&e)[K]) {
#pragma unroll
      for (int i = 0; i < K; ++i) {
        float4 tm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < K; ++k) if (mt[i] == k) tm = t[k];
        float4 o;
        o.x = fmaf(a[i], e[i].x, fmaf(bm[i], tm.x, c[i]));
        o.y = fmaf(a[i], e[i].y, fmaf(bm[i], tm.y, c[i]));
        o.z = fmaf(a[i], e[i].z, fmaf(bm[i], tm.z, c[i]));
        o.w = fmaf(a[i], e[i].w, fmaf(bm[i], tm.w, c[i]));
        g4[i * es4 + v] = o;
      }
    };
    int64_t v = v0 + threadIdx.x;
    for (; K <= 4 && v + kStatsThreads < v1; v += 2 * kStatsThreads) {   // 4K 16-byte loads in flight
      float4 ta[K], ea[K], tb[K], eb[K];
#pragma unroll
      for (int k = 0; k < K; ++k) { ta[k] = __ldg(t4 + k * ts4 + v); tb[k] = __ldg(t4 + k * ts4 + v + kStatsThreads); }
#pragma unroll
      for (int i = 0; i < K; ++i) { ea[i] = __ldg(e4 + i * es4 + v); eb[i] = __ldg(e4 + i * es4 + v + kStatsThreads); }
      emit(v, ta, ea);
      emit(v + kStatsThreads, tb, eb);
    }
    for (; v < v1; v += kStatsThreads) {
      float4 t[K], e[K];
#pragma unroll
      for (int k = 0; k < K; ++k) t[k] = __ldg(t4 + k * ts4 + v);
#pragma unroll
      for (int i = 0; i < K; ++i) e[i] = __ldg(e4 + i * es4 + v);
      emit(v, t, e);
    }
    if (chunk != nchunks - 1) return;
  }
  const int64_t n0 = vec ? (T & ~(int64_t)3) : T * chunk / nchunks, n1 = vec ? T : T * (chunk + 1) / nchunks;
  for (int64_t n = n0 + threadIdx.x; n < n1; n += kStatsThreads) {
    float t[K];
#pragma unroll
    for (int k = 0; k < K; ++k) t[k] = __ldg(t_ + k * tgt_stride + n);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float tm = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) if (match[i] == k) tm = t[k];
      const float e = __ldg(e_ + i * est_stride + n);
      g_[i * est_stride + n] = fmaf(ca[i], e, fmaf(cb[i], tm, cc[i]));
    }
  }
}


// ------------------------------------------------------------------------------------------- pairwise loss matrix
// compute_pairwise_losses (padertorch/ops/losses/source_separation.py:127-241): matrix[ex][i][j] =
// reduce over the example's `inner` groups of l(e_i, t_j) -- every entry is a function of the statistics the
// forward pass already holds, so the K x K matrix costs no further read of the signals.  One CTA per example.
__global__ void __launch_bounds__(64)
pair_loss_matrix_kernel(const double* __restrict__ stats, const int64_t* __restrict__ meta, int64_t inner, int K,
                        int kind, int flags, double tau, int reduction, float* __restrict__ matrix) {
  const int64_t ex = blockIdx.x;
  const int NV = stats_per_group(K);
  for (int ij = threadIdx.x; ij < K * K; ij += blockDim.x) {
    const int i = ij / K, j = ij - i * K;
    double v = 0.0;
    for (int64_t c = 0; c < inner; ++c) {  // fixed order
      const int64_t g = ex * inner + c;
      const double* s = stats + g * NV;
      const double T = (double)meta[g * B2S_PAIR_META];
      v += eval_pair(kind, flags, tau, T, s[K * K + i], s[i * K + j], s[K * K + K + j], s[K * K + 2 * K + i],
                     s[K * K + 3 * K + j]).value;
    }
    if (reduction == B2S_REDUCE_MEAN) v /= (double)inner;
    matrix[ex * K * K + ij] = (float)v;
  }
}

// Gradient of sum_{i,j} G[ex][i][j] matrix[ex][i][j] w.r.t. the estimate rows: every pair loss is a function of
// (Ee_i, D_ij, Se_i), so  grad e_i = (sum_j 2 G_ij dEe_ij) e_i + sum_j (G_ij dD_ij) t_j + sum_j G_ij dSe_ij  -- one
// streaming pass that reads the K estimate rows and the K target rows of a group once.
template <int K>
__global__ void __launch_bounds__(kStatsThreads)
pair_matrix_backward_kernel(const float* __restrict__ est, const float* __restrict__ tgt,
                            const int64_t* __restrict__ meta, int nchunks, int64_t inner, int64_t est_stride,
                            int64_t tgt_stride, const double* __restrict__ stats, int kind, int flags, double tau,
                            int reduction, const float* __restrict__ grad_matrix, float* __restrict__ grad_est) {
  constexpr int NV = K * K + 4 * K;
  __shared__ float ca[K], cb[K][K], cc[K];
  const int g = blockIdx.x, chunk = blockIdx.y;
  const int64_t ex = g / inner;
  const int64_t T = meta[g * B2S_PAIR_META + 0];
  const int64_t eoff = meta[g * B2S_PAIR_META + 1];
  const float* e_ = est + eoff;
  const float* t_ = tgt + meta[g * B2S_PAIR_META + 2];
  float* g_ = grad_est + eoff;
  if (threadIdx.x < K) {
    const int i = threadIdx.x;
    const double* s = stats + (int64_t)g * NV;
    const double w = reduction == B2S_REDUCE_MEAN ? 1.0 / (double)inner : 1.0;
    double a = 0.0, c = 0.0;
    for (int j = 0; j < K; ++j) {
      const PairEval r = eval_pair(kind, flags, tau, (double)T, s[K * K + i], s[i * K + j], s[K * K + K + j],
                                   s[K * K + 2 * K + i], s[K * K + 3 * K + j]);
      const double up = w * (double)grad_matrix[(ex * K + i) * K + j];
      a += up * 2.0 * r.dEe;
      c += up * r.dSe;
      cb[i][j] = (float)(up * r.dD);
    }
    ca[i] = (float)a;
    cc[i] = (float)c;
  }
  __syncthreads();
  const int64_t n0 = T * chunk / nchunks, n1 = T * (chunk + 1) / nchunks;
  for (int64_t n = n0 + threadIdx.x; n < n1; n += kStatsThreads) {
    float t[K];
#pragma unroll
    for (int j = 0; j < K; ++j) t[j] = __ldg(t_ + j * tgt_stride + n);
#pragma unroll
    for (int i = 0; i < K; ++i) {
      float v = fmaf(ca[i], __ldg(e_ + i * est_stride + n), cc[i]);
#pragma unroll
      for (int j = 0; j < K; ++j) v = fmaf(cb[i][j], t[j], v);
      g_[i * est_stride + n] = v;
    }
  }
}

// ------------------------------------------------------------------------------------------- assignment
// Optimal assignment of a batch of K x K cost matrices by exhaustive search (K <= 8: at most 40 320 candidates,
// one CTA per matrix), replacing the host round trip of pit_loss_from_loss_matrix
// (padertorch/ops/losses/source_separation.py:285-288: to_numpy + scipy.optimize.linear_sum_assignment).
//   orientation 0: out[k] = row (estimate) matched to column (target) k, candidates in itertools.permutations
//                  order, first minimum wins -- the convention of pit_loss (:112-122);
//   orientation 1: out[i] = column matched to row i (scipy's col_ind for row_ind = 0..K-1), first minimum in
//                  lexicographic order of col_ind;
//   greedy != 0  : orientation 1 by repeatedly taking the smallest remaining entry (ties: lowest row, then
//                  lowest column) -- the 'greedy' algorithm of :291-296 (pb_bss is not on disk: parity unpinned).
__global__ void __launch_bounds__(256)
assign_kernel(const float* __restrict__ cost, int K, int orientation, int greedy, int32_t* __restrict__ out,
              float* __restrict__ value) {
  __shared__ double c[B2S_MAX_SOURCES * B2S_MAX_SOURCES];
  const int64_t ex = blockIdx.x;
  const float* m = cost + ex * K * K;
  for (int ij = threadIdx.x; ij < K * K; ij += blockDim.x) {
    const int i = ij / K, j = ij - i * K;
    // search_permutations minimises sum_k c[p[k] * K + k]
    c[orientation == 0 ? ij : j * K + i] = (double)m[ij];
  }
  __syncthreads();
  if (greedy) {
    if (threadIdx.x == 0) {
      bool row_used[B2S_MAX_SOURCES] = {}, col_used[B2S_MAX_SOURCES] = {};
      double total = 0.0;
      for (int step = 0; step < K; ++step) {
        int bi = -1, bj = -1;
        double bv = 0.0;
        for (int i = 0; i < K; ++i) {
          if (row_used[i]) continue;
          for (int j = 0; j < K; ++j) {
            if (col_used[j]) continue;
            const double v = (double)m[i * K + j];
            if (bi < 0 || v < bv) { bi = i; bj = j; bv = v; }
          }
        }
        row_used[bi] = true; col_used[bj] = true;
        out[ex * K + bi] = bj;
        total += bv;
      }
      if (value) value[ex] = (float)total;
    }
    return;
  }
  double best;
  int bp[B2S_MAX_SOURCES];
  search_permutations(c, K, best, bp);
  if (threadIdx.x == 0) {
    for (int k = 0; k < K; ++k) out[ex * K + k] = bp[k];
    if (value) value[ex] = (float)best;
  }
}

}  // namespace

extern "C" {

int64_t b2s_pair_workspace_bytes(int64_t groups, int64_t max_length, int sources) {
  if (groups <= 0 || sources <= 0) return 16;
  const int chunks = std::max(pair_chunks(groups, max_length), seg_grid(groups, max_length, sources).slots);
  return kTicketBytes + (int64_t)sizeof(double) * groups * chunks * stats_per_group(sources) + 16;
}

}  // extern "C"

namespace {
// The statistics pass; `fused` (may be NULL) asks the (group, chunk) kernel to evaluate a loss set in the same launch.
// Returns through `*fused_done` whether it did (the segment-staged kernel and K > 4 do not).
int pair_stats_launch(const float* estimate, const float* target, const int64_t* meta,
                      int64_t groups, int64_t max_length, int sources,
                      int64_t estimate_source_stride, int64_t target_source_stride, double* stats,
                      void* workspace, b2s_stream stream, const FusedLossSet* fused, bool* fused_done) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES,
              "sources=%d outside the supported range 1..%d", sources, B2S_MAX_SOURCES);
  B2S_REQUIRE(groups >= 0 && groups <= kMaxTickets && max_length >= 0, "bad extents (groups=%lld)",
              (long long)groups);
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(estimate && target && meta && stats && workspace, "NULL device pointer");
  const int chunks = pair_chunks(groups, max_length);
  double* partial = ws_partials(workspace);
  int* counters = ws_counters(workspace);
  cudaStream_t st = (cudaStream_t)stream;
  // Rows that start at 16-byte aligned addresses go to the (group, chunk) kernel with 16-byte loads; anything
  // else to the segment-staged kernel, whose TMA copies serve any alignment at the same rate (equal timings on
  // aligned rows: 26.6 vs 28.1 us for the TasNet forward).  Offsets inside `meta` are multiples of the strides
  // for every wrapper in this package.  B2S_PAIR_SEG=1 / 0 forces the choice.
  static const int force_seg = [] { const char* e = getenv("B2S_PAIR_SEG"); return e ? atoi(e) : -1; }();
  const bool aligned = ((reinterpret_cast<uintptr_t>(estimate) | reinterpret_cast<uintptr_t>(target)) & 15) == 0 &&
                       estimate_source_stride % 4 == 0 && target_source_stride % 4 == 0;
  if ((force_seg == 1 || (force_seg < 0 && !aligned)) && max_length >= 1) {
    const SegGrid sg = seg_grid(groups, max_length, sources);
    const size_t smem = sizeof(float) * 2 * 2 * sources * seg_row_floats(sources);
#define CALL_SEG(K) do {                                                                                      \
      static bool configured[64] = {};                                                                         \
      int dev = 0;                                                                                             \
      B2S_CUDA(cudaGetDevice(&dev));                                                                           \
      if (!configured[dev & 63]) {                                                                             \
        B2S_CUDA(cudaFuncSetAttribute(pair_stats_seg_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                      (int)smem));                                                             \
        configured[dev & 63] = true;                                                                           \
      }                                                                                                        \
      pair_stats_seg_kernel<K><<<sg.ctas, kStatsThreads, smem, st>>>(estimate, target, meta, sg.units_per_group, \
          sg.total, sg.slots, estimate_source_stride, target_source_stride, partial, counters, stats);          \
    } while (0)
    switch (sources) {
      case 1: CALL_SEG(1); break;
      case 2: CALL_SEG(2); break;
      case 3: CALL_SEG(3); break;
      case 4: CALL_SEG(4); break;
      case 5: CALL_SEG(5); break;
      case 6: CALL_SEG(6); break;
      case 7: CALL_SEG(7); break;
      default: CALL_SEG(8); break;
    }
#undef CALL_SEG
    B2S_LAUNCH_CHECK("pair_stats_seg_kernel");
    return B2S_OK;
  }
  const dim3 grid((unsigned)groups, chunks);
  // B2S_PAIR_PIPE=0: the loop without the software pipeline (A/B measurements)
  static const bool pipe = [] { const char* e = getenv("B2S_PAIR_PIPE"); return e ? atoi(e) != 0 : true; }();
  FusedLossSet f = {};
  if (fused && sources <= 4) {
    f = *fused;
    f.done = counters + (kMaxTickets - 1);
    *fused_done = true;
  }
  // the chunks of a group as one thread-block cluster (B2S_PAIR_CLUSTER=0: global partials + ticket)
  static const bool use_cluster = [] { const char* e = getenv("B2S_PAIR_CLUSTER"); return e ? atoi(e) != 0 : true; }();
  if (use_cluster && pipe && chunks >= 2 && sources <= 4) {
    const int cchunks = std::min(chunks, 8);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)groups, cchunks);
    cfg.blockDim = dim3(kStatsThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 1;
    attr[0].val.clusterDim.y = cchunks;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define CALLC(K) B2S_CUDA(cudaLaunchKernelEx(&cfg, pair_stats_kernel<K, true, true>, estimate, target, meta, cchunks, \
        estimate_source_stride, target_source_stride, partial, counters, stats, f))
    switch (sources) {
      case 1: CALLC(1); break;
      case 2: CALLC(2); break;
      case 3: CALLC(3); break;
      default: CALLC(4); break;
    }
#undef CALLC
    B2S_LAUNCH_CHECK("pair_stats_kernel (cluster)");
    return B2S_OK;
  }
#define CALL(K) do {                                                                                           \
    if (pipe) pair_stats_kernel<K, true, false><<<grid, kStatsThreads, 0, st>>>(estimate, target, meta, chunks, \
        estimate_source_stride, target_source_stride, partial, counters, stats, f);                             \
    else pair_stats_kernel<K, false, false><<<grid, kStatsThreads, 0, st>>>(estimate, target, meta, chunks,     \
        estimate_source_stride, target_source_stride, partial, counters, stats, f);                             \
  } while (0)
  switch (sources) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 3: CALL(3); break;
    case 4: CALL(4); break;
    case 5: CALL(5); break;
    case 6: CALL(6); break;
    case 7: CALL(7); break;
    default: CALL(8); break;
  }
#undef CALL
  B2S_LAUNCH_CHECK("pair_stats_kernel");
  return B2S_OK;
}
}  // namespace

extern "C" {

int b2s_pair_stats_forward(const float* estimate, const float* target, const int64_t* meta,
                           int64_t groups, int64_t max_length, int sources,
                           int64_t estimate_source_stride, int64_t target_source_stride, double* stats,
                           void* workspace, b2s_stream stream) {
  return pair_stats_launch(estimate, target, meta, groups, max_length, sources, estimate_source_stride,
                           target_source_stride, stats, workspace, stream, nullptr, nullptr);
}

int b2s_pair_loss(const double* stats, const int64_t* meta, int64_t groups, int64_t inner, int sources,
                  int kind, int flags, double tau, int reduction, int pit, float* loss, int32_t* perm,
                  b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d unsupported", sources);
  B2S_REQUIRE(kind >= B2S_LOSS_MSE && kind <= B2S_LOSS_SA_SDR, "unknown loss kind %d", kind);
  B2S_REQUIRE(inner >= 1 && groups % inner == 0, "groups must be a multiple of inner");
  B2S_REQUIRE(!pit || (reduction == B2S_REDUCE_SUM || reduction == B2S_REDUCE_MEAN),
              "PIT needs reduction sum or mean");
  B2S_REQUIRE(!(pit && kind == B2S_LOSS_SA_SDR), "SA-SDR is not additive over sources: no PIT variant");
  B2S_REQUIRE(!(flags & B2S_FLAG_OFFSET_INVARIANT) || kind == B2S_LOSS_SI_SDR,
              "offset_invariant exists for si_sdr only");
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(stats && meta && loss && (!pit || perm), "NULL device pointer");
  pair_loss_kernel<<<(unsigned)(groups / inner), 128, 0, (cudaStream_t)stream>>>(
      stats, meta, inner, sources, kind, flags, tau, reduction, pit, loss, perm);
  B2S_LAUNCH_CHECK("pair_loss_kernel");
  return B2S_OK;
}

int b2s_pair_loss_set(const double* stats, const int64_t* meta, int64_t groups, int64_t inner, int sources,
                      int count, const int* kinds, const int* reductions, int flags, double tau, float* loss,
                      int32_t* perm, float* mean, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d unsupported", sources);
  B2S_REQUIRE(count >= 1 && count <= B2S_MAX_LOSS_SET && kinds && reductions, "1..%d loss kinds per set (got %d)",
              B2S_MAX_LOSS_SET, count);
  B2S_REQUIRE(inner >= 1 && groups % inner == 0, "groups must be a multiple of inner");
  LossSet set;
  set.n = count;
  for (int i = 0; i < count; ++i) {
    B2S_REQUIRE(kinds[i] >= B2S_LOSS_MSE && kinds[i] < B2S_LOSS_SA_SDR, "loss kind %d has no PIT variant", kinds[i]);
    B2S_REQUIRE(reductions[i] == B2S_REDUCE_SUM || reductions[i] == B2S_REDUCE_MEAN, "PIT needs reduction sum or mean");
    B2S_REQUIRE(!(flags & B2S_FLAG_OFFSET_INVARIANT) || kinds[i] == B2S_LOSS_SI_SDR,
                "offset_invariant exists for si_sdr only");
    set.kind[i] = kinds[i];
    set.reduction[i] = reductions[i];
  }
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(stats && meta && loss && perm && mean, "NULL device pointer");
  const int64_t examples = groups / inner;
  cudaStream_t st = (cudaStream_t)stream;
  if (sources <= 4) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)count);
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
#define CALL(K) B2S_CUDA(cudaLaunchKernelEx(&cfg, pair_loss_set_kernel<K>, stats, meta, examples, inner, set, flags, tau, loss, perm, mean))
    switch (sources) {
      case 1: CALL(1); break;
      case 2: CALL(2); break;
      case 3: CALL(3); break;
      default: CALL(4); break;
    }
#undef CALL
    B2S_LAUNCH_CHECK("pair_loss_set_kernel");
  } else {
    for (int i = 0; i < count; ++i) {
      pair_loss_kernel<<<(unsigned)examples, 128, 0, st>>>(stats, meta, inner, sources, set.kind[i], flags, tau,
                                                            set.reduction[i], 1, loss + i * examples,
                                                            perm + i * examples * sources);
      B2S_LAUNCH_CHECK("pair_loss_kernel");
    }
    row_mean_kernel<<<count, 256, 0, st>>>(loss, examples, mean);
    B2S_LAUNCH_CHECK("row_mean_kernel");
  }
  return B2S_OK;
}

int b2s_pair_stats_loss_set(const float* estimate, const float* target, const int64_t* meta, int64_t groups,
                            int64_t max_length, int sources, int64_t estimate_source_stride,
                            int64_t target_source_stride, int count, const int* kinds, const int* reductions,
                            int flags, double tau, double* stats, float* loss, int32_t* perm, float* mean,
                            void* workspace, b2s_stream stream) {
  B2S_REQUIRE(count >= 1 && count <= B2S_MAX_LOSS_SET && kinds && reductions, "1..%d loss kinds per set (got %d)",
              B2S_MAX_LOSS_SET, count);
  B2S_REQUIRE(groups < kMaxTickets, "too many groups (%lld)", (long long)groups);
  FusedLossSet f = {};
  f.set.n = count;
  for (int i = 0; i < count; ++i) {
    B2S_REQUIRE(kinds[i] >= B2S_LOSS_MSE && kinds[i] < B2S_LOSS_SA_SDR, "loss kind %d has no PIT variant", kinds[i]);
    B2S_REQUIRE(reductions[i] == B2S_REDUCE_SUM || reductions[i] == B2S_REDUCE_MEAN, "PIT needs reduction sum or mean");
    B2S_REQUIRE(!(flags & B2S_FLAG_OFFSET_INVARIANT) || kinds[i] == B2S_LOSS_SI_SDR,
                "offset_invariant exists for si_sdr only");
    f.set.kind[i] = kinds[i];
    f.set.reduction[i] = reductions[i];
  }
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(loss && perm && mean, "NULL device pointer");
  f.flags = flags;
  f.tau = tau;
  f.loss = loss;
  f.perm = perm;
  f.mean = mean;
  f.examples = groups;
  bool fused_done = false;
  const int rc = pair_stats_launch(estimate, target, meta, groups, max_length, sources, estimate_source_stride,
                                   target_source_stride, stats, workspace, stream, &f, &fused_done);
  if (rc != B2S_OK || fused_done) return rc;
  return b2s_pair_loss_set(stats, meta, groups, 1, sources, count, kinds, reductions, flags, tau, loss, perm, mean,
                           stream);
}

int b2s_pair_backward(const float* estimate, const float* target, const int64_t* meta, int64_t groups,
                      int64_t inner, int64_t max_length, int sources, int64_t estimate_source_stride,
                      int64_t target_source_stride, const double* stats, int kind, int flags, double tau,
                      int reduction, int pit, const int32_t* perm, const float* grad_loss,
                      int64_t grad_loss_stride, double grad_scale, float* grad_estimate, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d unsupported", sources);
  B2S_REQUIRE(kind >= B2S_LOSS_MSE && kind <= B2S_LOSS_SA_SDR, "unknown loss kind %d", kind);
  B2S_REQUIRE(inner >= 1 && groups % inner == 0, "groups must be a multiple of inner");
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(estimate && target && meta && stats && grad_loss && grad_estimate, "NULL device pointer");
  // one wave of 4 resident CTAs per SM (the register limit of the kernel): 2 / 3 / 4 / 6 / 8 CTAs per SM worth of
  // chunks -> 31.4 / 28.6 / 25.0 / 29.1 / 27.3 us at batch 64 x 2 x 4 s; at least 2 float4 units per thread and chunk
  static const int per_sm = [] { const char* e = getenv("B2S_PAIR_BWD_CTAS"); return e ? std::max(1, atoi(e)) : 4; }();
  const int64_t want = std::max<int64_t>(1, (int64_t)kNumSMs * per_sm / groups);
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(max_length / (kStatsThreads * 8) + 1, std::min<int64_t>(want, 4096)));
  const dim3 grid((unsigned)groups, chunks);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K) pair_backward_kernel<K><<<grid, kStatsThreads, 0, st>>>(estimate, target, meta, chunks, \
      inner, estimate_source_stride, target_source_stride, stats, kind, flags, tau, reduction, pit, perm, \
      grad_loss, grad_loss_stride, grad_scale, grad_estimate)
  switch (sources) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 3: CALL(3); break;
    case 4: CALL(4); break;
    case 5: CALL(5); break;
    case 6: CALL(6); break;
    case 7: CALL(7); break;
    default: CALL(8); break;
  }
#undef CALL
  B2S_LAUNCH_CHECK("pair_backward_kernel");
  return B2S_OK;
}

int b2s_pair_loss_matrix(const double* stats, const int64_t* meta, int64_t groups, int64_t inner, int sources,
                         int kind, int flags, double tau, int reduction, float* matrix, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d unsupported", sources);
  B2S_REQUIRE(kind >= B2S_LOSS_MSE && kind < B2S_LOSS_SA_SDR, "loss kind %d has no pairwise form", kind);
  B2S_REQUIRE(inner >= 1 && groups % inner == 0, "groups must be a multiple of inner");
  B2S_REQUIRE(reduction == B2S_REDUCE_SUM || reduction == B2S_REDUCE_MEAN, "reduction must be sum or mean");
  B2S_REQUIRE(!(flags & B2S_FLAG_OFFSET_INVARIANT) || kind == B2S_LOSS_SI_SDR,
              "offset_invariant exists for si_sdr only");
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(stats && meta && matrix, "NULL device pointer");
  pair_loss_matrix_kernel<<<(unsigned)(groups / inner), 64, 0, (cudaStream_t)stream>>>(
      stats, meta, inner, sources, kind, flags, tau, reduction, matrix);
  B2S_LAUNCH_CHECK("pair_loss_matrix_kernel");
  return B2S_OK;
}

int b2s_pair_matrix_backward(const float* estimate, const float* target, const int64_t* meta, int64_t groups,
                             int64_t inner, int64_t max_length, int sources, int64_t estimate_source_stride,
                             int64_t target_source_stride, const double* stats, int kind, int flags, double tau,
                             int reduction, const float* grad_matrix, float* grad_estimate, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES, "sources=%d unsupported", sources);
  B2S_REQUIRE(kind >= B2S_LOSS_MSE && kind < B2S_LOSS_SA_SDR, "loss kind %d has no pairwise form", kind);
  B2S_REQUIRE(inner >= 1 && groups % inner == 0, "groups must be a multiple of inner");
  B2S_REQUIRE(reduction == B2S_REDUCE_SUM || reduction == B2S_REDUCE_MEAN, "reduction must be sum or mean");
  if (groups == 0) return B2S_OK;
  B2S_REQUIRE(estimate && target && meta && stats && grad_matrix && grad_estimate, "NULL device pointer");
  const int64_t want = std::max<int64_t>(1, (int64_t)kNumSMs * 4 / groups);
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>(max_length / (kStatsThreads * 4) + 1, std::min<int64_t>(want, 4096)));
  const dim3 grid((unsigned)groups, chunks);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(K) pair_matrix_backward_kernel<K><<<grid, kStatsThreads, 0, st>>>(estimate, target, meta, chunks, \
      inner, estimate_source_stride, target_source_stride, stats, kind, flags, tau, reduction, grad_matrix, \
      grad_estimate)
  switch (sources) {
    case 1: CALL(1); break;
    case 2: CALL(2); break;
    case 3: CALL(3); break;
    case 4: CALL(4); break;
    case 5: CALL(5); break;
    case 6: CALL(6); break;
    case 7: CALL(7); break;
    default: CALL(8); break;
  }
#undef CALL
  B2S_LAUNCH_CHECK("pair_matrix_backward_kernel");
  return B2S_OK;
}

int b2s_assign(const float* cost, int64_t batch, int sources, int orientation, int greedy, int32_t* assignment,
               float* value, b2s_stream stream) {
  B2S_REQUIRE(sources >= 1 && sources <= B2S_MAX_SOURCES,
              "device assignment supports 1..%d sources (got %d)", B2S_MAX_SOURCES, sources);
  B2S_REQUIRE(orientation == 0 || orientation == 1, "orientation must be 0 (pit_loss) or 1 (col_ind)");
  B2S_REQUIRE(!greedy || orientation == 1, "the greedy assignment returns col_ind (orientation 1)");
  B2S_REQUIRE(batch >= 0, "negative batch");
  if (batch == 0) return B2S_OK;
  B2S_REQUIRE(cost && assignment, "NULL device pointer");
  assign_kernel<<<(unsigned)batch, 256, 0, (cudaStream_t)stream>>>(cost, sources, orientation, greedy, assignment, value);
  B2S_LAUNCH_CHECK("assign_kernel");
  return B2S_OK;
}

}  // extern "C"
