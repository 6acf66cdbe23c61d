// Dense projections of the mask networks on the 5th-generation tensor cores (tcgen05 / UMMA, sm_100a):
//     C[M, N] = act(A[M, K] . W[N, K]^T + bias[N])                      (torch.nn.functional.linear)
// for the packed-sequence GEMMs of the PIT / deep-clustering mask estimators
// (padertorch/contrib/examples/source_separation/pit/model.py:68-72, 96-102: Linear(1200, 1200) + ReLU,
//  Linear(1200, F K) + sigmoid, and the LSTM input projections [sum T_b, F] x [F, 4 * 600]) and the 1 x 1
// convolutions of the ConvNet separator (padertorch/modules/convnet.py:120-167).
//
// Arithmetic: the reference computes these in fp32 (cuBLAS SGEMM, TF32 off by default in PyTorch).  A single
// TF32 product keeps 10 mantissa bits per operand (relative error ~5e-4 per product, far outside the 1e-4
// budget of the path); the kernel therefore evaluates the 3-term split
//     a b ~= a_hi b_hi + a_hi b_lo + a_lo b_hi,   x_hi = tf32(x) (truncation, done by the tensor core itself when
//     it reads an fp32 word),   x_lo = x - x_hi (exact in fp32, <= 13 significant bits),
// three tcgen05.mma per k-step into the same fp32 TMEM accumulator: error ~2^-21 |a||b|, i.e. fp32-faithful.
// `products` = 1 runs the plain TF32 GEMM (documented tolerance 2e-3) for callers that accept it.
// The lo parts are operands of their own ([M, K] / [N, K] fp32 arrays): W_lo is computed once per weight update,
// A_lo by the producer of A (b2s_tf32_split, or the epilogue of the previous projection: `c_lo` output).
//
// Structure (one CTA per SM, persistent over 128 x 128 output tiles, 256 threads):
//   warp 0   TMA producer: 2-D tensor maps (cuTensorMapEncodeTiled, SWIZZLE_128B, box 32 x 128 fp32) bring the
//            A / A_lo / W / W_lo tiles of a k-block (K = 32 fp32 = one 128-byte swizzle atom) into a stage of the
//            shared-memory ring, completing on the stage's `full` mbarrier;
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 128, K 8) from
//            shared-memory descriptors, tcgen05.commit releases the stage (`empty`) and, after the last k-block,
//            publishes the accumulator (`tmem_full`);
//   warp 2   allocates / frees the tensor memory (2 accumulators x 128 columns: the epilogue of tile i overlaps
//            the MMAs of tile i + 1);
//   warps 4-7 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> registers -> bias + activation ->
//            transposition through shared memory -> 128-byte row-segment stores (optionally also the lo part for a following projection).
#include <cuda.h>
#include <cuda_runtime.h>
#include <mutex>
#include <stdlib.h>

#include "common.cuh"

using namespace b2s;

namespace {

constexpr int kBlockM = 128, kBlockN = 128, kBlockK = 32;     // fp32 elements; kBlockK * 4 = 128 bytes = swizzle atom
constexpr int kUmmaK = 8;                                     // tf32: 32 bytes per instruction
constexpr int kTileBytes = kBlockM * kBlockK * 4;             // 16 KB (A and W tiles have the same shape)
constexpr int kThreads = 256;
constexpr int kAccumColumns = kBlockN;                        // fp32 accumulator: one column per n
constexpr int kTmemColumns = 2 * kAccumColumns;               // double buffered

__host__ __device__ constexpr int stages_for(int products) { return products == 3 ? 3 : 6; }
__host__ __device__ constexpr int tiles_per_stage(int products) { return products == 3 ? 4 : 2; }

// ---- PTX wrappers --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile("{\n .reg .pred p;\n GEMM_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
               " @p bra GEMM_DONE;\n bra GEMM_WAIT;\n GEMM_DONE:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void prefetch_map(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
// shared-memory matrix descriptor of a K-major tile with 128-byte rows, SWIZZLE_128B: rows of an 8-row group are
// 128 bytes apart, groups 1024 bytes (stride byte offset); descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)1 << 16;                               // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset
  d |= (uint64_t)1 << 46;                               // version
  d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32, A / B tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n"
               " tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
               ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
               "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                 "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                 "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr));
}

struct GemmArgs {
  int M, N, K;
  const float* bias;     // [N] or NULL
  int activation;        // B2S_ACT_*
  float* c;              // [M, N] row-major
  float* c_lo;           // optional: c - tf32(c), the lo operand of a following projection
};

__device__ __forceinline__ float activate(float x, int act) {
  if (act == B2S_ACT_RELU) return fmaxf(x, 0.f);
  if (act == B2S_ACT_SIGMOID) return 1.f / (1.f + __expf(-x));
  return x;
}
__device__ __forceinline__ float tf32_lo(float x) { return x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

template <int PRODUCTS>
__global__ void __launch_bounds__(kThreads, 1)
linear_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_a_lo,
                   const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_w_lo,
                   const GemmArgs args) {
  constexpr int kStages = stages_for(PRODUCTS);
  constexpr int kStageBytes = tiles_per_stage(PRODUCTS) * kTileBytes;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the swizzle atoms need 1024-byte aligned tiles
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_holder;
  __shared__ float epi_stage[4][32 * 33];   // per epilogue warp: transposition tile of a 32 x 32 chunk
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int tiles_m = (args.M + kBlockM - 1) / kBlockM, tiles_n = (args.N + kBlockN - 1) / kBlockN;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (args.K + kBlockK - 1) / kBlockK;     // the tensor maps zero-fill beyond K

  if (warp == 0 && lane == 0) {
    prefetch_map(&map_a); prefetch_map(&map_w);
    if (PRODUCTS == 3) { prefetch_map(&map_a_lo); prefetch_map(&map_w_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full[a], 1); mbar_init(&tmem_empty[a], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {   // whole warp: allocate the tensor memory, publish its base address
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                 ::"r"(smem_u32(&tmem_base_holder)), "r"(kTmemColumns) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0; unsigned phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* st = smem + stage * kStageBytes;
          mbar_expect_tx(&full_bar[stage], kStageBytes);
          tma_load_2d(st, &map_a, kb * kBlockK, m0, &full_bar[stage]);
          tma_load_2d(st + kTileBytes, &map_w, kb * kBlockK, n0, &full_bar[stage]);
          if (PRODUCTS == 3) {
            tma_load_2d(st + 2 * kTileBytes, &map_a_lo, kb * kBlockK, m0, &full_bar[stage]);
            tma_load_2d(st + 3 * kTileBytes, &map_w_lo, kb * kBlockK, n0, &full_bar[stage]);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc(kBlockM, kBlockN);
      int stage = 0; unsigned phase = 0;
      int acc = 0; unsigned acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);     // the epilogue has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem_base + acc * kAccumColumns;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t st = smem_u32(smem + stage * kStageBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint32_t off = k * kUmmaK * 4;     // 32 bytes along K inside the swizzle atom
            const uint64_t a = umma_desc(st + off), w = umma_desc(st + kTileBytes + off);
            umma_tf32(d, a, w, idesc, (kb | k) != 0);
            if (PRODUCTS == 3) {
              const uint64_t a_lo = umma_desc(st + 2 * kTileBytes + off), w_lo = umma_desc(st + 3 * kTileBytes + off);
              umma_tf32(d, a, w_lo, idesc, 1);
              umma_tf32(d, a_lo, w, idesc, 1);
            }
          }
          umma_commit(&empty_bar[stage]);            // the stage is free once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);                // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: warp w reads TMEM lanes 32 (w % 4) .. +31 = rows of the tile =====
    const int quarter = warp & 3;
    int acc = 0; unsigned acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / tiles_n) * kBlockM, n0 = (tile % tiles_n) * kBlockN;
      mbar_wait(&tmem_full[acc], acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + acc * kAccumColumns + ((uint32_t)(quarter * 32) << 16);
      // A thread of the warp receives ONE accumulator row (32 consecutive columns per tcgen05.ld): stored as they
      // come, a warp instruction would touch 32 different rows.  The 32 x 32 chunk is transposed through a padded
      // shared-memory tile instead: every global store instruction then writes one 128-byte row segment.
      float* stage = epi_stage[quarter];
#pragma unroll 1
      for (int c0 = 0; c0 < kBlockN; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(taddr + c0, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float x = __uint_as_float(v[j]);
          const int n = n0 + c0 + j;
          if (args.bias && n < args.N) x += __ldg(args.bias + n);
          stage[lane * 33 + j] = activate(x, args.activation);
        }
        __syncwarp();
        const int n = n0 + c0 + lane;
        if (n < args.N) {
          const int rows = min(32, args.M - (m0 + quarter * 32));
          float* cp = args.c + (int64_t)(m0 + quarter * 32) * args.N + n;
          float* lp = args.c_lo ? args.c_lo + (int64_t)(m0 + quarter * 32) * args.N + n : nullptr;
          for (int r = 0; r < rows; ++r) {
            const float o = stage[r * 33 + lane];
            cp[(int64_t)r * args.N] = o;
            if (lp) lp[(int64_t)r * args.N] = tf32_lo(o);
          }
        }
        __syncwarp();
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemColumns) : "memory");
  }
}

// x_lo = x - tf32(x): the second operand of the 3-term split
__global__ void __launch_bounds__(256)
tf32_split_kernel(const float* __restrict__ x, float* __restrict__ lo, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i + 4 <= n && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(lo)) & 15) == 0) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + i));
    *reinterpret_cast<float4*>(lo + i) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  } else {
    for (int64_t j = i; j < n && j < i + 4; ++j) lo[j] = tf32_lo(x[j]);
  }
}

// ---- tensor maps (driver entry point fetched at run time: the library does not link against libcuda) ------------
typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiled encode_tiled() {
  static EncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiled>(p);
  });
  return fn;
}

// [rows, cols] fp32 row-major with `row_stride` floats between rows; box = kBlockK columns x 128 rows
int make_map(CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t row_stride) {
  EncodeTiled enc = encode_tiled();
  B2S_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 4};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockM};
  const cuuint32_t elem[2] = {1, 1};
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, elem,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2S_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a [%lld, %lld] operand", (int)r,
              (long long)rows, (long long)cols);
  return B2S_OK;
}

template <int PRODUCTS>
int launch_linear(const CUtensorMap& ma, const CUtensorMap& mal, const CUtensorMap& mw, const CUtensorMap& mwl,
                  const GemmArgs& args, cudaStream_t stream) {
  constexpr size_t smem = (size_t)stages_for(PRODUCTS) * tiles_per_stage(PRODUCTS) * kTileBytes + 1024;
  static_assert(smem <= 227 * 1024, "stage ring exceeds the shared memory of an SM");
  auto kernel = linear_umma_kernel<PRODUCTS>;
  static bool configured[64] = {};
  int dev = 0;
  B2S_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2S_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63] = true;
  }
  const int tiles = ((args.M + kBlockM - 1) / kBlockM) * ((args.N + kBlockN - 1) / kBlockN);
  const int grid = std::max(1, std::min(tiles, kNumSMs));
  kernel<<<grid, kThreads, smem, stream>>>(ma, mal, mw, mwl, args);
  B2S_LAUNCH_CHECK("linear_umma_kernel");
  return B2S_OK;
}

}  // namespace

extern "C" {

int b2s_tf32_split(const float* x, int64_t count, float* lo, b2s_stream stream) {
  B2S_REQUIRE(count >= 0, "negative count");
  if (count == 0) return B2S_OK;
  B2S_REQUIRE(x && lo, "NULL device pointer");
  const int64_t blocks = ceil_div(count, 1024);
  B2S_REQUIRE(blocks < ((int64_t)1 << 31), "too many elements for one launch");
  tf32_split_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, lo, count);
  B2S_LAUNCH_CHECK("tf32_split_kernel");
  return B2S_OK;
}

int b2s_linear_forward(const float* a, const float* a_lo, const float* weight, const float* weight_lo,
                       const float* bias, int64_t m, int64_t n, int64_t k, int64_t a_row_stride,
                       int64_t weight_row_stride, int activation, float* c, float* c_lo, b2s_stream stream) {
  B2S_REQUIRE(m >= 0 && n >= 1 && k >= 1 && m < ((int64_t)1 << 31) && n < ((int64_t)1 << 31) && k < ((int64_t)1 << 31),
              "bad extents (m=%lld n=%lld k=%lld)", (long long)m, (long long)n, (long long)k);
  B2S_REQUIRE(activation >= B2S_ACT_NONE && activation <= B2S_ACT_SIGMOID, "unknown activation %d", activation);
  if (m == 0) return B2S_OK;
  B2S_REQUIRE(a && weight && c, "NULL device pointer");
  B2S_REQUIRE((a_lo == nullptr) == (weight_lo == nullptr),
              "pass both lo operands (3-term split, fp32-faithful) or neither (plain TF32)");
  // TMA: 16-byte aligned base addresses and row pitches
  B2S_REQUIRE(a_row_stride >= k && weight_row_stride >= k && a_row_stride % 4 == 0 && weight_row_stride % 4 == 0,
              "row strides must be multiples of 4 floats (got %lld, %lld): pad K", (long long)a_row_stride,
              (long long)weight_row_stride);
  B2S_REQUIRE(((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(weight) | reinterpret_cast<uintptr_t>(a_lo) |
                reinterpret_cast<uintptr_t>(weight_lo)) & 15) == 0, "operands must be 16-byte aligned");
  CUtensorMap ma, mal, mw, mwl;
  if (int rc = make_map(&ma, a, m, k, a_row_stride)) return rc;
  if (int rc = make_map(&mw, weight, n, k, weight_row_stride)) return rc;
  const GemmArgs args{(int)m, (int)n, (int)k, bias, activation, c, c_lo};
  if (a_lo) {
    if (int rc = make_map(&mal, a_lo, m, k, a_row_stride)) return rc;
    if (int rc = make_map(&mwl, weight_lo, n, k, weight_row_stride)) return rc;
    return launch_linear<3>(ma, mal, mw, mwl, args, (cudaStream_t)stream);
  }
  return launch_linear<1>(ma, ma, mw, mw, args, (cudaStream_t)stream);
}

}  // extern "C"
