// Warp-specialised persistent GEMM for sm_100a: TMA -> shared memory (128-byte
// swizzle) -> tcgen05.mma (BF16 x BF16 -> FP32 accumulator in TMEM) -> tcgen05.ld
// epilogue with the fused element-wise layers.
//
// Replaces GEMM<float>::Gemm (src/gemm.cc:69-125) and its AVX2 micro-kernel
// (src/gemm_haswell.cc:72-632) for LinearLayer::Propagate (src/nnet.cc:22-36), with
// ReLULayer / NormalizeLayer (src/nnet.cc:49-75) and the exp/sum half of
// SoftmaxLayer (src/vector.cc:264-277) folded into the epilogue. Weights are used
// in their on-disk [out][in] layout (K-major for both operands), so the transpose
// of LinearLayer::LinearLayer (src/nnet.cc:16-17) disappears.
//
// CTA = 8 warps: warp 0 lane 0 TMA producer, warp 1 lane 0 MMA issuer, warp 2 TMEM
// allocator, warps 4-7 epilogue (TMEM lane quadrant = warp % 4). Three pipelines:
// smem full/empty ring (TMA <-> MMA), TMEM full/empty double buffer (MMA <->
// epilogue), static persistent tile loop with N fastest so that co-resident CTAs
// share the activation tile in L2.
//
// BF16X3: every operand is carried as two BF16 planes (hi, lo = bf16(x - hi)) and
// each K step issues hi*hi + lo*hi + hi*lo into the same accumulator, which brings
// the product error down to ~2^-16 relative (FP32-class for this workload).

#include <algorithm>

#include "gemm_sm100.cuh"

namespace pkb {

namespace {

constexpr int kThreads = 256;
constexpr int kMaxStages = 8;
constexpr uint32_t kSmemBudget = 227 * 1024;

// ---------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// Bounded wait: a pipeline bug turns into a trap (reported as a launch failure)
// instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000ll) __trap();
  }
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar,
                                            int c_inner, int c_outer) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
      : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t *dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, BF16 inputs, FP32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Arrives on `bar` once all previously issued tcgen05.mma of this thread finished.
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
        "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes
// apart (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// cute::UMMA::InstrDescriptor for kind::f16: D=F32 (1<<4), A=B=BF16 (1<<7, 1<<10),
// both K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t *>(&v);
}

struct SmemLayout {
  uint32_t stage_bytes, a_plane, w_plane, stages, bar_off, total;
};

__host__ __device__ inline SmemLayout smem_layout(int block_n, int planes) {
  SmemLayout L;
  L.a_plane = kBlockM * kBlockK * 2;
  L.w_plane = block_n * kBlockK * 2;
  L.stage_bytes = planes * (L.a_plane + L.w_plane);
  uint32_t avail = kSmemBudget - 1024 /* alignment slack */ - 256 /* barriers */;
  L.stages = avail / L.stage_bytes;
  if (L.stages > kMaxStages) L.stages = kMaxStages;
  L.bar_off = L.stages * L.stage_bytes;
  L.total = L.bar_off + 256 + 1024;
  return L;
}

// ---------------------------------------------------------------- kernel
template <int BN, int PLANES, bool FINAL>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tm_a_hi, const __grid_constant__ CUtensorMap tm_a_lo,
            const __grid_constant__ CUtensorMap tm_w_hi, const __grid_constant__ CUtensorMap tm_w_lo,
            const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const SmemLayout L = smem_layout(BN, PLANES);
  uint8_t *smem = reinterpret_cast<uint8_t *>(
      (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bar_off);
  uint64_t *full = bars;                       // [stages]
  uint64_t *empty = bars + kMaxStages;         // [stages]
  uint64_t *tfull = bars + 2 * kMaxStages;     // [2]
  uint64_t *tempty = bars + 2 * kMaxStages + 2;  // [2]
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * kMaxStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t S = L.stages;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a_hi);
    tma_prefetch_desc(&tm_w_hi);
    if (PLANES == 2) {
      tma_prefetch_desc(&tm_a_lo);
      tma_prefetch_desc(&tm_w_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < S; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 2 * BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int m0 = (tile / p.n_tiles_n) * kBlockM;
        const int n0 = (tile % p.n_tiles_n) * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t *st = smem + s * L.stage_bytes;
          mbar_expect_tx(&full[s], L.stage_bytes);
          tma_load_2d(st, &tm_a_hi, &full[s], kb * kBlockK, m0);
          tma_load_2d(st + PLANES * L.a_plane, &tm_w_hi, &full[s], kb * kBlockK, n0);
          if (PLANES == 2) {
            tma_load_2d(st + L.a_plane, &tm_a_lo, &full[s], kb * kBlockK, m0);
            tma_load_2d(st + 2 * L.a_plane + L.w_plane, &tm_w_lo, &full[s], kb * kBlockK, n0);
          }
          if (++s == S) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(kBlockM, BN);
      uint32_t s = 0, ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const uint32_t as = it & 1, aph = (it >> 1) & 1;
        mbar_wait(&tempty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * L.stage_bytes);
          const uint32_t sw = sa + PLANES * L.a_plane;
          const uint64_t a_hi = make_smem_desc(sa);
          const uint64_t w_hi = make_smem_desc(sw);
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
            umma_bf16(d_tmem, a_hi + adv, w_hi + adv, idesc, (kb | k) != 0);
          }
          if (PLANES == 2) {
            const uint64_t a_lo = make_smem_desc(sa + L.a_plane);
            const uint64_t w_lo = make_smem_desc(sw + L.w_plane);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t adv = static_cast<uint64_t>((k * kUmmaK * 2) >> 4);
              umma_bf16(d_tmem, a_lo + adv, w_hi + adv, idesc, 1);
              umma_bf16(d_tmem, a_hi + adv, w_lo + adv, idesc, 1);
            }
          }
          umma_commit(&empty[s]);  // frees the smem stage once these MMAs retire
          if (++s == S) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull[as]);  // accumulator complete
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;  // TMEM lane quadrant
    int it = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const uint32_t as = it & 1, aph = (it >> 1) & 1;
      const int n_blk = tile % p.n_tiles_n;
      const int m0 = (tile / p.n_tiles_n) * kBlockM;
      const int n0 = n_blk * BN;
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < p.M;

      float rs = 1.0f;
      if (p.in_sumsq != nullptr && row_ok) {
        float ss = 0.0f;
        for (int i = 0; i < p.in_sumsq_tiles; ++i)
          ss += p.in_sumsq[static_cast<size_t>(row) * p.in_sumsq_tiles + i];
        rs = sqrtf(p.in_dim / ss);  // NormalizeLayer: no floor (src/nnet.cc:71-73)
      }

      mbar_wait(&tfull[as], aph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BN;

      float sumsq = 0.0f;
      float run_max = -INFINITY, run_sum = 0.0f;
      int dest = -1;
      if (FINAL && row_ok) dest = p.row_map ? p.row_map[row] : row;

#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t v[32];
        tmem_ld32(taddr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;
        float z[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4 *>(p.bias + col0 + i));
          z[i + 0] = fmaf(__uint_as_float(v[i + 0]), rs, b.x);
          z[i + 1] = fmaf(__uint_as_float(v[i + 1]), rs, b.y);
          z[i + 2] = fmaf(__uint_as_float(v[i + 2]), rs, b.z);
          z[i + 3] = fmaf(__uint_as_float(v[i + 3]), rs, b.w);
        }
        if (!FINAL) {
          if (p.relu) {
#pragma unroll
            for (int i = 0; i < 32; ++i) z[i] = fmaxf(z[i], 0.0f);
          }
          if (p.out_sumsq != nullptr) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sumsq = fmaf(z[i], z[i], sumsq);
          }
          if (row_ok) {
            uint32_t hi[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) hi[i] = pack_bf16(z[2 * i], z[2 * i + 1]);
            uint4 *dst = reinterpret_cast<uint4 *>(p.out_hi + static_cast<size_t>(row) * p.ld_out + col0);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              dst[i] = make_uint4(hi[4 * i], hi[4 * i + 1], hi[4 * i + 2], hi[4 * i + 3]);
            if (PLANES == 2) {
              uint32_t lo[16];
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162 *>(&hi[i]);
                lo[i] = pack_bf16(z[2 * i] - __low2float(h), z[2 * i + 1] - __high2float(h));
              }
              uint4 *dl = reinterpret_cast<uint4 *>(p.out_lo + static_cast<size_t>(row) * p.ld_out + col0);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                dl[i] = make_uint4(lo[4 * i], lo[4 * i + 1], lo[4 * i + 2], lo[4 * i + 3]);
            }
          }
        } else {
          const int nvalid = p.N_valid - col0;  // columns of this chunk that are real
          if (p.lse_part != nullptr && nvalid > 0) {
            float cmax = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) cmax = fmaxf(cmax, z[i]);
            const float nm = fmaxf(run_max, cmax);
            float acc = 0.0f;
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (i < nvalid) acc += __expf(z[i] - nm);
            run_sum = run_sum * __expf(run_max - nm) + acc;
            run_max = nm;
          }
          if (dest >= 0 && nvalid > 0) {
            float *dst = p.out_f32 + static_cast<size_t>(dest) * p.ld_f32 + col0;
            if (nvalid >= 32 && (p.ld_f32 & 3) == 0) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4 *>(dst + i) = make_float4(z[i], z[i + 1], z[i + 2], z[i + 3]);
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nvalid) dst[i] = z[i];
            }
          }
        }
      }
      // all TMEM reads of this accumulator are done: hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[as]);

      if (!FINAL) {
        if (p.out_sumsq != nullptr && row_ok)
          p.out_sumsq[static_cast<size_t>(row) * p.n_tiles_n + n_blk] = sumsq;
      } else {
        if (p.lse_part != nullptr && row_ok)
          p.lse_part[static_cast<size_t>(row) * p.n_tiles_n + n_blk] = make_float2(run_max, run_sum);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 2 * BN);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void *sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || sym == nullptr)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(sym);
  return fn;
}

template <int BN, int PLANES, bool FINAL>
int launch_one(Ctx *c, const CUtensorMap *a_hi, const CUtensorMap *a_lo, const CUtensorMap *w_hi,
               const CUtensorMap *w_lo, const GemmParams &p) {
  const SmemLayout L = smem_layout(BN, PLANES);
  auto kern = gemm_kernel<BN, PLANES, FINAL>;
  PKB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
  const int grid = std::min(p.num_tiles, c->sm_count);
  LaunchScope scope(c, PKB_KERNEL_GEMM);
  kern<<<grid, kThreads, L.total, c->stream>>>(*a_hi, *a_lo, *w_hi, *w_lo, p);
  PKB_CUDA(cudaGetLastError());
  return PKB_OK;
}

}  // namespace

int gemm_max_smem_bytes(int block_n, int planes) { return smem_layout(block_n, planes).total; }

int make_tensor_map(CUtensorMap *map, const void *base, uint64_t cols, uint64_t rows,
                    uint64_t pitch_bytes, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return PKB_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d) cols=%llu rows=%llu pitch=%llu", (int)r,
              (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_bytes);
    return PKB_ERR_CUDA;
  }
  return PKB_OK;
}

int launch_gemm(Ctx *c, int block_n, int planes, bool final, const CUtensorMap *a_hi,
                const CUtensorMap *a_lo, const CUtensorMap *w_hi, const CUtensorMap *w_lo,
                const GemmParams &p) {
  if (p.num_tiles <= 0) return PKB_OK;
  if (planes == 1) { a_lo = a_hi; w_lo = w_hi; }
#define PKB_GEMM_CASE(BN, PL, FN) \
  if (block_n == BN && planes == PL && final == FN) return launch_one<BN, PL, FN>(c, a_hi, a_lo, w_hi, w_lo, p);
  PKB_GEMM_CASE(128, 1, false)
  PKB_GEMM_CASE(128, 1, true)
  PKB_GEMM_CASE(128, 2, false)
  PKB_GEMM_CASE(128, 2, true)
  PKB_GEMM_CASE(256, 1, false)
  PKB_GEMM_CASE(256, 1, true)
  PKB_GEMM_CASE(256, 2, false)
  PKB_GEMM_CASE(256, 2, true)
#undef PKB_GEMM_CASE
  set_error("launch_gemm: unsupported configuration block_n=%d planes=%d", block_n, planes);
  return PKB_ERR_INVALID;
}

}  // namespace pkb
