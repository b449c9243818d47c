// Tensor-core implicit-GEMM convolution for sm_100a:  TMA (im2col-free, OOB zero fill = conv padding)
//   -> 128B-swizzled shared-memory stages -> tcgen05.mma (bf16 x bf16 -> fp32 accumulators in TMEM)
//   -> tcgen05.ld epilogue (+bias, +residual) -> fp32 NHWC.
//
// GEMM view per CTA: D[128 pixels][BN couts] = sum over (tap, 64-channel block) of
//     A_tap[128 pixels][64 ch] * W_tap[BN couts][64 ch]^T
// The 128-pixel M tile is a (tn images) x (th rows) x (tw cols) box of the NHWC activation, fetched by
// ONE 4-D TMA per (tap, channel block) at coordinates shifted by the tap offset; out-of-image
// coordinates are zero-filled by the TMA unit, which is exactly the conv's zero padding.  Both operands
// land K-major with the 128-byte swizzle, i.e. the canonical UMMA SW128 layout (8-row atoms of 1024 B).
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer,
// warps 2..5 = epilogue (warp%4 selects the 32-lane TMEM quadrant it may read).
#include <cuda.h>

#include "common.cuh"

namespace pdae {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                         // bf16 elements = 128 bytes = one swizzle row
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;     // 16 KB
constexpr int TC_MAX_STAGES = 8;
constexpr int TC_THREADS = 192;

struct ConvTcArgs {
  const float* bias;
  const float* residual;
  float* out;
  int B, H, W, Cout;
  int tw, th, tn;        // pixel box of one M tile (tw*th*tn == 128)
  int tiles_x, tiles_y;  // tiles per image row / column
  int taps, ksize, kblocks;  // kblocks = Cin/64
  int stages;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must surface as a trapped kernel (error code), never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  const long long t0 = clock64();
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (=1, unused for swizzled K-major) | [32,46) SBO>>4 (=64: 8 rows * 128 B)
//   [46,48) version=1 (Blackwell) | [61,64) layout type 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

template <int BN>
__global__ void __launch_bounds__(TC_THREADS) conv_tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                             const __grid_constant__ CUtensorMap tmB, ConvTcArgs p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_empty[TC_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc;
  __shared__ uint32_t tmem_slot;

  constexpr int B_BYTES = BN * TC_BK * 2;
  constexpr int STAGE_BYTES = TC_A_BYTES + B_BYTES;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SW128 atoms need 1024-B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.stages;

  // ---- tile coordinates ----
  int mt = blockIdx.x;
  const int tx = mt % p.tiles_x;
  mt /= p.tiles_x;
  const int ty = mt % p.tiles_y;
  const int bt = mt / p.tiles_y;
  const int x0 = tx * p.tw, y0 = ty * p.th, b0 = bt * p.tn;
  const int n0 = blockIdx.y * BN;
  const int total_k = p.taps * p.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(smem_u32(&bar_full[s]), 1);
      mbar_init(smem_u32(&bar_empty[s]), 1);
    }
    mbar_init(smem_u32(&bar_acc), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ===== TMA producer =====
      for (int it = 0; it < total_k; ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        mbar_wait(smem_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = smem_u32(&bar_full[s]);
        mbar_expect_tx(full, STAGE_BYTES);
        const int tap = it / p.kblocks, kb = it - tap * p.kblocks;
        const int dy = p.ksize == 3 ? tap / 3 - 1 : 0, dx = p.ksize == 3 ? tap % 3 - 1 : 0;
        const uint32_t sa = smem0 + (uint32_t)s * STAGE_BYTES;
        tma_load_4d(sa, &tmA, full, kb * TC_BK, x0 + dx, y0 + dy, b0);
        tma_load_3d(sa + TC_A_BYTES, &tmB, full, kb * TC_BK, n0, tap);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===== MMA issuer =====
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=F32 [4,6)=1, a=BF16 [7,10)=1, b=BF16 [10,13)=1,
      // a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      for (int it = 0; it < total_k; ++it) {
        const int s = it % S;
        const uint32_t ph = (uint32_t)((it / S) & 1);
        mbar_wait(smem_u32(&bar_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + (uint32_t)s * STAGE_BYTES;
        const uint64_t adesc = make_sw128_desc(sa);
        const uint64_t bdesc = make_sw128_desc(sa + TC_A_BYTES);
#pragma unroll
        for (int k = 0; k < TC_BK / 16; ++k) {
          // advance 16 bf16 = 32 bytes along K inside the swizzle atom: +2 in the (addr>>4) field
          umma_bf16(tmem_base, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), IDESC, (uint32_t)((it | k) != 0));
        }
        umma_commit(smem_u32(&bar_empty[s]));  // frees this smem stage when the MMAs above retire
      }
      umma_commit(smem_u32(&bar_acc));  // accumulator complete
    }
  } else {
    // ===== epilogue: TMEM -> registers -> (+bias, +residual) -> global fp32 NHWC =====
    const int q = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = q * 32 + lane;       // accumulator row == pixel index inside the tile
    const int ni = r / (p.th * p.tw);
    const int rem = r - ni * (p.th * p.tw);
    const int yy = rem / p.tw, xx = rem - yy * p.tw;
    const int b = b0 + ni;
    const bool valid = b < p.B;
    const long long pix = ((long long)b * p.H + (y0 + yy)) * p.W + (x0 + xx);
    float* __restrict__ orow = p.out + pix * p.Cout + n0;
    const float* __restrict__ rrow = p.residual ? p.residual + pix * p.Cout + n0 : nullptr;
    mbar_wait(smem_u32(&bar_acc), 0u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
            "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
            "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
            "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (valid) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 o;
          o.x = __uint_as_float(v[j + 0]);
          o.y = __uint_as_float(v[j + 1]);
          o.z = __uint_as_float(v[j + 2]);
          o.w = __uint_as_float(v[j + 3]);
          if (p.bias) {
            const float4 bv = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + c0 + j));
            o.x += bv.x; o.y += bv.y; o.z += bv.z; o.w += bv.w;
          }
          if (rrow) {
            const float4 rv = *reinterpret_cast<const float4*>(rrow + c0 + j);
            o.x += rv.x; o.y += rv.y; o.z += rv.z; o.w += rv.w;
          }
          *reinterpret_cast<float4*>(orow + c0 + j) = o;
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn)sym;
  return fn;
}

}  // namespace pdae

using namespace pdae;

struct pdae_conv_tc_plan {
  CUtensorMap tmA, tmB;
  ConvTcArgs args;
  int BN;
  dim3 grid;
  size_t smem;
};

static int pick_pow2_tile(int W, int cap) {
  int t = 1;
  while (t * 2 <= cap && W % (t * 2) == 0) t *= 2;
  return t;
}

extern "C" int pdae_conv_tc_create(pdae_conv_tc_plan** plan_out, const void* in_bf16, const void* w_bf16,
                                   const float* bias, const float* residual, float* out, int B, int H, int W, int Cin,
                                   int Cout, int ksize) {
  PDAE_REQUIRE(plan_out && in_bf16 && w_bf16 && out, "conv_tc_create: null pointer");
  PDAE_REQUIRE(ksize == 1 || ksize == 3, "conv_tc_create: ksize must be 1 or 3");
  PDAE_REQUIRE(Cin % TC_BK == 0 && Cout % 64 == 0, "conv_tc_create: Cin=%d Cout=%d not multiples of 64", Cin, Cout);
  PDAE_REQUIRE(((uintptr_t)in_bf16 & 15) == 0 && ((uintptr_t)w_bf16 & 15) == 0 && ((uintptr_t)out & 15) == 0,
               "conv_tc_create: pointers must be 16-byte aligned");
  EncodeTiledFn enc = get_encode_fn();
  PDAE_REQUIRE(enc != nullptr, "conv_tc_create: cuTensorMapEncodeTiled unavailable (no driver)");

  pdae_conv_tc_plan* pl = new pdae_conv_tc_plan();
  ConvTcArgs& a = pl->args;
  a.bias = bias; a.residual = residual; a.out = out;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout;
  a.tw = pick_pow2_tile(W, TC_BM);
  a.th = pick_pow2_tile(H, TC_BM / a.tw);
  a.tn = TC_BM / (a.tw * a.th);
  if (W % a.tw != 0 || H % a.th != 0 || a.tw * a.th * a.tn != TC_BM || a.tn > 256) {
    delete pl;
    PDAE_REQUIRE(false, "conv_tc_create: H=%d W=%d cannot be tiled into 128-pixel boxes", H, W);
  }
  a.tiles_x = W / a.tw; a.tiles_y = H / a.th;
  a.taps = ksize * ksize; a.ksize = ksize; a.kblocks = Cin / TC_BK;
  pl->BN = (Cout % 128 == 0) ? 128 : 64;
  const int stage_bytes = TC_A_BYTES + pl->BN * TC_BK * 2;
  a.stages = 3;
  if (a.taps * a.kblocks < a.stages) a.stages = a.taps * a.kblocks;
  pl->smem = (size_t)a.stages * stage_bytes + 1024;
  const int b_tiles = (B + a.tn - 1) / a.tn;
  pl->grid = dim3((unsigned)(a.tiles_x * a.tiles_y * b_tiles), (unsigned)(Cout / pl->BN), 1);

  {  // activations: [B][H][W][Cin] bf16, box (64 ch, tw, th, tn)
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)Cin * 2, (cuuint64_t)W * Cin * 2, (cuuint64_t)H * W * Cin * 2};
    cuuint32_t box[4] = {(cuuint32_t)TC_BK, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(in_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      PDAE_REQUIRE(false, "conv_tc_create: cuTensorMapEncodeTiled(A) failed with %d", (int)r);
    }
  }
  {  // weights: [taps][Cout][Cin] bf16, box (64 ch, BN couts, 1 tap)
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)a.taps};
    cuuint64_t strides[2] = {(cuuint64_t)Cin * 2, (cuuint64_t)Cout * Cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)pl->BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(w_bf16), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      PDAE_REQUIRE(false, "conv_tc_create: cuTensorMapEncodeTiled(W) failed with %d", (int)r);
    }
  }
  // opt in to the largest dynamic shared-memory footprint any plan of this BN can ask for (3 stages)
  const int max_smem = 3 * stage_bytes + 1024;
  cudaError_t e = pl->BN == 128
                      ? cudaFuncSetAttribute(conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem)
                      : cudaFuncSetAttribute(conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  if (e != cudaSuccess) {
    delete pl;
    set_error("conv_tc_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return PDAE_ECUDA;
  }
  *plan_out = pl;
  return PDAE_OK;
}

extern "C" int pdae_conv_tc_run(const pdae_conv_tc_plan* pl, pdae_stream_t stream) {
  PDAE_REQUIRE(pl, "conv_tc_run: null plan");
  cudaStream_t s = (cudaStream_t)stream;
  if (pl->BN == 128)
    conv_tc_kernel<128><<<pl->grid, TC_THREADS, pl->smem, s>>>(pl->tmA, pl->tmB, pl->args);
  else
    conv_tc_kernel<64><<<pl->grid, TC_THREADS, pl->smem, s>>>(pl->tmA, pl->tmB, pl->args);
  PDAE_LAUNCH_CHECK("conv_tc_kernel");
  return PDAE_OK;
}

extern "C" void pdae_conv_tc_destroy(pdae_conv_tc_plan* pl) { delete pl; }
