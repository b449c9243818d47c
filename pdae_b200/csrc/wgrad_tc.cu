// wgrad_tc -- weight gradient of a stride-1 3x3 / 1x1 "same" convolution on the tensor cores (sm_100a), fp32-grade.
//
//   dW[tap][cin][cout] = sum over (b, y, x) of  act[b, y+dy, x+dx, cin] * dY[b, y, x, cout]        (autograd of F.conv2d,
//   model/module.py:241-243, 255-259 under trainer/train_representation_learning.py:112 loss.backward())
//
// As a GEMM the contraction runs over PIXELS while both operands are NHWC (channels contiguous): both are MN-major UMMA
// operands.  A TMA box [64 pixels][64 channels] with SWIZZLE_128B is exactly the canonical MN-major SW128 atom stack
// (8 pixel rows x 128 B per atom, stride-byte-offset 1024 B between 8-row groups); a second box one leading-byte-offset
// further supplies channels 64-127.  One MMA = 128 (M channels) x BN (N channels) x 16 pixels.
//
// Operands arrive split, [hi | lo | hi] channel blocks (a = hi + lo, bf16 each): every product is
// a_hi*d_hi + a_lo*d_hi + a_hi*d_lo accumulated in fp32 in TMEM -- the same fp32-grade scheme as the forward / dgrad convs.
//
// Work item = (tap, M chunk of 128 channels, N chunk of BN channels, 64-pixel tile).  Items are dealt to the persistent CTAs
// in contiguous ranges (split-K over pixels); a CTA accumulates in TMEM while consecutive items belong to the same
// (tap, M chunk, N chunk) and then adds its partial sums to dW with fp32 reductions.
//   warp 4: TMA producer, warp 5: MMA issuer (warp-uniform loops, elect.sync), warps 0-3: epilogue (one TMEM lane quadrant each).
// The tap shift is applied to whichever operand is the activation (4-D box at shifted coordinates, out-of-image = zero = padding).
// `a_is_act` selects which tensor sits on the M side: with dY there (M = cout) the 32 lanes of a warp reduce into 32
// consecutive dW addresses (coalesced), which is the preferred form whenever Cout % 128 == 0.  When neither channel count is a
// multiple of 128 (64 -> 64, 192 -> 64) the M side is the activation in 64-channel chunks and the two 64-row halves of the
// accumulator hold two different TAPS of it ("pair" mode; the odd ninth tap is paired with a discarded duplicate).
#include <cuda.h>

#include "common.cuh"

namespace pdae {

constexpr int WG_KT = 64;                 // pixels per k-tile (= rows of one TMA box)
constexpr int WG_BOX = WG_KT * 128;       // bytes of one [64 px][64 ch] box
constexpr int WG_THREADS = 192;
constexpr int WG_MAX_ST = 4;

struct WgradArgs {
  float* dw;
  long long sm, sn, stap;   // element strides of the (m, n, tap) indices inside dw
  int a_is_act;             // 1: M side = activation (shifted per tap), N side = dY; 0: M side = dY, N side = activation
  int Ma, Nb;               // real channel counts on the M / N side (the tensors hold 3x: [hi | lo | hi])
  int mchunks, nchunks, taps, ksize;
  int pair;                 // 1: 64-channel M chunks, the 128 accumulator rows hold TWO taps (activation on the M side)
  int tw, th, tn, tiles_x, tiles_y, tiles_b, ktiles;
  int B, H, W;
  int stages;
  long long items;          // taps * mchunks * nchunks * ktiles
};

namespace wg {
__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mb_try(uint32_t bar, uint32_t parity, uint32_t hint) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(hint)
      : "memory");
  return done;
}
__device__ __noinline__ void mb_wait(uint32_t bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mb_try(bar, parity, 20000u))
    if (++n > 4000000u) __trap();  // a protocol bug must trap, never hang the GPU
}
__device__ __forceinline__ void tma_ld4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// MN-major SWIZZLE_128B operand: 64-element (128 B) channel blocks `lbo` bytes apart, 8-pixel row groups 1024 B apart
__device__ __forceinline__ uint64_t mn_desc(uint32_t saddr, uint32_t lbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) |
         (2ull << 61);
}
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a), "l"(b), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_commit_to(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
}  // namespace wg

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, WgradArgs p) {
  using namespace wg;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[WG_MAX_ST], bar_empty[WG_MAX_ST], bar_acc_full, bar_acc_empty;
  __shared__ uint32_t tmem_slot;
  constexpr int NBOX_B = BN / 64;                       // 64-channel boxes of the N-side tile
  constexpr int A_BYTES = 2 * WG_BOX;                   // M = 128 channels = two boxes
  constexpr int B_BYTES = NBOX_B * WG_BOX;
  constexpr int STAGE = 2 * A_BYTES + 2 * B_BYTES;      // (hi, lo) of both operands
  const uint32_t smem0 = (s_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long per_cta = (p.items + gridDim.x - 1) / gridDim.x;
  const long long it_begin = (long long)blockIdx.x * per_cta;
  const long long it_end = it_begin + per_cta < p.items ? it_begin + per_cta : p.items;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mb_init(s_u32(&bar_full[s]), 1);
      mb_init(s_u32(&bar_empty[s]), 1);
    }
    mb_init(s_u32(&bar_acc_full), 1);
    mb_init(s_u32(&bar_acc_empty), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 5) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "n"(BN < 32 ? 32 : BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 4) {
    // ================= TMA producer =================
    int s = 0;
    uint32_t ph = 0;
    for (long long it = it_begin; it < it_end; ++it) {
      const long long g = it / p.ktiles;
      int kt = (int)(it - g * p.ktiles);
      int gg = (int)g;
      const int nc = gg % p.nchunks; gg /= p.nchunks;
      const int mc = gg % p.mchunks;
      const int tap = gg / p.mchunks;
      const int tx = kt % p.tiles_x; kt /= p.tiles_x;
      const int ty = kt % p.tiles_y;
      const int bt = kt / p.tiles_y;
      const int x0 = tx * p.tw, y0 = ty * p.th, b0 = bt * p.tn;
      // the activation carries the tap shift; in pair mode the two M boxes are two taps of the same 64 channels
      int tj[2], ax[2], ay[2], am[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        tj[j] = p.pair ? min(2 * tap + j, p.taps - 1) : tap;
        const int dy = p.ksize == 3 ? tj[j] / 3 - 1 : 0, dx = p.ksize == 3 ? tj[j] % 3 - 1 : 0;
        ax[j] = p.a_is_act ? x0 + dx : x0;
        ay[j] = p.a_is_act ? y0 + dy : y0;
        am[j] = p.pair ? mc * 64 : mc * 128 + j * 64;
      }
      const int bdy = p.ksize == 3 ? tap / 3 - 1 : 0, bdx = p.ksize == 3 ? tap % 3 - 1 : 0;
      const int bx = p.a_is_act ? x0 : x0 + bdx, by = p.a_is_act ? y0 : y0 + bdy;
      const int n0 = nc * BN;
      mb_wait(s_u32(&bar_empty[s]), ph ^ 1u);
      const uint32_t full = s_u32(&bar_full[s]);
      const uint32_t base = smem0 + (uint32_t)(s * STAGE);
      if (elect_one()) {
        mb_expect_tx(full, (uint32_t)STAGE);
#pragma unroll
        for (int h = 0; h < 2; ++h) {          // hi block (channel offset 0), lo block (channel offset Ma / Nb)
#pragma unroll
          for (int j = 0; j < 2; ++j)
            tma_ld4(base + (uint32_t)(h * A_BYTES + j * WG_BOX), &tmA, full, h * p.Ma + am[j], ax[j], ay[j], b0);
#pragma unroll
          for (int j = 0; j < NBOX_B; ++j)
            tma_ld4(base + (uint32_t)(2 * A_BYTES + h * B_BYTES + j * WG_BOX), &tmB, full, h * p.Nb + n0 + j * 64, bx, by, b0);
        }
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1u; }
    }
  } else if (warp == 5) {
    // ================= MMA issuer =================
    constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(BN >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);   // fp32 accumulate, bf16 x bf16, A and B MN-major
    const uint32_t tmem_d = __shfl_sync(0xffffffffu, tmem_base, 0);
    int s = 0, ngroups = 0;
    uint32_t ph = 0;
    long long cur_g = -1;
    for (long long it = it_begin; it < it_end; ++it) {
      const long long g = it / p.ktiles;
      const bool first = g != cur_g;
      if (first) {
        if (cur_g >= 0) {                      // previous group complete: hand the accumulator to the epilogue ...
          if (elect_one()) umma_commit_to(s_u32(&bar_acc_full));
          __syncwarp();
        }
        if (ngroups > 0) {                     // ... and wait until it has been drained
          mb_wait(s_u32(&bar_acc_empty), (uint32_t)((ngroups - 1) & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        cur_g = g;
        ++ngroups;
      }
      mb_wait(s_u32(&bar_full[s]), ph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t base = smem0 + (uint32_t)(s * STAGE);
      const uint64_t a_hi = mn_desc(base, WG_BOX), a_lo = mn_desc(base + A_BYTES, WG_BOX);
      const uint64_t b_hi = mn_desc(base + 2 * A_BYTES, WG_BOX), b_lo = mn_desc(base + 2 * A_BYTES + B_BYTES, WG_BOX);
      const uint32_t bar_e = s_u32(&bar_empty[s]);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < WG_KT / 16; ++k) {           // 16 pixels per MMA = 16 rows x 128 B = 2048 B = 128 descriptor units
          const uint64_t o = (uint64_t)(k * 128);
          umma(tmem_d, a_hi + o, b_hi + o, IDESC, (uint32_t)(!(first && k == 0)));
          umma(tmem_d, a_lo + o, b_hi + o, IDESC, 1u);
          umma(tmem_d, a_hi + o, b_lo + o, IDESC, 1u);
        }
        umma_commit_to(bar_e);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    if (cur_g >= 0) {
      if (elect_one()) umma_commit_to(s_u32(&bar_acc_full));
      __syncwarp();
    }
  } else {
    // ================= epilogue: TMEM -> fp32 reductions into dW =================
    const int q = warp & 3;
    const int m = q * 32 + lane;               // accumulator row = channel on the M side
    const int et = threadIdx.x;                // 0..127
    int ngroups = 0;
    long long cur_g = -1;
    for (long long it = it_begin; it <= it_end; ++it) {
      const long long g = it < it_end ? it / p.ktiles : -2;
      if (g == cur_g) continue;
      if (cur_g >= 0) {
        mb_wait(s_u32(&bar_acc_full), (uint32_t)((ngroups - 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int gg = (int)cur_g;
        const int nc = gg % p.nchunks; gg /= p.nchunks;
        const int mc = gg % p.mchunks;
        const int tap = gg / p.mchunks;
        const int tap_w = p.pair ? 2 * tap + (m >> 6) : tap;             // pair mode: rows 64-127 belong to the second tap
        const int ch_m = p.pair ? mc * 64 + (m & 63) : mc * 128 + m;
        const bool live = tap_w < p.taps;
        float* dst = p.dw + (long long)tap_w * p.stap + (long long)ch_m * p.sm + (long long)(nc * BN) * p.sn;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(tacc + (uint32_t)(c * 32), v);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
          if (live) {
#pragma unroll
            for (int j = 0; j < 32; ++j) atomicAdd(dst + (long long)(c * 32 + j) * p.sn, __uint_as_float(v[j]));
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) mb_arrive(s_u32(&bar_acc_empty));
      }
      cur_g = g;
      if (g >= 0) ++ngroups;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 5) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BN < 32 ? 32 : BN) : "memory");
  }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnW encode_fnw() {
  static EncodeTiledFnW fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFnW)sym;
  return fn;
}

static int pow2_tile_w(int W, int cap) {
  int t = 1;
  while (t * 2 <= cap && W % (t * 2) == 0) t *= 2;
  return t;
}

template <int BN>
static cudaError_t launch_wg(const CUtensorMap& a, const CUtensorMap& b, const WgradArgs& args, int grid, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  wgrad_tc_kernel<BN><<<grid, WG_THREADS, smem, s>>>(a, b, args);
  return cudaPeekAtLastError();
}

}  // namespace pdae

using namespace pdae;

struct pdae_wgrad_tc_plan {
  CUtensorMap tmA, tmB;
  WgradArgs args;
  int BN, grid;
  size_t smem;
};

static int g_num_sms_w = 0;

extern "C" int pdae_wgrad_tc_supported(int H, int W, int Cin, int Cout, int ksize) {
  if (ksize != 1 && ksize != 3) return 0;
  if (Cin % 64 || Cout % 64) return 0;
  const int tw = pow2_tile_w(W, WG_KT), th = pow2_tile_w(H, WG_KT / tw);
  const int tn = WG_KT / (tw * th);
  return (W % tw == 0 && H % th == 0 && tw * th * tn == WG_KT && tn <= 64) ? 1 : 0;
}

extern "C" int pdae_wgrad_tc_create(pdae_wgrad_tc_plan** plan_out, const void* act3_bf16, const void* dy3_bf16, float* dw, int B,
                                    int H, int W, int Cin, int Cout, int ksize) {
  PDAE_REQUIRE(plan_out && act3_bf16 && dy3_bf16 && dw, "wgrad_tc_create: null pointer");
  PDAE_REQUIRE(pdae_wgrad_tc_supported(H, W, Cin, Cout, ksize), "wgrad_tc_create: unsupported shape H=%d W=%d Cin=%d Cout=%d k=%d", H, W,
               Cin, Cout, ksize);
  PDAE_REQUIRE(!(((uintptr_t)act3_bf16 | (uintptr_t)dy3_bf16) & 15) && !((uintptr_t)dw & 3),
               "wgrad_tc_create: act3 / dy3 must be 16-byte aligned (TMA), dw 4-byte aligned");
  EncodeTiledFnW enc = encode_fnw();
  PDAE_REQUIRE(enc != nullptr, "wgrad_tc_create: cuTensorMapEncodeTiled unavailable (no driver)");
  if (g_num_sms_w == 0) {
    int dev = 0;
    PDAE_CUDA(cudaGetDevice(&dev));
    PDAE_CUDA(cudaDeviceGetAttribute(&g_num_sms_w, cudaDevAttrMultiProcessorCount, dev));
  }
  pdae_wgrad_tc_plan* pl = new pdae_wgrad_tc_plan();
  WgradArgs& a = pl->args;
  a.dw = dw;
  a.a_is_act = (Cout % 128 == 0) ? 0 : 1;        // prefer dY on the M side: coalesced reductions into dW[tap][cin][cout]
  a.pair = (a.a_is_act && Cin % 128) ? 1 : 0;    // neither side fills 128 accumulator rows: two taps of 64 channels do
  a.Ma = a.a_is_act ? Cin : Cout;
  a.Nb = a.a_is_act ? Cout : Cin;
  a.sm = a.a_is_act ? Cout : 1;
  a.sn = a.a_is_act ? 1 : Cout;
  a.stap = (long long)Cin * Cout;
  const int BN = (a.Nb % 128 == 0) ? 128 : 64;
  pl->BN = BN;
  a.mchunks = a.pair ? a.Ma / 64 : a.Ma / 128; a.nchunks = a.Nb / BN; a.taps = ksize * ksize; a.ksize = ksize;
  a.tw = pow2_tile_w(W, WG_KT); a.th = pow2_tile_w(H, WG_KT / a.tw); a.tn = WG_KT / (a.tw * a.th);
  a.tiles_x = W / a.tw; a.tiles_y = H / a.th; a.tiles_b = (B + a.tn - 1) / a.tn;
  a.ktiles = a.tiles_x * a.tiles_y * a.tiles_b;
  a.B = B; a.H = H; a.W = W;
  a.items = (long long)(a.pair ? (a.taps + 1) / 2 : a.taps) * a.mchunks * a.nchunks * a.ktiles;
  const int stage = 2 * (2 * WG_BOX) + 2 * (BN / 64) * WG_BOX;
  int stages = (220 * 1024 - 1024) / stage;
  if (stages > WG_MAX_ST) stages = WG_MAX_ST;
  a.stages = stages;
  pl->smem = (size_t)stages * stage + 1024;
  pl->grid = a.items < g_num_sms_w ? (int)a.items : g_num_sms_w;
  const void* At = a.a_is_act ? act3_bf16 : dy3_bf16;
  const void* Bt = a.a_is_act ? dy3_bf16 : act3_bf16;
  const int Ca = 3 * a.Ma, Cb = 3 * a.Nb;
  cuuint32_t estr4[4] = {1, 1, 1, 1};
  cuuint32_t box[4] = {64, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
  for (int i = 0; i < 2; ++i) {
    const int C = i ? Cb : Ca;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    CUresult r = enc(i ? &pl->tmB : &pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(i ? Bt : At), dims, strides, box,
                     estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      set_error("wgrad_tc_create: cuTensorMapEncodeTiled failed with %d", (int)r);
      return PDAE_EINVAL;
    }
  }
  *plan_out = pl;
  return PDAE_OK;
}

extern "C" int pdae_wgrad_tc_run(const pdae_wgrad_tc_plan* pl, pdae_stream_t stream) {
  PDAE_REQUIRE(pl, "wgrad_tc_run: null plan");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = pl->BN == 128 ? launch_wg<128>(pl->tmA, pl->tmB, pl->args, pl->grid, pl->smem, s)
                                : launch_wg<64>(pl->tmA, pl->tmB, pl->args, pl->grid, pl->smem, s);
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("launch of wgrad_tc_kernel<%d> failed: %s", pl->BN, cudaGetErrorString(e));
    return PDAE_ECUDA;
  }
  return PDAE_OK;
}

extern "C" void pdae_wgrad_tc_destroy(pdae_wgrad_tc_plan* pl) { delete pl; }
