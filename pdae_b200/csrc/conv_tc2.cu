// conv_tc v2 -- persistent, warp-specialised tensor-core implicit-GEMM convolution for sm_100a.
//
//   warp 8   : TMA producer  (4-D activation boxes with OOB zero fill = conv padding, 3-D weight boxes)
//   warp 9   : tcgen05.mma issuer; accumulators are DOUBLE-BUFFERED in TMEM (2 x BN columns), so
//   warps 0-7: (two groups) the epilogue of tile i overlaps the main loop of tile i+1:
//   (the two single-thread roles have the HIGHEST warp ids: the scheduler favours higher ids among eligible warps, so the
//    issuer / producer are not starved of issue slots by the eight epilogue warps)
//              tcgen05.ld -> +bias (+fp32 residual fetched by TMA) -> 128B-swizzled staging tile in smem ->
//              per-channel GroupNorm partial sums (sum, sum^2) -> TMA store (fp32 or bf16 NHWC).
// One CTA per SM, static round-robin tile schedule (tile = blockIdx.x + i * gridDim.x; all CTAs sweep the same
// weight tile at the same time, so it stays hot in L2).
//
// The BN=16 instantiation is the UNet image head (Cout=3 zero-padded to 16): it skips the TMA store and writes the
// first `cout_valid` columns as NCHW fp32 planes (coalesced along x).
#include <cuda.h>

#include "common.cuh"

namespace pdae {

constexpr int T2_BM = 128;
constexpr int T2_BK = 64;
constexpr int T2_A_BYTES = T2_BM * T2_BK * 2;  // 16 KB
constexpr int T2_STG_BYTES = 128 * 128;        // staging tile: 128 rows x 128 B
constexpr int T2_MAX_STAGES = 8;
constexpr int T2_THREADS = 320;   // warp 0 TMA, warp 1 MMA, warps 2-5 / 6-9 = two epilogue groups (one per TMEM accumulator)
constexpr int T2_EPI_THREADS = 128;

struct ConvTc2Args {
  const float* bias;
  float* ch_stats;    // [B][Cout][2] fp32 partial (sum, sum^2) accumulators or nullptr
  float* out_nchw;    // BN==16 head: NCHW fp32 output
  const long long* fuse;  // BN==16 head: optional device-side descriptor of the fused DDIM update (see pdae_conv_tc2_set_head_fuse)
  int B, H, W, Cout;
  int tw, th, tn;
  int tiles_x, tiles_y, tiles_b, tiles_m, tiles_total;
  int taps, ksize, kblocks;
  int stages;
  int has_res, out_bf16, cout_valid;
  int w_batched;  // B operand is a per-image matrix (batched GEMM): 3rd TMA coordinate = image index
  int kblocks2;   // fused 1x1 skip conv: extra K blocks from a second (activation, weight) pair, accumulated into the same tile
  int kblocks2a;  // ... of which the first kblocks2a come from tmA2, the rest from tmA3 (virtual channel concat of two tensors)
  int w_stat;     // weights stationary: every B tile of the (single) n-tile is loaded ONCE per CTA into its own shared-memory
                  // region and reused by all of the CTA's tiles; pipeline stages then hold A tiles only
  float softmax_alpha;  // > 0: the epilogue stores softmax_row(alpha * acc) (bf16) instead of acc -- attention scores whose
                        // whole row lives in this tile's TMEM accumulator (N == BN); model/module.py:452-455,483-486
  int dbg_shift;  // probe (scripts/desc_shift_probe.py): load the A box dbg_shift pixels EARLY and start the UMMA descriptor
  int dbg_boff;   // dbg_shift rows (x 128 B) into it, with the descriptor's base_offset field = dbg_boff -- 0 in production
};

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
__device__ __forceinline__ uint32_t mb_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done;
}
// Slow path of a wait: mbarrier.try_wait with a SUSPEND-TIME HINT, so a waiting warp sleeps in hardware (it is woken by the
// completing arrive) instead of re-issuing the poll every ~100 cycles.  With 18+ warps per CTA of which most are waiting at
// any time, hot polling took more than half of all issue slots away from the warps that had work (ncu: 8.4 M polls per
// launch, profiles/r02_ncu_conv_tc3_spin.txt).
__device__ __forceinline__ uint32_t mb_try_sleep(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity), "r"(20000u)
      : "memory");
  return done;
}
__device__ __noinline__ void mb_wait_slow(uint32_t bar, uint32_t parity) {
  uint32_t n = 0;
  while (!mb_try_sleep(bar, parity))
    if (++n > 4000000u) __trap();  // a protocol bug must trap, never hang the GPU
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
  if (mb_try(bar, parity)) return;   // fast path: already complete
  mb_wait_slow(bar, parity);
}
__device__ __forceinline__ void tma_ld4(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_ld3(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(m), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_st4(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(m), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
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
// One elected lane of a fully converged warp: the producer / issuer loops run on all 32 lanes (warp-uniform operands stay in
// uniform registers); only the TMA / tcgen05 instruction is predicated.  Under `if (lane == 0)` every descriptor was a
// per-lane value and each MMA paid an ELECT + 5 x R2UR.BROADCAST + branch waterfall.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void epi_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory"); }

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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

// byte offset of (row, 16-byte chunk) inside a 128-row x 128-byte SWIZZLE_128B staging tile
__device__ __forceinline__ uint32_t swz(int row, int chunk16) { return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4)); }

template <int BN>
__global__ void __launch_bounds__(T2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmR,
                const __grid_constant__ CUtensorMap tmA2, const __grid_constant__ CUtensorMap tmB2,
                const __grid_constant__ CUtensorMap tmA3, ConvTc2Args p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[T2_MAX_STAGES], bar_empty[T2_MAX_STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2], bar_acc_empty[2], bar_res[2], bar_w;
  __shared__ uint32_t tmem_slot;
  __shared__ float st_acc[2][2][BN < 32 ? 32 : BN];  // [epilogue group][sum | sum^2][channel] of the group's current (image, n-tile)

  constexpr int B_BYTES = BN * T2_BK * 2;
  constexpr int STAGE_BYTES = T2_A_BYTES + ((B_BYTES + 1023) / 1024) * 1024;
  constexpr int TMEM_COLS = (2 * BN < 32) ? 32 : 2 * BN;
  const uint32_t smem0 = (s_u32(smem_raw) + 1023u) & ~1023u;
  const int S = p.stages;
  const uint32_t stage_stride = p.w_stat ? (uint32_t)T2_A_BYTES : (uint32_t)STAGE_BYTES;
  const uint32_t wbase = smem0 + (uint32_t)S * stage_stride;                       // stationary weights (w_stat only)
  const uint32_t stg_out = wbase + (p.w_stat ? (uint32_t)((p.taps * p.kblocks + p.kblocks2) * B_BYTES) : 0u);   // 2 x 16 KB output staging   (one per epilogue group)
  const uint32_t stg_res = stg_out + 2u * T2_STG_BYTES;         // 2 x 16 KB residual staging (one per group; only if has_res)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_k = p.taps * p.kblocks;          // main conv
  const int total_all = total_k + p.kblocks2;      // + fused 1x1 skip conv
  // contiguous tile range per CTA: consecutive tiles of a CTA mostly belong to the same image, so the per-channel
  // GroupNorm partial sums can be accumulated in shared memory across tiles and flushed once per image
  const int per_cta = (p.tiles_total + (int)gridDim.x - 1) / (int)gridDim.x;
  const int tile_begin = (int)blockIdx.x * per_cta;
  const int tile_end = min(p.tiles_total, tile_begin + per_cta);

  for (int j = threadIdx.x; j < (BN < 32 ? 32 : BN); j += T2_THREADS) {
    st_acc[0][0][j] = 0.f; st_acc[0][1][j] = 0.f;
    st_acc[1][0][j] = 0.f; st_acc[1][1][j] = 0.f;
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mb_init(s_u32(&bar_full[s]), 1);
      mb_init(s_u32(&bar_empty[s]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mb_init(s_u32(&bar_acc_full[i]), 1);
      mb_init(s_u32(&bar_acc_empty[i]), 1);
      mb_init(s_u32(&bar_res[i]), 1);
    }
    mb_init(s_u32(&bar_w), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
  }
  if (warp == 9) {
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;

  if (warp == 8) {
    // ================= TMA producer (warp-uniform loop, elected lane issues) =================
    int s = 0;
    uint32_t ph = 0;
    const bool ws = p.w_stat != 0;
    const uint32_t stage_tx = ws ? (uint32_t)T2_A_BYTES : (uint32_t)(T2_A_BYTES + B_BYTES);
    if (ws && tile_begin < tile_end) {   // all B tiles of the layer, once (single n-tile: n0 = 0)
      const uint32_t wb = s_u32(&bar_w);
      if (elect_one()) {
        mb_expect_tx(wb, (uint32_t)(total_all * B_BYTES));
        for (int it = 0; it < total_k; ++it)
          tma_ld3(wbase + (uint32_t)(it * B_BYTES), &tmB, wb, (it % p.kblocks) * T2_BK, 0, it / p.kblocks);
        for (int kb2 = 0; kb2 < p.kblocks2; ++kb2)
          tma_ld3(wbase + (uint32_t)((total_k + kb2) * B_BYTES), &tmB2, wb, kb2 * T2_BK, 0, 0);
      }
      __syncwarp();
    }
    for (int tile = tile_begin; tile < tile_end; ++tile) {
      const int nt = tile / p.tiles_m;
      int mt = tile - nt * p.tiles_m;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int bt = mt / p.tiles_y;
      const int x0 = tx * p.tw, y0 = ty * p.th, b0 = bt * p.tn, n0 = nt * BN;
      int tap = 0, kb = 0, dy = p.ksize == 3 ? -1 : 0, dx = dy;
      for (int it = 0; it < total_k; ++it) {
        mb_wait(s_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = s_u32(&bar_full[s]);
        const uint32_t sa = smem0 + (uint32_t)s * stage_stride;
        if (elect_one()) {
          mb_expect_tx(full, stage_tx);
          tma_ld4(sa, &tmA, full, kb * T2_BK, x0 + dx - p.dbg_shift, y0 + dy, b0);
          if (!ws) tma_ld3(sa + T2_A_BYTES, &tmB, full, kb * T2_BK, n0, p.w_batched ? b0 : tap);
        }
        __syncwarp();
        if (++kb == p.kblocks) {
          kb = 0;
          ++tap;
          if (p.ksize == 3 && ++dx == 2) { dx = -1; ++dy; }
        }
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      for (int kb2 = 0; kb2 < p.kblocks2; ++kb2) {   // fused 1x1 skip conv: un-normalised input, centre tap
        mb_wait(s_u32(&bar_empty[s]), ph ^ 1u);
        const uint32_t full = s_u32(&bar_full[s]);
        const uint32_t sa = smem0 + (uint32_t)s * stage_stride;
        if (elect_one()) {
          mb_expect_tx(full, stage_tx);
          if (kb2 < p.kblocks2a) tma_ld4(sa, &tmA2, full, kb2 * T2_BK, x0, y0, b0);
          else tma_ld4(sa, &tmA3, full, (kb2 - p.kblocks2a) * T2_BK, x0, y0, b0);
          if (!ws) tma_ld3(sa + T2_A_BYTES, &tmB2, full, kb2 * T2_BK, n0, 0);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 9) {
    // ================= MMA issuer (warp-uniform loop, elected lane issues) =================
    constexpr uint32_t IDESC =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(T2_BM >> 4) << 24);
    const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
    int s = 0, tl = 0;
    uint32_t ph = 0;
    if (p.w_stat && tile_begin < tile_end) mb_wait(s_u32(&bar_w), 0u);   // stationary weights have landed
    for (int tile = tile_begin; tile < tile_end; ++tile, ++tl) {
      const int ab = tl & 1;
      mb_wait(s_u32(&bar_acc_empty[ab]), (uint32_t)(((tl >> 1) & 1) ^ 1));  // epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tmem_d = tmem_u + (uint32_t)(ab * BN);
      for (int it = 0; it < total_all; ++it) {
        mb_wait(s_u32(&bar_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t sa = smem0 + (uint32_t)s * stage_stride;
        const uint64_t ad = sw128_desc(sa + (uint32_t)p.dbg_shift * 128u) | ((uint64_t)p.dbg_boff << 49),
                       bd = sw128_desc(p.w_stat ? wbase + (uint32_t)(it * B_BYTES) : sa + T2_A_BYTES);
        const uint32_t first = (uint32_t)(it != 0);
        const uint32_t bar_e = s_u32(&bar_empty[s]);
        if (elect_one()) {
          umma(tmem_d, ad, bd, IDESC, first);
#pragma unroll
          for (int k = 1; k < T2_BK / 16; ++k) umma(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), IDESC, 1u);
          umma_commit_to(bar_e);
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit_to(s_u32(&bar_acc_full[ab]));
      __syncwarp();
    }
  } else {
    // ================= epilogue: two groups of 128 threads; group g drains accumulator buffer g =================
    // (tile tl of this CTA lands in accumulator tl & 1, so the groups work on alternate tiles concurrently: the per-tile
    //  epilogue chain -- tcgen05.ld, residual, staging, TMA store, statistics -- has twice the throughput)
    const int eg = warp >> 2;                  // epilogue group 0 | 1
    const int et = threadIdx.x - eg * 128;     // 0..127 inside the group
    const bool elected = et == 0;
    const int q = warp & 3;                    // TMEM lane quadrant of this warp
    const int r = q * 32 + lane;               // accumulator row = pixel index in the tile
    const int ppi = p.th * p.tw;               // pixels per image inside a tile
    int rc = 0;                                // residual loads consumed by this group (mbarrier phase)
    const uint32_t obuf = stg_out + (uint32_t)eg * T2_STG_BYTES, rbuf = stg_res + (uint32_t)eg * T2_STG_BYTES;
    const uint32_t rbar = s_u32(&bar_res[eg]);
    for (int tile = tile_begin + eg, tl = eg; tile < tile_end; tile += 2, tl += 2) {
      const int nt = tile / p.tiles_m;
      int mt = tile - nt * p.tiles_m;
      const int tx = mt % p.tiles_x;
      mt /= p.tiles_x;
      const int ty = mt % p.tiles_y;
      const int bt = mt / p.tiles_y;
      const int x0 = tx * p.tw, y0 = ty * p.th, b0 = bt * p.tn, n0 = nt * BN;
      const int ab = tl & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(ab * BN) + ((uint32_t)(q * 32) << 16);
      mb_wait(s_u32(&bar_acc_full[ab]), (uint32_t)((tl >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

      if constexpr (BN == 16) {
        // ---- image head: first cout_valid columns -> NCHW fp32 planes ----
        uint32_t v[16];
        tmem_ld16(tmem_acc, v);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int ni = r / ppi, rem = r - ni * ppi;
        const int yy = rem / p.tw, xx = rem - yy * p.tw;
        const int b = b0 + ni;
        if (b < p.B) {
          const long long hw = (long long)p.H * p.W;
          const long long pix = (long long)(y0 + yy) * p.W + (x0 + xx);
          float* o = p.out_nchw + (long long)b * p.cout_valid * hw + pix;
          float val[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            val[j] = __uint_as_float(v[j]) + (p.bias ? __ldg(p.bias + j) : 0.f);
            if (j < p.cout_valid) o[j * hw] = val[j];
          }
          // ---- fused DDIM update (diffusion/ddim.py:43-55, 66-79, 91-107, 123-138): this head's output is the last tensor the
          // step needs, so x_t -> x_{t-1} (or x_{t+1}) is finished here, in place, with the exact arithmetic of ddim_step_kernel.
          // The descriptor lives in device memory and is (re)written by the sampling loop; flags == 0 -> plain head conv.
          if (p.fuse) {
            const long long flags = __ldg(p.fuse);
            if (flags & 1) {
              const int C = (int)((flags >> 8) & 0xff), Ce = (int)((flags >> 16) & 0xff);
              const float* eps = reinterpret_cast<const float*>(__ldg(p.fuse + 1));
              float* xt = reinterpret_cast<float*>(__ldg(p.fuse + 2));
              const long long tb = reinterpret_cast<const long long*>(__ldg(p.fuse + 3))[b];
              const float A = reinterpret_cast<const float*>(__ldg(p.fuse + 4))[tb];
              const float Bm = reinterpret_cast<const float*>(__ldg(p.fuse + 5))[tb];
              const float abar = reinterpret_cast<const float*>(__ldg(p.fuse + 7))[tb];
              const float s1m = (flags & 2) ? reinterpret_cast<const float*>(__ldg(p.fuse + 6))[tb] : 0.f;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (j < C) {
                  const long long xi = ((long long)b * C + j) * hw + pix;
                  float e = val[j];                                  // this head predicts epsilon itself
                  if (flags & 2) e = __fsub_rn(eps[((long long)b * Ce + j) * hw + pix], __fmul_rn(s1m, val[j]));   // own output = shift term
                  else if (flags & 4) e = eps[((long long)b * Ce + j) * hw + pix];   // shift unused this step: epsilon from the other head
                  const float ax = __fmul_rn(A, xt[xi]);
                  float x0 = __fsub_rn(ax, __fmul_rn(Bm, e));
                  x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
                  const float e2 = __fdiv_rn(__fsub_rn(ax, x0), Bm);
                  xt[xi] = __fadd_rn(__fmul_rn(x0, sqrtf(abar)), __fmul_rn(sqrtf(__fsub_rn(1.0f, abar)), e2));
                }
              }
            }
          }
        }
      } else {
        const int CW = p.out_bf16 ? 64 : 32;  // accumulator columns per staging tile (128-byte rows)
        const int nch = BN / CW;
        if (p.has_res && elected) {            // residual chunk 0 of this tile (issued before the accumulator is needed)
          mb_expect_tx(rbar, T2_STG_BYTES);
          tma_ld4(rbuf, &tmR, rbar, n0, x0, y0, b0);
        }
        // softmax epilogue: three passes over the accumulator row held in TMEM (max, sum of exp, normalised store)
        float sm_mx = 0.f, sm_inv = 1.f;
        const float sm_a = p.softmax_alpha * 1.4426950408889634f;
        if (p.softmax_alpha > 0.f) {
          float mx = -3.0e38f;
          for (int c2 = 0; c2 < BN / 32; ++c2) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)(c2 * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(v[j]));
          }
          float sum = 0.f;
          for (int c2 = 0; c2 < BN / 32; ++c2) {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)(c2 * 32), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += exp2f((__uint_as_float(v[j]) - mx) * sm_a);
          }
          sm_mx = mx;
          sm_inv = 1.0f / sum;
        }
        for (int c = 0; c < nch; ++c) {
          float val[64];
          {
            uint32_t v[32];
            tmem_ld32(tmem_acc + (uint32_t)(c * CW), v);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) val[j] = __uint_as_float(v[j]);
            if (p.out_bf16) {
              tmem_ld32(tmem_acc + (uint32_t)(c * CW + 32), v);
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 32; ++j) val[32 + j] = __uint_as_float(v[j]);
            }
          }
          if (p.softmax_alpha > 0.f) {
#pragma unroll
            for (int j = 0; j < 64; ++j) val[j] = exp2f((val[j] - sm_mx) * sm_a) * sm_inv;
          }
          if (p.bias) {
            const float4* bp = reinterpret_cast<const float4*>(p.bias + n0 + c * CW);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j * 4 < CW) {
                const float4 bv = __ldg(bp + j);
                val[4 * j + 0] += bv.x; val[4 * j + 1] += bv.y; val[4 * j + 2] += bv.z; val[4 * j + 3] += bv.w;
              }
            }
          }
          if (p.has_res) {                      // the residual has the output's dtype: CW columns = one 128-byte row
            mb_wait(rbar, (uint32_t)(rc & 1));
            if (p.out_bf16) {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint32_t w[4];
                asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3])
                             : "r"(rbuf + swz(r, j)));
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                  val[8 * j + 2 * h] += __uint_as_float(w[h] << 16);
                  val[8 * j + 2 * h + 1] += __uint_as_float(w[h] & 0xffff0000u);
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                float4 rv;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(rv.x), "=f"(rv.y), "=f"(rv.z), "=f"(rv.w)
                             : "r"(rbuf + swz(r, j)));
                val[4 * j + 0] += rv.x; val[4 * j + 1] += rv.y; val[4 * j + 2] += rv.z; val[4 * j + 3] += rv.w;
              }
            }
            ++rc;
          }
          // the group's staging buffer must no longer be read by its previous TMA store
          if (elected) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          epi_bar(eg);                          // (also: every thread has consumed the residual buffer)
          if (p.has_res && elected && c + 1 < nch) {   // residual for the next chunk overlaps this chunk's store + statistics
            mb_expect_tx(rbar, T2_STG_BYTES);
            tma_ld4(rbuf, &tmR, rbar, n0 + (c + 1) * CW, x0, y0, b0);
          }
          if (p.out_bf16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                __nv_bfloat162 b2 = __floats2bfloat162_rn(val[8 * j + 2 * h], val[8 * j + 2 * h + 1]);
                w[h] = *reinterpret_cast<uint32_t*>(&b2);
              }
              asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(obuf + swz(r, j)), "r"(w[0]), "r"(w[1]), "r"(w[2]),
                           "r"(w[3])
                           : "memory");
            }
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
              asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(obuf + swz(r, j)), "f"(val[4 * j]), "f"(val[4 * j + 1]),
                           "f"(val[4 * j + 2]), "f"(val[4 * j + 3])
                           : "memory");
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          epi_bar(eg);
          if (elected) {
            tma_st4(&tmO, obuf, n0 + c * CW, x0, y0, b0);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          if (p.ch_stats) {
            // per-channel partial sums over this tile's rows, read back from the staged (rounded) values:
            // thread -> column (et % CW), rows [(et / CW) * CW, +CW)
            const int col = et % CW, r0 = (et / CW) * CW;
            const uint32_t cbyte = p.out_bf16 ? (uint32_t)((col & 7) * 2) : (uint32_t)((col & 3) * 4);
            const int cchunk = p.out_bf16 ? (col >> 3) : (col >> 2);
            auto ldv = [&](int rr) -> float {
              float x;
              if (p.out_bf16) {
                unsigned short h;
                asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(obuf + swz(rr, cchunk) + cbyte));
                x = __uint_as_float(((uint32_t)h) << 16);
              } else {
                asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(obuf + swz(rr, cchunk) + cbyte));
              }
              return x;
            };
            if (p.tn == 1) {
              // whole tile = one image: accumulate in shared memory across this CTA's consecutive tiles
              float s = 0.f, qq = 0.f;
#pragma unroll 8
              for (int rr = r0; rr < r0 + CW; ++rr) {
                const float x = ldv(rr);
                s += x;
                qq = fmaf(x, x, qq);
              }
              atomicAdd(&st_acc[eg][0][c * CW + col], s);
              atomicAdd(&st_acc[eg][1][c * CW + col], qq);
            } else if (ppi % CW == 0) {
              // several whole images per tile and this thread's CW rows lie inside ONE image
              float s = 0.f, qq = 0.f;
#pragma unroll 8
              for (int rr = r0; rr < r0 + CW; ++rr) {
                const float x = ldv(rr);
                s += x;
                qq = fmaf(x, x, qq);
              }
              const int cur = r0 / ppi;
              if (b0 + cur < p.B) {
                float* dst = p.ch_stats + ((long long)(b0 + cur) * p.Cout + n0 + c * CW + col) * 2;
                if (ppi == CW && p.tiles_x * p.tiles_y == 1) {   // whole image inside this tile and this thread covers all of
                                                                 // its rows: the only contributor -> plain store, no atomic
                  *reinterpret_cast<float2*>(dst) = make_float2(s, qq);
                } else {
                  atomicAdd(dst, s);
                  atomicAdd(dst + 1, qq);
                }
              }
            } else {
              float s = 0.f, qq = 0.f;
              int cur = r0 / ppi, nxt = (cur + 1) * ppi;  // `nxt` = first row of the next image
              float* dst0 = p.ch_stats + ((long long)b0 * p.Cout + n0 + c * CW + col) * 2;
              for (int rr = r0; rr < r0 + CW; ++rr) {
                if (rr == nxt) {
                  if (b0 + cur < p.B) {
                    atomicAdd(dst0 + (long long)cur * p.Cout * 2, s);
                    atomicAdd(dst0 + (long long)cur * p.Cout * 2 + 1, qq);
                  }
                  s = qq = 0.f;
                  ++cur;
                  nxt += ppi;
                }
                const float x = ldv(rr);
                s += x;
                qq = fmaf(x, x, qq);
              }
              if (b0 + cur < p.B) {
                atomicAdd(dst0 + (long long)cur * p.Cout * 2, s);
                atomicAdd(dst0 + (long long)cur * p.Cout * 2 + 1, qq);
              }
            }
          }
        }
      }
      // accumulator buffer fully read by every epilogue thread -> hand it back to the MMA warp
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      epi_bar(eg);
      if (elected) mb_arrive(s_u32(&bar_acc_empty[ab]));
      if constexpr (BN != 16) {
        if (p.ch_stats && p.tn == 1) {
          bool flush = tile + 2 >= tile_end;     // this group's next tile is tile + 2
          if (!flush) {
            const int nt2 = (tile + 2) / p.tiles_m;
            const int bt2 = ((tile + 2) - nt2 * p.tiles_m) / (p.tiles_x * p.tiles_y);
            flush = nt2 != nt || bt2 != bt;
          }
          if (flush) {  // (the epi_bar above ordered every thread's shared-memory atomics before these reads)
            for (int j = et; j < BN; j += 128) {
              float* dst = p.ch_stats + ((long long)b0 * p.Cout + n0 + j) * 2;
              atomicAdd(dst, st_acc[eg][0][j]);
              atomicAdd(dst + 1, st_acc[eg][1][j]);
              st_acc[eg][0][j] = 0.f;
              st_acc[eg][1][j] = 0.f;
            }
            epi_bar(eg);
          }
        }
      }
    }
    if (elected) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 9) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn2 encode_fn2() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn) return fn;
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess) {
    (void)cudaGetLastError();
    return nullptr;
  }
  fn = (EncodeTiledFn2)sym;
  return fn;
}

static int pow2_tile(int W, int cap) {
  int t = 1;
  while (t * 2 <= cap && W % (t * 2) == 0) t *= 2;
  return t;
}

template <int BN>
static cudaError_t launch_tc2(const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& o, const CUtensorMap& r,
                              const CUtensorMap& a2, const CUtensorMap& b2, const CUtensorMap& a3, const ConvTc2Args& args,
                              int grid, size_t smem, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 221 * 1024);  // + static (barriers, 2 x per-group stats <= 4 KB) <= 227 KB
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  conv_tc2_kernel<BN><<<grid, T2_THREADS, smem, s>>>(a, b, o, r, a2, b2, a3, args);
  return cudaPeekAtLastError();
}

}  // namespace pdae

using namespace pdae;

struct pdae_conv_tc2_plan {
  CUtensorMap tmA, tmB, tmO, tmR, tmA2, tmB2, tmA3;
  ConvTc2Args args;
  int BN, grid;
  size_t smem;
};

static int g_num_sms = 0;

struct Tc2Desc {
  const void* in; const void* w; const float* bias; const void* residual; void* out;
  int out_dtype; float* ch_stats;
  int B, H, W, Cin, Cout, ksize, cout_valid, bn_override;
  long long in_ld;            // elements between consecutive pixels of the A operand
  long long in_bs;            // elements between consecutive images of the A operand
  int w_batched;              // 0: weights [taps][Cout][Cin] ; 1: per-image matrix [B][Cout rows][Cin], strides below
  long long w_ld, w_bs;
  long long out_ld, out_bs;   // elements between consecutive pixels / images of the output
  const void* in2 = nullptr; const void* w2 = nullptr; int Cin2 = 0;   // fused 1x1 skip conv (bf16 NHWC input, [Cout][Cin2] weights)
  float softmax_alpha = 0.f;  // batched GEMM only: store softmax_row(alpha * out) as bf16 (needs N == BN)
  const void* in3 = nullptr; int Cin2a = 0;   // skip input = channel concat of in2 [..,Cin2a] and in3 [..,Cin2-Cin2a] (in3 == nullptr: in2 alone)
};

static int tc2_create(pdae_conv_tc2_plan** plan_out, const Tc2Desc& d) {
  const int B = d.B, H = d.H, W = d.W, Cin = d.Cin, Cout = d.Cout, ksize = d.ksize, cout_valid = d.cout_valid;
  PDAE_REQUIRE(plan_out && d.in && d.w && d.out, "conv_tc2_create: null pointer");
  PDAE_REQUIRE(ksize == 1 || ksize == 3, "conv_tc2_create: ksize must be 1 or 3");
  PDAE_REQUIRE(Cin % T2_BK == 0, "conv_tc2_create: Cin=%d not a multiple of 64", Cin);
  const bool head = cout_valid > 0;
  PDAE_REQUIRE(head ? (Cout == 16 && cout_valid <= 16 && d.out_dtype == PDAE_F32 && !d.residual && !d.ch_stats) : (Cout % 64 == 0),
               "conv_tc2_create: unsupported Cout=%d (cout_valid=%d)", Cout, cout_valid);
  PDAE_REQUIRE(d.out_dtype == PDAE_F32 || d.out_dtype == PDAE_BF16, "conv_tc2_create: bad out dtype");
  // (a residual is read in the output's dtype: fp32 with an fp32 output, bf16 with a bf16 output)
  PDAE_REQUIRE(((uintptr_t)d.in & 15) == 0 && ((uintptr_t)d.w & 15) == 0 && ((uintptr_t)d.out & 15) == 0 &&
                   ((uintptr_t)d.residual & 15) == 0 && ((uintptr_t)d.bias & 15) == 0,
               "conv_tc2_create: pointers must be 16-byte aligned");
  PDAE_REQUIRE(d.in_ld % 8 == 0 && d.in_bs % 8 == 0 && d.w_ld % 8 == 0 && d.w_bs % 8 == 0 && d.out_ld % 8 == 0 &&
                   d.out_bs % 8 == 0, "conv_tc2_create: strides must be multiples of 16 bytes");
  EncodeTiledFn2 enc = encode_fn2();
  PDAE_REQUIRE(enc != nullptr, "conv_tc2_create: cuTensorMapEncodeTiled unavailable (no driver)");
  if (g_num_sms == 0) {
    int dev = 0;
    PDAE_CUDA(cudaGetDevice(&dev));
    PDAE_CUDA(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  }
  pdae_conv_tc2_plan* pl = new pdae_conv_tc2_plan();
  ConvTc2Args& a = pl->args;
  a.bias = d.bias; a.ch_stats = d.ch_stats; a.out_nchw = head ? (float*)d.out : nullptr;
  a.fuse = nullptr;
  a.B = B; a.H = H; a.W = W; a.Cout = Cout;
  a.tw = pow2_tile(W, T2_BM);
  a.th = pow2_tile(H, T2_BM / a.tw);
  a.tn = T2_BM / (a.tw * a.th);
  if (W % a.tw != 0 || H % a.th != 0 || a.tw * a.th * a.tn != T2_BM || a.tn > 256 || (d.w_batched && a.tn != 1)) {
    delete pl;
    PDAE_REQUIRE(false, "conv_tc2_create: H=%d W=%d cannot be tiled into 128-pixel boxes%s", H, W,
                 d.w_batched ? " of a single image (batched GEMM)" : "");
  }
  a.tiles_x = W / a.tw; a.tiles_y = H / a.th; a.tiles_b = (B + a.tn - 1) / a.tn;
  a.tiles_m = a.tiles_x * a.tiles_y * a.tiles_b;
  a.taps = ksize * ksize; a.ksize = ksize; a.kblocks = Cin / T2_BK;
  a.has_res = d.residual != nullptr; a.out_bf16 = d.out_dtype == PDAE_BF16; a.cout_valid = cout_valid;
  a.w_batched = d.w_batched;
  a.kblocks2 = d.Cin2 / T2_BK;
  a.softmax_alpha = d.softmax_alpha;
  if (d.softmax_alpha > 0.f && !(d.w_batched && d.out_dtype == PDAE_BF16 && !d.residual && !d.bias && !d.ch_stats &&
                                 d.bn_override == Cout)) {
    delete pl;
    PDAE_REQUIRE(false, "conv_tc2_create: the softmax epilogue needs a bf16 batched GEMM whose row fits one tile (N=%d)", Cout);
  }
  {   // descriptor row-shift probe knobs (never set in production)
    const char* e1 = getenv("PDAE_TC_DBG_SHIFT");
    const char* e2 = getenv("PDAE_TC_DBG_BOFF");
    a.dbg_shift = e1 ? atoi(e1) : 0;
    a.dbg_boff = e2 ? atoi(e2) : 0;
  }
  int BN;
  if (head) BN = 16;
  else if (d.bn_override == 64 || d.bn_override == 128 || d.bn_override == 256) BN = d.bn_override;
  else if (Cout % 256 == 0 && (long long)a.tiles_m * (Cout / 256) >= g_num_sms) BN = 256;  // fewer A re-reads per FLOP
  else BN = (Cout % 128 == 0) ? 128 : 64;
  if (!head && Cout % BN != 0) {
    delete pl;
    PDAE_REQUIRE(false, "conv_tc2_create: Cout=%d not a multiple of BN=%d", Cout, BN);
  }
  pl->BN = BN;
  a.tiles_total = a.tiles_m * (head ? 1 : Cout / BN);
  const int b_bytes = ((BN * T2_BK * 2 + 1023) / 1024) * 1024;
  int stage_bytes = T2_A_BYTES + b_bytes;
  const int staging = head ? 0 : (a.has_res ? 4 : 2) * T2_STG_BYTES;
  const int total_all = a.taps * a.kblocks + a.kblocks2;
  pl->grid = a.tiles_total < g_num_sms ? a.tiles_total : g_num_sms;
  // Weights stationary in shared memory: when the whole layer's B operand fits beside >= 4 A-only stages and every CTA
  // runs several tiles of the single n-tile, the weights are fetched once per CTA instead of once per tile (the narrow
  // 64-channel layers are L2->SM bandwidth-bound: this removes a third of their bytes).  PDAE_TC_WSTAT=0 disables (A/B aid).
  static int wstat_env = -1;
  if (wstat_env < 0) { const char* e = getenv("PDAE_TC_WSTAT"); wstat_env = (e && e[0] == '0') ? 0 : 1; }
  int wbytes = 0;
  a.w_stat = 0;
  if (wstat_env && !head && !d.w_batched && Cout == BN && b_bytes == BN * T2_BK * 2 && a.tiles_total >= 2 * pl->grid) {
    const int wb = total_all * b_bytes;
    if ((220 * 1024 - 1024 - staging - wb) / T2_A_BYTES >= 4) {
      a.w_stat = 1;
      wbytes = wb;
      stage_bytes = T2_A_BYTES;
    }
  }
  int stages = (220 * 1024 - 1024 - staging - wbytes) / stage_bytes;
  if (stages > T2_MAX_STAGES) stages = T2_MAX_STAGES;
  if (stages > total_all) stages = total_all;
  if (stages < 2) stages = 2;
  a.stages = stages;
  pl->smem = (size_t)stages * stage_bytes + wbytes + staging + 1024;

  auto fail = [&](const char* what, int code) {
    delete pl;
    set_error("conv_tc2_create: cuTensorMapEncodeTiled(%s) failed with %d", what, code);
    return PDAE_EINVAL;
  };
  cuuint32_t estr4[4] = {1, 1, 1, 1};
  {
    cuuint64_t dims[4] = {(cuuint64_t)Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)d.in_ld * 2, (cuuint64_t)W * d.in_ld * 2, (cuuint64_t)d.in_bs * 2};
    cuuint32_t box[4] = {(cuuint32_t)T2_BK, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
    CUresult r = enc(&pl->tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d.in), dims, strides, box, estr4,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("A", (int)r);
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)Cin, (cuuint64_t)Cout, (cuuint64_t)(d.w_batched ? B : a.taps)};
    cuuint64_t strides[2] = {(cuuint64_t)d.w_ld * 2, (cuuint64_t)d.w_bs * 2};
    cuuint32_t box[3] = {(cuuint32_t)T2_BK, (cuuint32_t)BN, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&pl->tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d.w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("W", (int)r);
  }
  pl->tmO = pl->tmA;
  pl->tmR = pl->tmA;
  pl->tmA2 = pl->tmA;
  pl->tmB2 = pl->tmB;
  pl->tmA3 = pl->tmA;
  if (d.Cin2 > 0) {
    const int Ca = d.in3 ? d.Cin2a : d.Cin2, Cb = d.Cin2 - Ca;
    if (!d.in2 || !d.w2 || d.Cin2 % T2_BK != 0 || head || d.w_batched || ((uintptr_t)d.in2 & 15) || ((uintptr_t)d.w2 & 15) ||
        (d.in3 && (Ca <= 0 || Cb <= 0 || Ca % T2_BK != 0 || ((uintptr_t)d.in3 & 15)))) {
      delete pl;
      PDAE_REQUIRE(false, "conv_tc2_create: bad fused-skip operands (Cin2=%d, Cin2a=%d)", d.Cin2, d.Cin2a);
    }
    a.kblocks2a = Ca / T2_BK;
    cuuint32_t box[4] = {(cuuint32_t)T2_BK, (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
    CUresult r = CUDA_SUCCESS;
    for (int part = 0; part < (d.in3 ? 2 : 1); ++part) {
      const int Cp = part ? Cb : Ca;
      cuuint64_t dims[4] = {(cuuint64_t)Cp, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
      cuuint64_t strides[3] = {(cuuint64_t)Cp * 2, (cuuint64_t)W * Cp * 2, (cuuint64_t)H * W * Cp * 2};
      r = enc(part ? &pl->tmA3 : &pl->tmA2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(part ? d.in3 : d.in2), dims,
              strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(part ? "A3" : "A2", (int)r);
    }
    cuuint64_t wd[3] = {(cuuint64_t)d.Cin2, (cuuint64_t)Cout, 1};
    cuuint64_t ws[2] = {(cuuint64_t)d.Cin2 * 2, (cuuint64_t)Cout * d.Cin2 * 2};
    cuuint32_t wb[3] = {(cuuint32_t)T2_BK, (cuuint32_t)BN, 1};
    cuuint32_t we[3] = {1, 1, 1};
    r = enc(&pl->tmB2, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(d.w2), wd, ws, wb, we, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("W2", (int)r);
  }
  if (!head) {
    const int esz = a.out_bf16 ? 2 : 4;
    cuuint64_t dims[4] = {(cuuint64_t)Cout, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)d.out_ld * esz, (cuuint64_t)W * d.out_ld * esz, (cuuint64_t)d.out_bs * esz};
    cuuint32_t box[4] = {(cuuint32_t)(a.out_bf16 ? 64 : 32), (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
    CUresult r = enc(&pl->tmO, a.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d.out, dims,
                     strides, box, estr4, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("O", (int)r);
    if (a.has_res) {
      cuuint64_t rstr[3] = {(cuuint64_t)Cout * esz, (cuuint64_t)W * Cout * esz, (cuuint64_t)H * W * Cout * esz};
      cuuint32_t rbox[4] = {(cuuint32_t)(a.out_bf16 ? 64 : 32), (cuuint32_t)a.tw, (cuuint32_t)a.th, (cuuint32_t)a.tn};
      r = enc(&pl->tmR, a.out_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4,
              const_cast<void*>((const void*)d.residual), dims, rstr, rbox, estr4,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail("R", (int)r);
    }
  }
  *plan_out = pl;
  return PDAE_OK;
}

extern "C" int pdae_conv_tc2_create(pdae_conv_tc2_plan** plan_out, const void* in_bf16, const void* w_bf16, const float* bias,
                                    const void* residual, void* out, int out_dtype, float* ch_stats, int B, int H, int W,
                                    int Cin, int Cout, int ksize, int cout_valid, int bn_override) {
  Tc2Desc d;
  d.in = in_bf16; d.w = w_bf16; d.bias = bias; d.residual = residual; d.out = out; d.out_dtype = out_dtype;
  d.ch_stats = ch_stats; d.B = B; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.ksize = ksize; d.cout_valid = cout_valid;
  d.bn_override = bn_override;
  d.in_ld = Cin; d.in_bs = (long long)H * W * Cin;
  d.w_batched = 0; d.w_ld = Cin; d.w_bs = (long long)Cout * Cin;
  d.out_ld = Cout; d.out_bs = (long long)H * W * Cout;
  return tc2_create(plan_out, d);
}

// conv (ksize 1|3) + fused 1x1 skip conv:  out = conv(in, w) + in2 * w2^T + bias (+ residual); bias must already hold b + b_skip
extern "C" int pdae_conv_tc2_create_skip(pdae_conv_tc2_plan** plan_out, const void* in_bf16, const void* w_bf16, const float* bias,
                                         const void* in2_bf16, const void* w2_bf16, int Cin2, void* out, int out_dtype,
                                         float* ch_stats, int B, int H, int W, int Cin, int Cout, int ksize, int bn_override) {
  Tc2Desc d;
  d.in = in_bf16; d.w = w_bf16; d.bias = bias; d.residual = nullptr; d.out = out; d.out_dtype = out_dtype;
  d.ch_stats = ch_stats; d.B = B; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.ksize = ksize; d.cout_valid = 0;
  d.bn_override = bn_override;
  d.in_ld = Cin; d.in_bs = (long long)H * W * Cin;
  d.w_batched = 0; d.w_ld = Cin; d.w_bs = (long long)Cout * Cin;
  d.out_ld = Cout; d.out_bs = (long long)H * W * Cout;
  d.in2 = in2_bf16; d.w2 = w2_bf16; d.Cin2 = Cin2;
  return tc2_create(plan_out, d);
}

// Same, the skip conv's input being the channel concat cat([in2a (Cin2a ch), in2b (Cin2b ch)]) of two NHWC bf16 tensors
// that is never materialised (unet.py:199 `torch.cat([h, hs.pop()], dim=1)` feeding module.py:297 skip_connection).
extern "C" int pdae_conv_tc2_create_skip2(pdae_conv_tc2_plan** plan_out, const void* in_bf16, const void* w_bf16,
                                          const float* bias, const void* in2a_bf16, int Cin2a, const void* in2b_bf16, int Cin2b,
                                          const void* w2_bf16, void* out, int out_dtype, float* ch_stats, int B, int H, int W,
                                          int Cin, int Cout, int ksize, int bn_override) {
  PDAE_REQUIRE(in2b_bf16 && Cin2b > 0, "conv_tc2_create_skip2: second skip source missing");
  Tc2Desc d;
  d.in = in_bf16; d.w = w_bf16; d.bias = bias; d.residual = nullptr; d.out = out; d.out_dtype = out_dtype;
  d.ch_stats = ch_stats; d.B = B; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.ksize = ksize; d.cout_valid = 0;
  d.bn_override = bn_override;
  d.in_ld = Cin; d.in_bs = (long long)H * W * Cin;
  d.w_batched = 0; d.w_ld = Cin; d.w_bs = (long long)Cout * Cin;
  d.out_ld = Cout; d.out_bs = (long long)H * W * Cout;
  d.in2 = in2a_bf16; d.w2 = w2_bf16; d.Cin2 = Cin2a + Cin2b; d.in3 = in2b_bf16; d.Cin2a = Cin2a;
  return tc2_create(plan_out, d);
}

// Batched GEMM on the same kernel: for every batch item i,  out_i[M x N] = A_i[M x K] * Bm_i[N x K]^T  (both K-major bf16).
// a_ld / b_ld / out_ld: elements between consecutive rows; *_bs: elements between consecutive batch items.
extern "C" int pdae_gemm_tc2_create(pdae_conv_tc2_plan** plan_out, const void* a_bf16, long long a_ld, long long a_bs,
                                    const void* b_bf16, long long b_ld, long long b_bs, void* out, int out_dtype,
                                    long long out_ld, long long out_bs, int batch, int M, int N, int K) {
  Tc2Desc d;
  d.in = a_bf16; d.w = b_bf16; d.bias = nullptr; d.residual = nullptr; d.out = out; d.out_dtype = out_dtype;
  d.ch_stats = nullptr; d.B = batch; d.H = 1; d.W = M; d.Cin = K; d.Cout = N; d.ksize = 1; d.cout_valid = 0; d.bn_override = 0;
  d.in_ld = a_ld; d.in_bs = a_bs;
  d.w_batched = 1; d.w_ld = b_ld; d.w_bs = b_bs;
  d.out_ld = out_ld; d.out_bs = out_bs;
  return tc2_create(plan_out, d);
}

// P_i = softmax_rows(alpha * A_i * Bm_i^T) stored as bf16: the attention-probability GEMM with the softmax folded into the
// epilogue (the fp32 score matrix never leaves TMEM).  N must be 64, 128 or 256 (one n-tile holds the whole row).
extern "C" int pdae_gemm_tc2_softmax_create(pdae_conv_tc2_plan** plan_out, const void* a_bf16, long long a_ld, long long a_bs,
                                            const void* b_bf16, long long b_ld, long long b_bs, void* out_bf16, long long out_ld,
                                            long long out_bs, int batch, int M, int N, int K, float alpha) {
  PDAE_REQUIRE(N == 64 || N == 128 || N == 256, "gemm_tc2_softmax_create: N=%d must be 64, 128 or 256", N);
  PDAE_REQUIRE(alpha > 0.f, "gemm_tc2_softmax_create: alpha must be positive");
  Tc2Desc d;
  d.in = a_bf16; d.w = b_bf16; d.bias = nullptr; d.residual = nullptr; d.out = out_bf16; d.out_dtype = PDAE_BF16;
  d.ch_stats = nullptr; d.B = batch; d.H = 1; d.W = M; d.Cin = K; d.Cout = N; d.ksize = 1; d.cout_valid = 0; d.bn_override = N;
  d.in_ld = a_ld; d.in_bs = a_bs;
  d.w_batched = 1; d.w_ld = b_ld; d.w_bs = b_bs;
  d.out_ld = out_ld; d.out_bs = out_bs;
  d.softmax_alpha = alpha;
  return tc2_create(plan_out, d);
}

extern "C" int pdae_conv_tc2_run(const pdae_conv_tc2_plan* pl, pdae_stream_t stream) {
  PDAE_REQUIRE(pl, "conv_tc2_run: null plan");
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e;
  switch (pl->BN) {
    case 16: e = launch_tc2<16>(pl->tmA, pl->tmB, pl->tmO, pl->tmR, pl->tmA2, pl->tmB2, pl->tmA3, pl->args, pl->grid, pl->smem, s); break;
    case 64: e = launch_tc2<64>(pl->tmA, pl->tmB, pl->tmO, pl->tmR, pl->tmA2, pl->tmB2, pl->tmA3, pl->args, pl->grid, pl->smem, s); break;
    case 128: e = launch_tc2<128>(pl->tmA, pl->tmB, pl->tmO, pl->tmR, pl->tmA2, pl->tmB2, pl->tmA3, pl->args, pl->grid, pl->smem, s); break;
    default: e = launch_tc2<256>(pl->tmA, pl->tmB, pl->tmO, pl->tmR, pl->tmA2, pl->tmB2, pl->tmA3, pl->args, pl->grid, pl->smem, s); break;
  }
  if (e != cudaSuccess) {
    (void)cudaGetLastError();
    set_error("launch of conv_tc2_kernel<%d> failed: %s", pl->BN, cudaGetErrorString(e));
    return PDAE_ECUDA;
  }
  return PDAE_OK;
}

// Image-head plans only: attach the device-side descriptor of the fused DDIM update, 8 x int64 =
// { flags, eps*, x_t*, t*, sqrt_recip_alphas_cumprod*, sqrt_recip_alphas_cumprod_m1*, sqrt_one_minus_alphas_cumprod*, alphas_cumprod_prev|next* }
// with flags = enabled | use_grad << 1 | eps_only << 2 | C << 8 | C_eps << 16.  The head keeps writing its own output; when enabled it also
// updates x_t in place (use_grad: this head produces the shift/gradient term and `eps` comes from the other head).
extern "C" int pdae_conv_tc2_set_head_fuse(pdae_conv_tc2_plan* pl, const int64_t* fuse_desc_device) {
  PDAE_REQUIRE(pl && pl->BN == 16, "conv_tc2_set_head_fuse: not an image-head plan");
  pl->args.fuse = reinterpret_cast<const long long*>(fuse_desc_device);
  return PDAE_OK;
}

extern "C" void pdae_conv_tc2_destroy(pdae_conv_tc2_plan* pl) { delete pl; }
