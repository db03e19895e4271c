// tcgen05 version of the hidden GCN layer:  Y (+)= relu( Ablk . (X . W) + b ),  X [M,200] fp32, W [200,200].
//
//   * one CTA = 128 rows (8 environments of 16 nodes / 4 of 32) x all 208 (padded) output columns
//   * X.W on the 5th-gen tensor cores: tcgen05.mma kind::tf32, M=128 N=208 K=8 per instruction, fp32
//     accumulators in TMEM (256 columns per CTA, two CTAs per SM)
//   * float32-equivalent accuracy by the 3xTF32 split: x = hi + lo with hi = x truncated to 10 mantissa bits,
//     X.W ~= Xhi.Whi + Xhi.Wlo + Xlo.Whi (three MMAs per k-step into the same accumulator)
//   * operands in the canonical no-swizzle K-major layout (8-row x 16-byte core matrices): W is pre-split and
//     pre-laid-out on the host so a K-chunk is ONE cp.async.bulk (TMA 1-D bulk copy, completion on an
//     mbarrier); X chunks are split on the fly by the loader threads
//   * epilogue: tcgen05.ld 32x32b -> shared memory tile -> block-diagonal adjacency product, bias, ReLU,
//     optional accumulation into the five-way sum -> global
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tactor {
namespace tc {

constexpr int TCM = 128;             // rows per CTA
constexpr int TCN = 208;             // padded output columns (UMMA N, multiple of 16)
constexpr int KCH = 32;              // K elements per chunk (4 MMA k-steps of 8)
constexpr int NKB = KCH / 4;         // 16-byte core-matrix columns per chunk
constexpr int A_LBO = TCM * 16;      // bytes between core matrices adjacent in K (A operand)
constexpr int B_LBO = TCN * 16;      // same for the B operand
constexpr int SBO = 128;             // bytes between 8-row groups
constexpr int A_BYTES = NKB * A_LBO; // one of {hi, lo}
constexpr int B_BYTES = NKB * B_LBO;
constexpr int LDT = 212;             // padded row length of the epilogue tile
constexpr int THREADS = 256;
constexpr int TMEM_COLS = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

__host__ __device__ constexpr int smem_bytes(int nodes) {
  return 2 * A_BYTES + 2 * B_BYTES + (TCM / nodes) * nodes * nodes * 4 + 64;
}
// W operand image in global memory: per chunk, [hi: kb][n][4 floats] then [lo: ...]
__host__ __device__ constexpr int chunk_kw(int K, int c) { return (K - c * KCH) < KCH ? (K - c * KCH) : KCH; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFFu) << 32) | (1ull << 46);        // version 1 (Blackwell), no swizzle
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 208
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)(TCM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a mis-programmed barrier must not hang the GPU; returns false on timeout
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

template <int NODES>
__global__ void __launch_bounds__(THREADS, 2)
gcn_layer_tc_kernel(const float* __restrict__ X, int K, const float* __restrict__ Wimg,
                    const float* __restrict__ bias, const float* __restrict__ adj, int adj_batched,
                    float* __restrict__ Y, int accumulate, int M, int* __restrict__ error_flag) {
  constexpr int ENVS = TCM / NODES;
  constexpr int LDX = 208;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Ahi = smem;
  unsigned char* Alo = Ahi + A_BYTES;
  unsigned char* Bimg = Alo + A_BYTES;                       // hi then lo, as laid out in global memory
  float* Ad = reinterpret_cast<float*>(Bimg + 2 * B_BYTES);  // [ENVS][NODES(j)][NODES(i)]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ad + ENVS * NODES * NODES);   // [0] bulk copy, [1] MMA done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  float* Ts = reinterpret_cast<float*>(smem);                // epilogue tile [64][LDT], aliases the operands

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TCM;
  const uint32_t bar_b = smem_u32(&bars[0]), bar_m = smem_u32(&bars[1]);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    mbar_init(bar_b, 1);
    mbar_init(bar_m, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < ENVS * NODES * NODES; idx += THREADS) {
    const int e = idx / (NODES * NODES), r = idx % (NODES * NODES), i = r / NODES, j = r % NODES;
    const int env = row0 / NODES + e;
    float v = 0.f;
    if (env * NODES < M) v = adj_batched ? adj[(size_t)env * NODES * NODES + r] : adj[r];
    Ad[(e * NODES + j) * NODES + i] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  const int nchunks = (K + KCH - 1) / KCH;
  const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(Wimg);
  uint32_t phase = 0;
  bool ok = true;
  for (int c = 0; c < nchunks; ++c) {
    const int kw = chunk_kw(K, c), nkb = kw / 4;
    const uint32_t bbytes = 2u * nkb * B_LBO;
    if (tid == 0) {                                          // W chunk: one bulk copy (hi + lo)
      mbar_expect_tx(bar_b, bbytes);
      bulk_g2s(smem_u32(Bimg), wsrc, bbytes, bar_b);
    }
    wsrc += bbytes;
    // X chunk: lanes run along rows (conflict-free 16-byte shared stores), split into hi / lo
    for (int idx = tid; idx < TCM * nkb; idx += THREADS) {
      const int r = idx % TCM, kb = idx / TCM;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < M) v = *reinterpret_cast<const float4*>(X + (size_t)(row0 + r) * LDX + c * KCH + 4 * kb);
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
      *reinterpret_cast<float4*>(Ahi + kb * A_LBO + r * 16) = h;
      *reinterpret_cast<float4*>(Alo + kb * A_LBO + r * 16) = l;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (MMA)
    __syncthreads();
    if (tid == 0) {
      ok = mbar_wait(bar_b, phase) && ok;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(Ahi), a_lo = smem_u32(Alo), b_hi = smem_u32(Bimg), b_lo = b_hi + nkb * B_LBO;
      for (int ks = 0; ks < kw / 8; ++ks) {
        const uint32_t ao = 2 * ks * A_LBO, bo = 2 * ks * B_LBO;
        const uint64_t dah = make_desc(a_hi + ao, A_LBO), dal = make_desc(a_lo + ao, A_LBO);
        const uint64_t dbh = make_desc(b_hi + bo, B_LBO), dbl = make_desc(b_lo + bo, B_LBO);
        mma_tf32(tmem_base, dah, dbh, (c | ks) != 0);
        mma_tf32(tmem_base, dah, dbl, 1);
        mma_tf32(tmem_base, dal, dbh, 1);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_m) : "memory");
    }
    ok = mbar_wait(bar_m, phase) && ok;                      // operands consumed, accumulator updated
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (!ok && error_flag) atomicExch(error_flag, 1);

  // ---- epilogue: two passes of 64 rows -------------------------------------------------------------------
  const int tx = tid % 26, ty = tid / 26;                    // 208 threads own an 8 x 8 output patch
  float4 bs0 = make_float4(0, 0, 0, 0), bs1 = bs0;
  if (tid < 208) {
    bs0 = reinterpret_cast<const float4*>(bias + 8 * tx)[0];
    bs1 = reinterpret_cast<const float4*>(bias + 8 * tx)[1];
  }
  const float bb[8] = {bs0.x, bs0.y, bs0.z, bs0.w, bs1.x, bs1.y, bs1.z, bs1.w};
  for (int p = 0; p < 2; ++p) {
    __syncthreads();                                         // Ts free (previous pass stored / MMAs done)
    const int q = warp & 3;                                  // TMEM lane quarter this warp may read
    if ((q >> 1) == p) {
      const int rl = (q & 1) * 32 + lane;                    // row inside the 64-row half
      const int cbase = (warp >> 2) * 104;
      for (int cc = 0; cc < 104; cc += 8) {
        float v[8];
        tmem_ld8(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(cbase + cc), v);
        float4* dst = reinterpret_cast<float4*>(Ts + rl * LDT + cbase + cc);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    __syncthreads();
    if (tid < 208) {
      float acc[8][8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) acc[i][cidx] = 0.f;
      const int rt = 64 * p + 8 * ty;                        // first row of the patch inside the 128-row tile
      const int e = rt / NODES, ri = rt % NODES;
      const int tbase = (e * NODES - 64 * p);                // env's first row inside the 64-row half
#pragma unroll 4
      for (int j = 0; j < NODES; ++j) {
        const float4 a0 = reinterpret_cast<const float4*>(Ad + (e * NODES + j) * NODES + ri)[0];
        const float4 a1 = reinterpret_cast<const float4*>(Ad + (e * NODES + j) * NODES + ri)[1];
        const float4 t0 = reinterpret_cast<const float4*>(Ts + (tbase + j) * LDT + 8 * tx)[0];
        const float4 t1 = reinterpret_cast<const float4*>(Ts + (tbase + j) * LDT + 8 * tx)[1];
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int cidx = 0; cidx < 8; ++cidx) acc[i][cidx] = fmaf(a[i], t[cidx], acc[i][cidx]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = row0 + rt + i;
        if (row >= M) continue;
        float v[8];
#pragma unroll
        for (int cidx = 0; cidx < 8; ++cidx) v[cidx] = fmaxf(acc[i][cidx] + bb[cidx], 0.f);
        float4* dst = reinterpret_cast<float4*>(Y + (size_t)row * LDX + 8 * tx);
        if (accumulate) {
          const float4 o0 = dst[0], o1 = dst[1];
          v[0] += o0.x; v[1] += o0.y; v[2] += o0.z; v[3] += o0.w;
          v[4] += o1.x; v[5] += o1.y; v[6] += o1.z; v[7] += o1.w;
        }
        dst[0] = make_float4(v[0], v[1], v[2], v[3]);
        dst[1] = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace tc
}  // namespace tactor
