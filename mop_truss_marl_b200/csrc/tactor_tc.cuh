// tcgen05 building blocks and the whole-network fused actor kernel.
//
//   * a CTA owns 128 rows (8 environments of 16 nodes / 4 of 32) and all 208 (padded) output columns
//   * X.W on the 5th-gen tensor cores: tcgen05.mma kind::tf32, M=128 N=208 K=8 per instruction, fp32
//     accumulators in TMEM
//   * float32-equivalent accuracy by the 3xTF32 split: x = hi + lo with hi = x truncated to 10 mantissa bits,
//     X.W ~= Xhi.Whi + Xhi.Wlo + Xlo.Whi (three MMAs per k-step into the same accumulator)
//   * operands in the canonical no-swizzle K-major layout (8-row x 16-byte core matrices): W is pre-split and
//     pre-laid-out on the host so a K-chunk is ONE cp.async.bulk (TMA 1-D bulk copy, completion on an
//     mbarrier); the A operand is generated and split on the fly by the loader threads
//   * epilogue: tcgen05.ld 32x32b -> shared memory tile -> block-diagonal adjacency product, bias, ReLU
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tactor {
namespace tc {

constexpr int TCM = 128;             // rows per CTA
constexpr int TCN = 208;             // padded output columns (UMMA N, multiple of 16)
constexpr int KCH = 16;              // K elements per chunk (2 MMA k-steps of 8)
constexpr int NKB = KCH / 4;         // 16-byte core-matrix columns per chunk
constexpr int A_LBO = TCM * 16;      // bytes between core matrices adjacent in K (A operand)
constexpr int B_LBO = TCN * 16;      // same for the B operand
constexpr int SBO = 128;             // bytes between 8-row groups
constexpr int A_BYTES = NKB * A_LBO; // one of {hi, lo}
constexpr int B_BYTES = NKB * B_LBO;
constexpr int LDT = 212;             // padded row length of the epilogue tile
constexpr int TMEM_COLS = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

constexpr int STAGES = 2;
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // Ahi, Alo, Bhi, Blo of one K chunk
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

// =========================================================================================================
// Whole-network fusion: one CTA carries a 128-row tile (ENVS environments) through all 13 GCN layers.
// Activations never leave the SM:
//   Z   = A_n . x_n                                   [128,13]   shared memory (registers per loader thread)
//   A operand of GEMM g, generated chunk by chunk by the loader threads:
//     g=0  x11 = relu(Z W11 + b11)      g=1,2  x12 = relu(Z W12 + b12)      g=3  x13 = relu(Z W13 + b13)
//     g=4  x14b[r][k] = pooled[env][(n*200+k)/N]       g=5,6  H (the five-way sum, shared memory)
//   D (TMEM) = A . W_g  by tcgen05.mma kind::tf32 with the 3xTF32 split
//   epilogue g: T = D -> shared; V = relu(Adj_g . T + b_g);  g<=4: H (+)= V;  g=5: geo head;  g=6: topo head
// HBM traffic per environment: x_n, three [N,N] adjacencies, pooled row in; 5N floats out.
namespace fused {

constexpr int FTHREADS = 288;                              // warps 0-7: operand generators + epilogue; warp 8: TMA + MMA issuer
constexpr int LDH = 212;                                   // padded row length of the H tile
constexpr int NGEMM = 7;
constexpr int KH = 200;

struct Params {
  const float* x_n;       // [B,N,13]
  const float* A_n;       // [N,N]
  const float* A_s;       // [B,N,N]
  const float* A_ts;      // [B,N,N]
  const float* A_cs;      // [B,N,N]
  const float* pooled;    // [B,208]
  const float* w1[3];     // layer-1 kernels packed [16,208]
  const float* b1[3];     // [208]
  const float* wimg[NGEMM];
  const float* bias[NGEMM];
  const float* w_head[2]; // packed [208,208], first 2 / 3 columns used
  const float* b_head[2];
  float* geo;             // [B,N,2]
  float* topo;            // [B,N,3]
  int M;                  // B*N
  int* error_flag;
};

template <int NODES>
__host__ __device__ constexpr int fused_smem_bytes() {
  return STAGES * STAGE_BYTES + TCM * LDH * 4 + (TCM / NODES) * NODES * NODES * 4 + NODES * NODES * 4 +
         14 * 208 * 4 + (TCM / NODES) * 208 * 4 + 64 * 4 * 4 + 128;
}

template <int NODES>
__global__ void __launch_bounds__(FTHREADS, 1)
actor_fused_kernel(const Params P) {
  constexpr int ENVS = TCM / NODES;
  extern __shared__ __align__(128) unsigned char smem[];
  float* H = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);          // [128][LDH]
  float* Ad = H + TCM * LDH;                                                 // [ENVS][N(j)][N(i)] adjacency of the current GEMM
  float* An = Ad + ENVS * NODES * NODES;                                     // [N(j)][N(i)] shared A_n, transposed
  float* W1s = An + NODES * NODES;                                           // [14][208] layer-1 kernel + bias row of the current GEMM;
                                                                             // for g >= 5: head kernel [200][4] + bias [4]
  float* Pl = W1s + 14 * 208;                                                // [ENVS][208] pooled Pareto embedding
  float* Zs = reinterpret_cast<float*>(smem);                                // [128][16], only until the first stage fill
  float* Us = Pl + ENVS * 208;                                               // [64][4] head pre-activations
  uint64_t* bars = reinterpret_cast<uint64_t*>(Us + 64 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  float* Ts = reinterpret_cast<float*>(smem);                                // [64][LDT], aliases the stages

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TCM;
  const int env0 = row0 / NODES;
  const int M = P.M;
  // mbarriers: [0,1] W chunk landed (TMA tx)  [2,3] stage consumed (tcgen05.commit)
  //            [4,5] A operand written (one arrive per generator warp)  [6] accumulator complete
  const uint32_t bar_b0 = smem_u32(&bars[0]), bar_m0 = smem_u32(&bars[2]), bar_a0 = smem_u32(&bars[4]);
  const uint32_t bar_acc = smem_u32(&bars[6]);
  const bool is_issuer = (warp == 8);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1);
    mbar_init(bar_a0, 8);
    mbar_init(bar_a0 + 8, 8);
    mbar_init(bar_acc, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // shared A_n (transposed), pooled rows, raw x_n rows (staged in H, which is free until the first epilogue)
  for (int idx = tid; idx < NODES * NODES; idx += FTHREADS) An[(idx % NODES) * NODES + idx / NODES] = P.A_n[idx];
  for (int idx = tid; idx < ENVS * 208; idx += FTHREADS) {
    const int env = env0 + idx / 208;
    Pl[idx] = (env * NODES < M) ? P.pooled[(size_t)env * 208 + idx % 208] : 0.f;
  }
  float* Xraw = H;                                                           // [128][13]
  for (int idx = tid; idx < TCM * 13; idx += FTHREADS)
    Xraw[idx] = (row0 + idx / 13 < M) ? P.x_n[(size_t)row0 * 13 + idx] : 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  // Z = A_n . x_n  (gcn_l1_k share the input and the adjacency, so the product is formed once)
  for (int idx = tid; idx < TCM * 16; idx += FTHREADS) {
    const int r = idx / 16, i = idx % 16, e = r / NODES, n = r % NODES;
    float z = 0.f;
    if (i < 13)
      for (int j = 0; j < NODES; ++j) z = fmaf(An[j * NODES + n], Xraw[(e * NODES + j) * 13 + i], z);
    Zs[idx] = z;
  }
  __syncthreads();
  const int lr = tid % TCM, lkb = tid / TCM;                 // loader role: row lr, core columns lkb and lkb + 2
  float zr[13];
#pragma unroll
  for (int i = 0; i < 13; ++i) zr[i] = Zs[lr * 16 + i];
  const int l_env = lr / NODES, l_n = lr % NODES;

  // ---- A operand generator: 4 consecutive k of row lr for GEMM g -------------------------------------------
  auto gen_a = [&](int g, int k) -> float4 {
    if (g <= 3) {
      const float* w = W1s + k;
      float4 acc = *reinterpret_cast<const float4*>(W1s + 13 * 208 + k);
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        const float4 wv = *reinterpret_cast<const float4*>(w + i * 208);
        acc.x = fmaf(zr[i], wv.x, acc.x); acc.y = fmaf(zr[i], wv.y, acc.y);
        acc.z = fmaf(zr[i], wv.z, acc.z); acc.w = fmaf(zr[i], wv.w, acc.w);
      }
      return make_float4(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f), fmaxf(acc.z, 0.f), fmaxf(acc.w, 0.f));
    } else if (g == 4) {
      const float* pl = Pl + l_env * 208;
      const int f = l_n * KH + k;
      return make_float4(pl[f / NODES], pl[(f + 1) / NODES], pl[(f + 2) / NODES], pl[(f + 3) / NODES]);
    }
    return *reinterpret_cast<const float4*>(H + lr * LDH + k);
  };
  auto split_store = [&](int stage, int kb, const float4& v) {
    unsigned char* Ahi = smem + stage * STAGE_BYTES;
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
    *reinterpret_cast<float4*>(Ahi + kb * A_LBO + lr * 16) = h;
    *reinterpret_cast<float4*>(Ahi + A_BYTES + kb * A_LBO + lr * 16) = l;
  };
  auto fill_stage = [&](int g, int c, int stage) {
    const int nkb = chunk_kw(KH, c) / 4;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int kb = lkb + 2 * q;
      if (kb < nkb) split_store(stage, kb, gen_a(g, c * KCH + 4 * kb));
    }
  };
  auto issue_w = [&](int g, int c, int stage) {
    const uint32_t bytes = 2u * (chunk_kw(KH, c) / 4) * B_LBO;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.wimg[g]) + (size_t)c * (2u * NKB * B_LBO);
    const uint32_t bar = bar_b0 + 8 * stage;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(smem + stage * STAGE_BYTES + 2 * A_BYTES), src, bytes, bar);
  };

  // ---- per-GEMM small operands: adjacency tile (<= 8 floats per thread), layer-1 kernel + bias or head
  //      kernel + bias (<= 3 float4 per thread); loaded one GEMM ahead so their latency hides in the epilogue
  constexpr int ADJ_PER = (ENVS * NODES * NODES + FTHREADS - 1) / FTHREADS;
  constexpr int W1_PER = (14 * 208 / 4 + FTHREADS - 1) / FTHREADS;
  float adj_reg[ADJ_PER];
  float4 w1_reg[W1_PER];
  auto prefetch_small = [&](int g) {
    const float* adj = (g == 1) ? P.A_ts : (g == 2) ? P.A_cs : (g == 3 || g == 6) ? P.A_s : nullptr;
#pragma unroll
    for (int q = 0; q < ADJ_PER; ++q) {
      const int idx = tid + q * FTHREADS;
      float v = 0.f;
      if (adj != nullptr && idx < ENVS * NODES * NODES) {
        const int e = idx / (NODES * NODES), r = idx % (NODES * NODES);
        if ((env0 + e) * NODES < M) v = __ldg(adj + (size_t)(env0 + e) * NODES * NODES + r);
      }
      adj_reg[q] = v;
    }
    if (g == 0 || g == 1 || g == 3) {                        // g = 2 reuses gcn_l1_2's kernel
      const int l1 = g == 0 ? 0 : (g == 3 ? 2 : 1);
#pragma unroll
      for (int q = 0; q < W1_PER; ++q) {
        const int idx = tid + q * FTHREADS;
        w1_reg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < 13 * 52) w1_reg[q] = __ldg(reinterpret_cast<const float4*>(P.w1[l1]) + idx);
        else if (idx < 14 * 52) w1_reg[q] = __ldg(reinterpret_cast<const float4*>(P.b1[l1]) + (idx - 13 * 52));
      }
    } else if (g >= 5) {                                     // head kernel rows [k][0..3] (+ bias as row 200)
      const int hd = g - 5;
#pragma unroll
      for (int q = 0; q < W1_PER; ++q) {
        const int k = tid + q * FTHREADS;
        w1_reg[q] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < KH) w1_reg[q] = __ldg(reinterpret_cast<const float4*>(P.w_head[hd] + (size_t)k * 208));
        else if (k == KH) w1_reg[q] = __ldg(reinterpret_cast<const float4*>(P.b_head[hd]));
      }
    }
  };
  auto commit_small = [&](int g) {
    const bool shared_an = !(g == 1 || g == 2 || g == 3 || g == 6);
#pragma unroll
    for (int q = 0; q < ADJ_PER; ++q) {
      const int idx = tid + q * FTHREADS;
      if (idx < ENVS * NODES * NODES) {
        const int e = idx / (NODES * NODES), r = idx % (NODES * NODES), i = r / NODES, j = r % NODES;
        Ad[(e * NODES + j) * NODES + i] = shared_an ? An[j * NODES + i] : adj_reg[q];
      }
    }
    if (g == 0 || g == 1 || g == 3) {
#pragma unroll
      for (int q = 0; q < W1_PER; ++q) {
        const int idx = tid + q * FTHREADS;
        if (idx < 14 * 52) reinterpret_cast<float4*>(W1s)[idx] = w1_reg[q];
      }
    } else if (g >= 5) {
#pragma unroll
      for (int q = 0; q < W1_PER; ++q) {
        const int k = tid + q * FTHREADS;
        if (k <= KH) reinterpret_cast<float4*>(W1s)[k] = w1_reg[q];
      }
    }
  };

#ifdef DEBUG_TIMING
  const long long Tstart = clock64();
#endif
  constexpr int NCH = (KH + KCH - 1) / KCH;                  // 13 chunks per GEMM
  uint32_t use = 0;                                          // running count of stage uses (both barriers flip per use)
  bool ok = true;
  const int cq = tid % 52, gq = tid / 52;                    // epilogue role (tid < 208): rows 16*gq.., columns 4*cq..

  for (int g = 0; g < NGEMM; ++g) {
    // adjacency (transposed per environment: A_n for g = 0, 4, 5; A_ts, A_cs, A_s otherwise) and the small
    // weights of this GEMM were prefetched into registers during the previous epilogue
#ifdef DEBUG_TIMING
    const long long T0 = clock64();
#endif
    if (g == 0) prefetch_small(0);
    commit_small(g);
    __syncthreads();
#ifdef DEBUG_TIMING
    const long long T1 = clock64();
#endif
    // ---- main loop, warp-specialised: generators fill the A operand two chunks ahead, the issuer warp
    //      streams the W chunks (TMA) and issues the MMAs; the only hand-offs are mbarriers ----
    const uint32_t use0 = use;
    if (is_issuer) {
      if (lane == 0) {
        // Measured with clock64 (round 1): issuing is latency-bound on this one thread, so the operand
        // descriptors are formed outside the loop, and the W chunk for c+1 is requested BEFORE the MMAs of
        // chunk c are issued so that it has a whole iteration to land.
        uint64_t da[2][2][2], db[2][2][2];                   // [stage][k-step][hi, lo]
#pragma unroll
        for (int st = 0; st < 2; ++st)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_hi = smem_u32(smem + st * STAGE_BYTES) + 2 * ks * A_LBO;
            const uint32_t b_hi = smem_u32(smem + st * STAGE_BYTES) + 2 * A_BYTES + 2 * ks * B_LBO;
            da[st][ks][0] = make_desc(a_hi, A_LBO);
            da[st][ks][1] = make_desc(a_hi + A_BYTES, A_LBO);
            db[st][ks][0] = make_desc(b_hi, B_LBO);
            db[st][ks][1] = make_desc(b_hi + NKB * B_LBO, B_LBO);      // lo half of a full chunk
          }
        issue_w(g, 0, use0 & 1);
        for (int c = 0; c < NCH; ++c) {
          const uint32_t u = use0 + c, s = u & 1, parity = (u >> 1) & 1;
          if (c + 1 < NCH) {
            // W chunk c+1 goes into the other stage: its previous user (chunk c-1) must have been consumed
            if (c >= 1) ok = mbar_wait(bar_m0 + 8 * (s ^ 1), ((u - 1) >> 1) & 1) && ok;
            issue_w(g, c + 1, s ^ 1);
          }
          ok = mbar_wait(bar_a0 + 8 * s, parity) && ok;
          ok = mbar_wait(bar_b0 + 8 * s, parity) && ok;
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const int kw = chunk_kw(KH, c);
          if (kw == KCH) {
            mma_tf32(tmem_base, da[s][0][0], db[s][0][0], c != 0);
            mma_tf32(tmem_base, da[s][0][0], db[s][0][1], 1);
            mma_tf32(tmem_base, da[s][0][1], db[s][0][0], 1);
            mma_tf32(tmem_base, da[s][1][0], db[s][1][0], 1);
            mma_tf32(tmem_base, da[s][1][0], db[s][1][1], 1);
            mma_tf32(tmem_base, da[s][1][1], db[s][1][0], 1);
          } else {                                           // tail chunk: one k-step, lo half right after hi
            const uint32_t b_hi = smem_u32(smem + s * STAGE_BYTES) + 2 * A_BYTES;
            const uint64_t dbl = make_desc(b_hi + (kw / 4) * B_LBO, B_LBO);
            mma_tf32(tmem_base, da[s][0][0], db[s][0][0], c != 0);
            mma_tf32(tmem_base, da[s][0][0], dbl, 1);
            mma_tf32(tmem_base, da[s][0][1], db[s][0][0], 1);
          }
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_m0 + 8 * s)
                       : "memory");
          if (c + 1 == NCH)
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_acc)
                         : "memory");
        }
      }
      __syncwarp();
    } else {
      for (int c = 0; c < NCH; ++c) {
        const uint32_t u = use0 + c, s = u & 1;
        if (c >= 2) ok = mbar_wait(bar_m0 + 8 * s, ((u - 2) >> 1) & 1) && ok;   // chunk c-2 consumed
        fill_stage(g, c, s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_a0 + 8 * s) : "memory");
      }
    }
    use = use0 + NCH;
    ok = mbar_wait(bar_acc, (uint32_t)(g & 1)) && ok;        // every MMA of this GEMM has completed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#ifdef DEBUG_TIMING
    const long long T2 = clock64();
#endif
    if (g + 1 < NGEMM) prefetch_small(g + 1);                // latency hidden behind the epilogue

    // ---- epilogue ----
    float4 bs0 = make_float4(0, 0, 0, 0);
    if (tid < 208) bs0 = __ldg(reinterpret_cast<const float4*>(P.bias[g] + 4 * cq));
    const float bb[4] = {bs0.x, bs0.y, bs0.z, bs0.w};
    for (int p = 0; p < 2; ++p) {
      __syncthreads();                                       // Ts / Us free
      const int q = warp & 3;
      if (warp < 8 && (q >> 1) == p) {
        const int rl = (q & 1) * 32 + lane;
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
      // each thread owns 16 rows (one environment, or half of a 32-node one) x 4 columns: every T row is
      // read once per 16 output rows and the adjacency values are warp broadcasts
      float acc[16][4];
      const int rt = 64 * p + 16 * gq;                       // first row of the patch inside the 128-row tile
      if (tid < 208) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
#pragma unroll
          for (int cidx = 0; cidx < 4; ++cidx) acc[i][cidx] = 0.f;
        const int e = rt / NODES, ri = rt % NODES;
        const int tbase = e * NODES - 64 * p;
#pragma unroll 2
        for (int j = 0; j < NODES; ++j) {
          const float4 t = *reinterpret_cast<const float4*>(Ts + (tbase + j) * LDT + 4 * cq);
          const float4* ap = reinterpret_cast<const float4*>(Ad + (e * NODES + j) * NODES + ri);
          const float4 a0 = ap[0], a1 = ap[1], a2 = ap[2], a3 = ap[3];
          const float a[16] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w, a3.x, a3.y, a3.z, a3.w};
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            acc[i][0] = fmaf(a[i], t.x, acc[i][0]); acc[i][1] = fmaf(a[i], t.y, acc[i][1]);
            acc[i][2] = fmaf(a[i], t.z, acc[i][2]); acc[i][3] = fmaf(a[i], t.w, acc[i][3]);
          }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
#pragma unroll
          for (int cidx = 0; cidx < 4; ++cidx) acc[i][cidx] = fmaxf(acc[i][cidx] + bb[cidx], 0.f);
      }
      if (g <= 4) {
        if (tid < 208) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float4* dst = reinterpret_cast<float4*>(H + (rt + i) * LDH + 4 * cq);
            float4 o0 = make_float4(0, 0, 0, 0);
            if (g > 0) o0 = dst[0];
            dst[0] = make_float4(acc[i][0] + o0.x, acc[i][1] + o0.y, acc[i][2] + o0.z, acc[i][3] + o0.w);
          }
        }
      } else {
        // output heads (truss2D_RL.py:121-125): sigmoid(A_n (x3 W4) + b4); x3 goes back through Ts
        const int hd = g - 5, nout = 2 + hd;
        __syncthreads();                                     // every thread is done reading T
        if (tid < 208) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            *reinterpret_cast<float4*>(Ts + (16 * gq + i) * LDT + 4 * cq) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        }
        __syncthreads();
        if (tid < 64 * nout) {
          const int r = tid / nout, o = tid % nout;
          const float* w = W1s + o;                          // head kernel staged as [k][4]
          float u0 = 0.f, u1 = 0.f, u2 = 0.f, u3 = 0.f;
          for (int k = 0; k < KH; k += 4) {
            const float4 x = *reinterpret_cast<const float4*>(Ts + r * LDT + k);
            u0 = fmaf(x.x, w[4 * k], u0); u1 = fmaf(x.y, w[4 * k + 4], u1);
            u2 = fmaf(x.z, w[4 * k + 8], u2); u3 = fmaf(x.w, w[4 * k + 12], u3);
          }
          Us[r * 4 + o] = (u0 + u1) + (u2 + u3);
        }
        __syncthreads();
        if (tid < 64 * nout) {
          const int r = tid / nout, o = tid % nout;
          const int e = r / NODES, n = r % NODES;
          float v = 0.f;
          for (int j = 0; j < NODES; ++j) v = fmaf(An[j * NODES + n], Us[(e * NODES + j) * 4 + o], v);
          v += W1s[4 * KH + o];
          const int row = row0 + 64 * p + r;
          if (row < M) (hd == 0 ? P.geo : P.topo)[(size_t)row * nout + o] = 1.f / (1.f + expf(-v));
        }
      }
    }
    __syncthreads();                                         // H / Ts settled before the next GEMM refills the stages
#ifdef DEBUG_TIMING
    if (blockIdx.x == 200 && (tid == 0 || tid == 256)) {
      long long* dbg = reinterpret_cast<long long*>(P.error_flag) + 16 + (tid ? 64 : 0);
      dbg[g * 4 + 0] = T1 - T0; dbg[g * 4 + 1] = T2 - T1; dbg[g * 4 + 2] = clock64() - T2; dbg[g * 4 + 3] = T0 - Tstart;
    }
#endif
  }
  if (!ok && P.error_flag) atomicExch(P.error_flag, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace fused

}  // namespace tc
}  // namespace tactor
