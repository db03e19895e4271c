// tcgen05 building blocks and the whole-network fused actor kernel.
//
//   * a CTA owns 128 rows (8 environments of 16 nodes / 4 of 32) and all 208 (padded) output columns
//   * X.W on the 5th-gen tensor cores: tcgen05.mma kind::tf32, M=128 N=208 K=8 per instruction, fp32
//     accumulators in TMEM
//   * float32-equivalent accuracy by the 3xTF32 split: x = hi + lo with hi = x truncated to 10 mantissa bits,
//     X.W ~= Xhi.Whi + Xhi.Wlo + Xlo.Whi (three MMAs per k-step into the same accumulator)
//   * the B operand (W) sits in shared memory in the canonical no-swizzle K-major layout (8-row x 16-byte core
//     matrices): W is pre-split and pre-laid-out on the host so a K-chunk is ONE cp.async.bulk (TMA 1-D bulk
//     copy, completion on an mbarrier)
//   * the A operand is generated and split on the fly by the generator threads and written straight into
//     TENSOR MEMORY (tcgen05.st, lane = row, column = k): it never touches shared memory, whose bandwidth the
//     tensor core's B reads, the TMA writes and the epilogue already compete for
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
constexpr int SBO = 128;             // bytes between 8-row groups
constexpr int A_BYTES = NKB * A_LBO; // one of {hi, lo}
constexpr int LDT = 212;             // padded row length of the epilogue tile
constexpr int TMEM_COLS = 512;       // [0,208) accumulator, [256, 256 + 32 * AST) A-operand stages
constexpr int AST = 4;               // A-operand stages in tensor memory: per stage [hi: 16 columns][lo: 16 columns]
constexpr int TM_A0 = 256;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

// NCTA = 1: one CTA per 128-row tile.  NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) works on two
// tiles with ONE M = 256 instruction stream: each CTA generates the A operand of its own 128 rows and holds only
// its half of the W columns, so the weight stream out of L2 and the tensor core's B reads per SM are halved.
template <int NCTA> struct Cfg {
  static constexpr int BN = TCN / NCTA;                     // W columns held by one CTA
  static constexpr int B_LBO = BN * 16;                     // bytes between core matrices adjacent in K (B operand)
  static constexpr int B_BYTES = NKB * B_LBO;               // one of {hi, lo} of a full chunk
  static constexpr int WST = NCTA == 1 ? 3 : 5;             // W stages in shared memory
  static constexpr int STAGE_BYTES = 2 * B_BYTES;           // Bhi, Blo of one K chunk
  static constexpr int CHUNK_IMG_BYTES = NCTA * 2 * B_BYTES;      // one full chunk of the W image (all CTAs)
  // kind::tf32, fp32 accumulate, A and B K-major, M = 128 * NCTA, N = 208
  static constexpr uint32_t IDESC =
      (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TCN >> 3) << 17) | ((uint32_t)((TCM * NCTA) >> 4) << 24);
};
// W operand image in global memory: per chunk, per CTA of the pair, [hi: kb][n (BN)][4 floats] then [lo: ...]
__host__ __device__ constexpr int chunk_kw(int K, int c) { return (K - c * KCH) < KCH ? (K - c * KCH) : KCH; }

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((SBO >> 4) & 0x3FFFu) << 32) | (1ull << 46);        // version 1 (Blackwell), no swizzle
}
// D[tmem] (+)= A[tmem] . B[smem]: A is [128 lanes = rows][8 columns = k] of tensor memory (of each CTA of a pair)
template <int NCTA>
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t accumulate) {
  if constexpr (NCTA == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(Cfg<1>::IDESC), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(Cfg<2>::IDESC), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
// all MMAs issued so far by this thread have completed -> one arrival on the barrier (of every CTA of the pair)
template <int NCTA>
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  if constexpr (NCTA == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
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
// the same for a barrier that peer-CTA threads arrive on
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; spin < SPIN_LIMIT; ++spin) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) return true;
  }
  return false;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(bar), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
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

constexpr int FTHREADS = 320;                              // warps 0-7: operand generators + epilogue; warp 8: MMA issuer
                                                           // (peer CTA: W-landed forwarder); warp 9: W producer (TMA)
constexpr int LDH = 204;                                   // padded row length of the H tile
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

template <int NODES, int NCTA>
__host__ __device__ constexpr int fused_smem_bytes() {
  return Cfg<NCTA>::WST * Cfg<NCTA>::STAGE_BYTES + TCM * LDH * 4 + (TCM / NODES) * NODES * NODES * 4 + NODES * NODES * 4 +
         14 * 208 * 4 + (TCM / NODES) * 208 * 4 + 64 * 4 * 4 + 256;
}

template <int NODES, int NCTA>
__global__ void __launch_bounds__(FTHREADS, 1)
actor_fused_kernel(const Params P) {
  constexpr int ENVS = TCM / NODES;
  constexpr int WST = Cfg<NCTA>::WST, STAGE_BYTES = Cfg<NCTA>::STAGE_BYTES;
  constexpr int B_LBO = Cfg<NCTA>::B_LBO;
  static_assert(WST * STAGE_BYTES >= 64 * LDT * 4, "the epilogue tile aliases the W stages");
  extern __shared__ __align__(128) unsigned char smem[];
  float* H = reinterpret_cast<float*>(smem + WST * STAGE_BYTES);             // [128][LDH]
  float* Ad = H + TCM * LDH;                                                 // [ENVS][N(j)][N(i)] adjacency of the current GEMM
  float* An = Ad + ENVS * NODES * NODES;                                     // [N(j)][N(i)] shared A_n, transposed
  float* W1s = An + NODES * NODES;                                           // [14][208] layer-1 kernel + bias row of the current GEMM;
                                                                             // for g >= 5: head kernel [200][4] + bias [4]
  float* Pl = W1s + 14 * 208;                                                // [ENVS][208] pooled Pareto embedding
  float* Zs = reinterpret_cast<float*>(smem);                                // [128][16], only until the first stage fill
  float* Us = Pl + ENVS * 208;                                               // [64][4] head pre-activations
  uint64_t* bars = reinterpret_cast<uint64_t*>(Us + 64 * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);
  float* Ts = reinterpret_cast<float*>(smem);                                // [64][LDT], aliases the stages

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * TCM;
  const int env0 = row0 / NODES;
  const int M = P.M;
  // mbarriers.  W ring (stage s < WST):   bar_b0+8s  this CTA's W half landed (TMA tx)
  //                                       bar_p0+8s  the peer CTA's W half landed (leader's copy, peer arrives)
  //                                       bar_m0+8s  W stage consumed (tcgen05.commit, multicast to the pair)
  //             A ring (stage s < AST):   bar_a0+8s  A operand written to tensor memory (leader's copy: one arrive
  //                                                  per generator warp of every CTA of the pair)
  //                                       bar_e0+8s  A stage consumed (tcgen05.commit, multicast)
  //             bar_acc                              accumulator complete
  const uint32_t bar_b0 = smem_u32(&bars[0]), bar_p0 = smem_u32(&bars[5]), bar_m0 = smem_u32(&bars[10]);
  const uint32_t bar_a0 = smem_u32(&bars[15]), bar_e0 = smem_u32(&bars[20]), bar_acc = smem_u32(&bars[25]);
  const bool is_issuer = (warp == 8), is_producer = (warp == 9);
  const uint32_t cta_rank = (NCTA == 1) ? 0u : cluster_ctarank();
  const bool is_leader = (cta_rank == 0);

  if (warp == 0) {
    if constexpr (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(TMEM_COLS)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (tid == 32) {
    for (int s = 0; s < WST; ++s) {
      mbar_init(bar_b0 + 8 * s, 1);
      mbar_init(bar_p0 + 8 * s, 1);
      mbar_init(bar_m0 + 8 * s, 1);
    }
    for (int s = 0; s < AST; ++s) {
      mbar_init(bar_a0 + 8 * s, 8 * NCTA);
      mbar_init(bar_e0 + 8 * s, 1);
    }
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
  if constexpr (NCTA == 2) cluster_sync_all();               // the peer's barriers are initialised before any remote arrive
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
  const int lr = tid % TCM, lkb = (tid / TCM) & 1;           // generator role: row lr, k = 8 * lkb .. 8 * lkb + 7 of a chunk
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
  // 3xTF32 split of 8 consecutive k of row lr, written to A stage `stage` of tensor memory
  auto fill_stage = [&](int g, int c, int stage) {
    if (8 * lkb < chunk_kw(KH, c)) {                         // warp-uniform (tail chunk: k-step 0 only)
      const float4 v0 = gen_a(g, c * KCH + 8 * lkb), v1 = gen_a(g, c * KCH + 8 * lkb + 4);
      const float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hi[i] = __uint_as_float(__float_as_uint(v[i]) & 0xFFFFE000u);
        lo[i] = v[i] - hi[i];
      }
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(TM_A0 + 32 * stage + 8 * lkb);
      tmem_st8(taddr, hi);
      tmem_st8(taddr + 16, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  };
  auto issue_w = [&](int g, int c, int stage) {
    const uint32_t bytes = 2u * (chunk_kw(KH, c) / 4) * B_LBO;                  // this CTA's half: hi then lo
    const unsigned char* src = reinterpret_cast<const unsigned char*>(P.wimg[g]) +
                               (size_t)c * Cfg<NCTA>::CHUNK_IMG_BYTES + (size_t)cta_rank * bytes;
    const uint32_t bar = bar_b0 + 8 * stage;
    mbar_expect_tx(bar, bytes);
    bulk_g2s(smem_u32(smem + stage * STAGE_BYTES), src, bytes, bar);
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
  long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define DBG_T(x) const long long x = clock64()
#define DBG_ACC(i, v) dbg_acc[i] += (v)
#else
#define DBG_T(x)
#define DBG_ACC(i, v)
#endif
  constexpr int NCH = (KH + KCH - 1) / KCH;                  // 13 chunks per GEMM
  uint32_t use = 0;                                          // running count of chunks: W stage = use % WST, A stage = use % AST
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
    // ---- main loop, warp-specialised; the only hand-offs are mbarriers.  For the running chunk count u: W stage
    //      u % WST (phase parity (u / WST) & 1), A stage u % AST (parity (u / AST) & 1).
    //        generators (warps 0-7)   A operand of chunk c -> tensor memory, up to AST chunks ahead
    //        producer   (warp 9)      this CTA's W half of chunk c -> shared memory (TMA), up to WST chunks ahead
    //        issuer     (warp 8)      leader CTA: MMAs of the whole pair;  peer CTA: forwards "my W half landed" ----
    const uint32_t use0 = use;
    if (is_producer) {
      if (lane == 0) {
        for (int c = 0; c < NCH; ++c) {
          const uint32_t u = use0 + c, s = u % WST;
          DBG_T(t0);
          if (c >= WST) ok = mbar_wait(bar_m0 + 8 * s, ((u / WST) - 1) & 1) && ok;    // chunk c-WST consumed
          DBG_T(t1);
          issue_w(g, c, s);
          DBG_T(t2);
          DBG_ACC(3, t1 - t0); DBG_ACC(4, t2 - t1);
        }
      }
      __syncwarp();
    } else if (is_issuer) {
      if (lane == 0) {
        if (is_leader) {
          uint64_t db[WST][2][2];                            // [stage][k-step][hi, lo] of a full chunk
#pragma unroll
          for (int st = 0; st < WST; ++st)
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {
              const uint32_t b_hi = smem_u32(smem + st * STAGE_BYTES) + 2 * ks * B_LBO;
              db[st][ks][0] = make_desc(b_hi, B_LBO);
              db[st][ks][1] = make_desc(b_hi + NKB * B_LBO, B_LBO);
            }
          for (int c = 0; c < NCH; ++c) {
            const uint32_t u = use0 + c, sw = u % WST, sa = u % AST;
            DBG_T(t0);
            ok = mbar_wait_cluster(bar_a0 + 8 * sa, (u / AST) & 1) && ok;
            DBG_T(t1);
            ok = mbar_wait(bar_b0 + 8 * sw, (u / WST) & 1) && ok;
            if constexpr (NCTA == 2) ok = mbar_wait_cluster(bar_p0 + 8 * sw, (u / WST) & 1) && ok;
            DBG_T(t2);
            DBG_ACC(0, t1 - t0); DBG_ACC(1, t2 - t1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int kw = chunk_kw(KH, c);
            const uint32_t a_hi = tmem_base + (uint32_t)(TM_A0 + 32 * sa), a_lo = a_hi + 16;
#pragma unroll
            for (int st = 0; st < WST; ++st) {
              if (st != (int)sw) continue;                   // compile-time stage index keeps the descriptors in registers
              if (kw == KCH) {
                mma_tf32<NCTA>(tmem_base, a_hi, db[st][0][0], c != 0);
                mma_tf32<NCTA>(tmem_base, a_hi, db[st][0][1], 1);
                mma_tf32<NCTA>(tmem_base, a_lo, db[st][0][0], 1);
                mma_tf32<NCTA>(tmem_base, a_hi + 8, db[st][1][0], 1);
                mma_tf32<NCTA>(tmem_base, a_hi + 8, db[st][1][1], 1);
                mma_tf32<NCTA>(tmem_base, a_lo + 8, db[st][1][0], 1);
              } else {                                       // tail chunk: one k-step, lo half right after hi
                const uint64_t dbl = make_desc(smem_u32(smem + st * STAGE_BYTES) + (kw / 4) * B_LBO, B_LBO);
                mma_tf32<NCTA>(tmem_base, a_hi, db[st][0][0], c != 0);
                mma_tf32<NCTA>(tmem_base, a_hi, dbl, 1);
                mma_tf32<NCTA>(tmem_base, a_lo, db[st][0][0], 1);
              }
            }
            mma_commit<NCTA>(bar_m0 + 8 * sw);
            mma_commit<NCTA>(bar_e0 + 8 * sa);
            if (c + 1 == NCH) mma_commit<NCTA>(bar_acc);
            DBG_T(t3);
            DBG_ACC(2, t3 - t2);
          }
        } else {
          for (int c = 0; c < NCH; ++c) {                    // peer CTA: tell the leader that W chunk c has landed here
            const uint32_t u = use0 + c, sw = u % WST;
            ok = mbar_wait(bar_b0 + 8 * sw, (u / WST) & 1) && ok;
            mbar_arrive_remote(bar_p0 + 8 * sw, 0);
          }
        }
      }
      __syncwarp();
    } else {
      for (int c = 0; c < NCH; ++c) {
        const uint32_t u = use0 + c, sa = u % AST;
        DBG_T(t0);
        if (c >= AST) ok = mbar_wait(bar_e0 + 8 * sa, ((u / AST) - 1) & 1) && ok;     // chunk c-AST consumed
        DBG_T(t1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        fill_stage(g, c, sa);
        DBG_T(t2);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(bar_a0 + 8 * sa);
          else mbar_arrive_remote(bar_a0 + 8 * sa, 0);
        }
        DBG_T(t3);
        DBG_ACC(5, t1 - t0); DBG_ACC(6, t2 - t1); DBG_ACC(7, t3 - t2);
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
        if (tid < 208 && 4 * cq < LDH) {                      // the last column group is padding beyond the H row
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
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // Ts (generic proxy) before the next TMA writes into the stages
    __syncthreads();                                         // H / Ts settled before the next GEMM refills the stages
#ifdef DEBUG_TIMING
    if (blockIdx.x == 200 && (tid == 0 || tid == 256 || tid == 288)) {
      long long* dbg = reinterpret_cast<long long*>(P.error_flag) + 16 + (tid == 0 ? 0 : tid == 256 ? 128 : 256);
      dbg[g * 4 + 0] = T1 - T0; dbg[g * 4 + 1] = T2 - T1; dbg[g * 4 + 2] = clock64() - T2; dbg[g * 4 + 3] = T0 - Tstart;
      for (int i = 0; i < 8; ++i) dbg[32 + g * 8 + i] = dbg_acc[i];      // cumulative over GEMMs 0..g
    }
#endif
  }
  if (!ok && P.error_flag) atomicExch(P.error_flag, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();               // neither CTA leaves while the pair's TMEM / barriers are in use
  if (warp == 0) {
    if constexpr (NCTA == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace fused

}  // namespace tc
}  // namespace tactor
