// Whole-network fused actor kernel, fully pipelined ("pipe") formulation.
//
// One CTA (or CTA pair, NCTA = 2) carries a 128-row tile (ENVS environments of NODES nodes) through all 13
// GCNConv layers of multimodes_actor.call (train/code/truss2D_RL.py:75-127).  The adjacency product is moved IN FRONT
// of the dense contraction,
//
//      A_g . (X_g . W_g) + b_g   ==   [A_g . X_g, 1] . [W_g ; b_g]        (spektral GCNConv, associativity)
//
// so that the tensor-core epilogue is row-local (scale, ReLU, accumulate) and four groups of warps run concurrently,
// coupled only by mbarriers:
//
//   generators  warps 0-15   row group q = warp & 3 (row r = 32 q + lane), k-half kh, chunk parity par: a warp owns
//                            every other 16-wide K chunk and 8 k of it per thread:
//                            x = layer-1 activations / Pareto embedding / H  ->  y = sum_j A_g[r,j] x[j]
//                            (neighbour rows live in the same warp: exchange through two warp-private tiles,
//                            __syncwarp only; <= 8 neighbours per row go through a compacted list, denser rows
//                            through the full row)  ->  y = hi + lo in fp16  ->  tcgen05.st into the A stage (TMEM,
//                            two k per column).  The stage is acquired right before the store and published in
//                            the middle of the warp's next chunk (hand_off), off the dependent chain.
//   producer    warp 21      streams the pre-split W chunks into shared memory (cp.async.bulk, mbarrier tx),
//                            running ahead across GEMM boundaries
//   issuer      warp 20      tcgen05.mma kind::f16 (Yhi.Whi + Yhi.Wlo + Ylo.Whi), A from tensor memory, B from
//                            shared memory, accumulator g&1 of two (TMEM columns [0,208) and [256,464))
//   epilogue    warps 16-19  tcgen05.ld the finished accumulator while the next GEMM is already running:
//                            relu(D / S) -> H (+)= (g <= 4), or the sigmoid heads gcn_l4_1/2 (g = 5, 6); the bias is
//                            row 200 of the W image (A's column 200 is the constant 1)
//
// TMEM map (512 columns): [0,208) acc0 | [208,256) A stages hi (6 x 8 columns of packed pairs) | [256,464) acc1
//                         | [464,512) A stages lo
#pragma once
#include <cuda_fp16.h>

#include "tactor_tc.cuh"

namespace tactor {
namespace tc {
namespace pipe {

using fused::Params;
using fused::NGEMM;
using fused::KH;

constexpr int PTHREADS = 704;
constexpr int NGENW = 16, NEPIW = 4, W_ISSUER = 20, W_PRODUCER = 21;
constexpr int KPT = 8;                                      // k values per generator thread and chunk (a k-half)
constexpr int PAST = 6;                                     // A-operand stages in tensor memory
constexpr int ACOLS = 8;                                    // TMEM columns of one A stage (two fp16 k per 32-bit column)
constexpr int MAXST = 6;                                    // barrier slots per ring
constexpr int TM_ACC1 = 256, TM_AHI = 208, TM_ALO = 464;
constexpr int LDH = 204;                                    // padded row length of the H tile (12 r mod 32 distinct for 8 rows)
constexpr int LDX = 4;                                      // row length of a warp-private exchange tile (two tiles per warp)
constexpr int NCH = (KH + KCH - 1) / KCH;                   // 13 chunks per GEMM
constexpr int DMAX = 8;                                     // neighbour slots of the compacted adjacency row

template <int NODES, int NCTA>
__host__ __device__ constexpr int pipe_smem_bytes() {
  return Cfg<NCTA>::WST * Cfg<NCTA>::STAGE_BYTES + TCM * LDH * 4 + 2 * NGENW * 32 * LDX * 4 + 3 * 14 * 208 * 4 +
         2 * (TCM / NODES) * 208 * 4 + 2 * TCM * 13 * 4 + NODES * NODES * 4 + 2 * 201 * 4 * 4 + TCM * 4 * 4 + 384;
}

__device__ __forceinline__ void tmem_st4u(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3])
               : "memory");
}
__device__ __forceinline__ uint32_t h2_bits(const __half2& h) { return *reinterpret_cast<const uint32_t*>(&h); }
// tcgen05.ld of 8 accumulator columns without the wait (software pipelining in the epilogue)
__device__ __forceinline__ void tmem_ld8_async(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}

template <int NODES, int NCTA>
__global__ void __launch_bounds__(PTHREADS, 1)
actor_pipe_kernel(const __grid_constant__ Params P) {
  constexpr int ENVS = TCM / NODES;
  constexpr int WST = Cfg<NCTA>::WST, STAGE_BYTES = Cfg<NCTA>::STAGE_BYTES, B_LBO = Cfg<NCTA>::B_LBO;
  static_assert(pipe_smem_bytes<NODES, NCTA>() <= 232448, "shared memory budget");
  static_assert(KH % KPT == 0 && KCH == 2 * KPT, "16-wide k-steps, two k-halves per chunk");
  extern __shared__ __align__(128) unsigned char smem[];
  float* H = reinterpret_cast<float*>(smem + WST * STAGE_BYTES);             // [128][LDH]
  float* Xt = H + TCM * LDH;                                                 // [8 warps][32][LDX] exchange tiles
  float* W1all = Xt + 2 * NGENW * 32 * LDX;                                  // [3][14][208] the three layer-1 kernels + bias rows, resident
  float* Pl2 = W1all + 3 * 14 * 208;                                         // [2][ENVS][208] pooled Pareto embedding of the item / the next item
  float* Xr2 = Pl2 + 2 * ENVS * 208;                                         // [2][128][13] raw x_n rows of the item / the next item
  float* AnT = Xr2 + 2 * TCM * 13;                                              // [N(j)][N(n)] shared A_n, transposed
  float* Wh = AnT + NODES * NODES;                                           // [2][201][4] head kernels, row 200 = bias
  float* Us = Wh + 2 * 201 * 4;                                              // [128][4] head pre-activations
  uint64_t* bars = reinterpret_cast<uint64_t*>(Us + TCM * 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 * MAXST + 8);

  // the shuffle tells the compiler that `warp` is warp-uniform: role branches become uniform branches and the
  // constant-bank reads of the generators go through the uniform datapath
  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  // CTAs below split_from own a full 128-row tile.  The tiles of the last, partial wave are cut into split_f pieces of
  // 128 / split_f rows, one CTA each, so that the SMs a partial wave would leave idle share its work: a CTA with fewer
  // live rows runs the same barrier protocol, but the warps of its dead 32-row groups skip their work.
  //
  // PERSISTENT: the grid is one CTA per SM and a CTA walks the work items blockIdx.x, blockIdx.x + gridDim.x, ...; TMEM,
  // barriers and the tile-independent constants are set up once, every barrier parity is derived from chunk / GEMM
  // counters that keep running across items, and the generators start on the next item while the issuer and the
  // epilogue still finish the current one.
  auto item_rows = [&](int item, int& row0, int& rows_here) {
    row0 = item * TCM; rows_here = TCM;
    if (item >= P.split_from) {
      const int j = item - P.split_from;
      rows_here = TCM / P.split_f;
      row0 = (P.split_from + j / P.split_f) * TCM + (j % P.split_f) * rows_here;
    }
  };
  const int M = P.M;
  // W ring (s < WST):  w_full  this CTA's W half landed (TMA tx)       w_peer  the peer's half landed (leader's copy)
  //                    w_empty stage consumed (tcgen05.commit, multicast to the pair)
  // A ring (s < PAST): a_full  A stage written (leader's copy; one arrive per generator warp of the pair)
  //                    a_empty stage consumed (commit)
  // acc_full[b] accumulator b complete (commit)   acc_empty[b] drained by the epilogue warps (leader's copy)
  // h_ready     the five-way sum H is complete (epilogue of GEMM 4 -> generators of GEMM 5)
  static_assert(WST <= MAXST && PAST <= MAXST, "barrier slots");
  const uint32_t w_full = smem_u32(&bars[0]), w_peer = smem_u32(&bars[MAXST]), w_empty = smem_u32(&bars[2 * MAXST]);
  const uint32_t a_full = smem_u32(&bars[3 * MAXST]), a_empty = smem_u32(&bars[4 * MAXST]);
  const uint32_t acc_full = smem_u32(&bars[5 * MAXST]), acc_empty = smem_u32(&bars[5 * MAXST + 2]), h_ready = smem_u32(&bars[5 * MAXST + 4]);
  const uint32_t x_full = smem_u32(&bars[5 * MAXST + 6]);      // [2] the item's x_n / pooled rows have landed (cp.async arrivals of all generator threads)
  const uint32_t cta_rank = (NCTA == 1) ? 0u : cluster_ctarank();
  const bool is_leader = (cta_rank == 0);

  if (warp == 0) {
    if constexpr (NCTA == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  if (tid == 32) {
    for (int s = 0; s < WST; ++s) {
      mbar_init(w_full + 8 * s, 1);
      mbar_init(w_peer + 8 * s, 1);
      mbar_init(w_empty + 8 * s, 1);
    }
    for (int s = 0; s < PAST; ++s) {
      mbar_init(a_full + 8 * s, (NGENW / 2) * NCTA);       // the eight warps of the chunk's parity
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(acc_full + 8 * b, 1);
      mbar_init(acc_empty + 8 * b, NEPIW * NCTA);
    }
    mbar_init(h_ready, NEPIW);
    mbar_init(x_full, NGENW * 32);
    mbar_init(x_full + 8, NGENW * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // item-independent constants: A_n (transposed) and the head kernels; the generators stage the per-item data
  for (int idx = tid; idx < NODES * NODES; idx += PTHREADS) AnT[(idx % NODES) * NODES + idx / NODES] = P.A_n[idx];
  for (int idx = tid; idx < 2 * 201; idx += PTHREADS) {
    const int hd = idx / 201, k = idx % 201;
    reinterpret_cast<float4*>(Wh)[idx] = (k < KH) ? __ldg(reinterpret_cast<const float4*>(P.w_head[hd] + (size_t)k * 208))
                                                  : __ldg(reinterpret_cast<const float4*>(P.b_head[hd]));
  }
  for (int idx = tid; idx < 3 * 14 * 52; idx += PTHREADS) {
    const int l1 = idx / (14 * 52), i = idx % (14 * 52);
    reinterpret_cast<float4*>(W1all)[idx] = (i < 13 * 52) ? __ldg(reinterpret_cast<const float4*>(P.w1[l1]) + i)
                                                          : __ldg(reinterpret_cast<const float4*>(P.b1[l1]) + (i - 13 * 52));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  bool ok = true;
  // programmatic dependent launch: the env-step kernel that follows in the stream may be scheduled as soon as SMs free up
  // (it waits in griddepcontrol.wait for this grid to finish before it reads the actions)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#ifdef DEBUG_TIMING
#ifndef DEBUG_BLOCK
#define DEBUG_BLOCK 100
#endif
  // per role (generator warp 0, first epilogue warp, issuer) and GEMM: 8 cycle counters
  long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long Tstart = clock64();
#define PDBG_T(x) const long long x = clock64()
#define PDBG_ACC(i, v) dbg_acc[i] += (v)
#define PDBG_FLUSH(role, g)                                                                         \
  if (blockIdx.x == DEBUG_BLOCK && lane == 0) {                                                           \
    long long* dbg = reinterpret_cast<long long*>(P.error_flag) + 16 + (role) * 64 + (g) * 8;      \
    for (int i = 0; i < 7; ++i) { dbg[i] = dbg_acc[i]; dbg_acc[i] = 0; }                             \
    dbg[7] = clock64() - Tstart;                                                                    \
  }
#else
#define PDBG_T(x)
#define PDBG_ACC(i, v)
#define PDBG_FLUSH(role, g)
#endif

  // =================================================== generators ===========================================
  auto generator_role = [&]() {
    // row group q, k-half kh of the chunk, chunk parity par: a warp works on every other chunk (8 k per thread), so
    // two chunks are always in flight in different warps and the per-chunk sync overhead is paid per 8 k
    const int q = warp & 3, kh = (warp >> 2) & 1, par = warp >> 3;
    float amax = 0.f;                                        // largest |A.X| this thread split (f16 range check)
    // asynchronous copy of one item's x_n rows and pooled rows into buffer `buf` (all 512 generator threads take part)
    auto stage_item = [&](int item_s, int buf) {
      int row0s, rows_s;
      item_rows(item_s, row0s, rows_s);
      const long long xbytes = (long long)M * 13 * 4;        // bytes of x_n that exist
      for (int idx = tid; idx < TCM * 13 / 4; idx += NGENW * 32) {
        const long long off = (long long)row0s * 13 * 4 + (long long)idx * 16;
        const long long left = xbytes - off;
        const int nbytes = left >= 16 ? 16 : (left > 0 ? (int)left : 0);
        const uint32_t dst = smem_u32(Xr2 + buf * TCM * 13 + idx * 4);
        const char* src = reinterpret_cast<const char*>(P.x_n) + (nbytes > 0 ? off : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      }
      const int env0s = row0s / NODES;
      for (int idx = tid; idx < ENVS * 52; idx += NGENW * 32) {
        const int env = env0s + idx / 52;
        const int nbytes = (env * NODES < M) ? 16 : 0;
        const uint32_t dst = smem_u32(Pl2 + buf * ENVS * 208 + idx * 4);
        const char* src = reinterpret_cast<const char*>(P.pooled) + (nbytes ? ((size_t)env * 208 * 4 + (size_t)(idx % 52) * 16) : 0);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
      }
      // the three per-environment adjacency tensors of that item are read straight from global memory (pattern masks,
      // coefficients): pull their lines into L2 now
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(x_full + 8 * buf) : "memory");
      if (row0s + TCM <= M) {
        constexpr int LINES = ENVS * NODES * NODES * 4 / 128;
        for (int idx = tid; idx < 3 * LINES; idx += NGENW * 32) {
          const int arr = idx / LINES, line = idx % LINES;
          const float* base = (arr == 0 ? P.A_s : arr == 1 ? P.A_ts : P.A_cs) + (size_t)env0s * NODES * NODES;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + 128 * line));
        }
      }
    };
    for (int item = blockIdx.x, it = 0; item < P.n_items; item += gridDim.x, ++it) {
    int row0, rows_here;
    item_rows(item, row0, rows_here);
    const int env0 = row0 / NODES;
    const uint32_t ubase = (uint32_t)it * (uint32_t)(NGEMM * NCH);
    // per-item data (x_n rows, pooled rows): double-buffered; the copies of THIS item were issued one item ago
    // (cp.async, 16 bytes each, zero-filled past the batch), so the item starts without global latency.
    // No rendezvous of the generator warps: the buffer's mbarrier completes when the copies of all 512 threads have
    // landed, and the buffer being refilled (item it-1's) is free because no warp can be more than PAST chunks behind.
    if (it == 0) {
      asm volatile("griddepcontrol.wait;" ::: "memory");      // the Pareto-branch kernel (pooled) and everything before it are done
      stage_item(item, 0);
    }
    if (item + (int)gridDim.x < P.n_items) stage_item(item + (int)gridDim.x, (it + 1) & 1);
    ok = mbar_wait(x_full + 8 * (it & 1), (uint32_t)((it >> 1) & 1)) && ok;
    const float* Xraw = Xr2 + (it & 1) * TCM * 13;
    const float* Pl = Pl2 + (it & 1) * ENVS * 208;
    if (32 * q >= rows_here) {
      // dead row group of a split tile: keep the barrier protocol going, produce nothing (the tensor core reads
      // whatever these TMEM lanes hold; rows are independent and the epilogue never looks at them)
      for (int g = 0; g < NGEMM; ++g) {
        for (int c = 0; c < NCH; ++c) {
          const uint32_t u = ubase + (uint32_t)(g * NCH + c), sa = u % PAST;
          if ((int)(u & 1u) != par) continue;
          if (u >= PAST) ok = mbar_wait(a_empty + 8 * sa, ((u / PAST) - 1) & 1) && ok;
          __syncwarp();
          if (lane == 0) {
            if (is_leader) mbar_arrive(a_full + 8 * sa);
            else mbar_arrive_remote(a_full + 8 * sa, 0);
          }
        }
      }
      continue;
    }
    const int r = 32 * q + lane;                             // row of the tile
    const int e = r / NODES, n = r % NODES;
    const int lane_env0 = lane & ~(NODES - 1);               // first lane of this row's environment inside the warp
    const bool env_valid = (env0 + e) * NODES < M;
    float* xt = Xt + warp * 2 * 32 * LDX;                    // warp-private exchange tiles: k0..k0+3 | k0+4..k0+7
    // Z = A_n . x_n (gcn_l1_1..3 share input and adjacency: formed once, kept in registers)
    float z[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) z[i] = 0.f;
    uint32_t mask_n = 0;                                     // pattern of this row of A_n
#pragma unroll 8
    for (int j = 0; j < NODES; ++j) mask_n |= (AnT[j * NODES + n] != 0.f) ? (1u << j) : 0u;
    for (uint32_t m = mask_n; m; m &= m - 1) {               // only the row's neighbours: a zero entry adds exactly nothing
      const int j = __ffs(m) - 1;
      const float a = AnT[j * NODES + n];
      const float* xr = Xraw + (e * NODES + j) * 13;
#pragma unroll
      for (int i = 0; i < 13; ++i) z[i] = fmaf(a, xr[i], z[i]);
    }
    // Sparsity pattern of this row, once per tile: the normalised adjacency A_n (adjacency + self loops) bounds the
    // pattern of A_s / A_n_ts / A_n_cs in every reference family.  If a row has more than DMAX entries, or a
    // per-environment matrix has an entry outside A_n's pattern, the warp uses the full row instead ("dense").
    float coef[DMAX];
    uint64_t nidx_packed = 0;                                // 8 neighbour lane ids, one byte each
    const int cnt = __popc(mask_n);
    uint32_t mask_o = 0;
    if (env_valid) {
      const size_t ro = ((size_t)(env0 + e) * NODES + n) * NODES;
      const float4* r0 = reinterpret_cast<const float4*>(P.A_s + ro);
      const float4* r1 = reinterpret_cast<const float4*>(P.A_ts + ro);
      const float4* r2 = reinterpret_cast<const float4*>(P.A_cs + ro);
#pragma unroll
      for (int j4 = 0; j4 < NODES / 4; ++j4) {               // independent 128-bit loads, compared afterwards
        const float4 a = __ldg(r0 + j4), b = __ldg(r1 + j4), c = __ldg(r2 + j4);
        const uint32_t m = ((a.x != 0.f) | (b.x != 0.f) | (c.x != 0.f)) | (((a.y != 0.f) | (b.y != 0.f) | (c.y != 0.f)) << 1) |
                           (((a.z != 0.f) | (b.z != 0.f) | (c.z != 0.f)) << 2) | (((a.w != 0.f) | (b.w != 0.f) | (c.w != 0.f)) << 3);
        mask_o |= m << (4 * j4);
      }
    }
    const bool dense = __any_sync(0xffffffffu, cnt > DMAX || (mask_o & ~mask_n) != 0u);
    const int dcnt = min(__reduce_max_sync(0xffffffffu, cnt), DMAX);       // neighbour slots in use (largest row of the warp)
    // slot 0 is the row itself when every row of the warp has a self loop (A_n always has): its x is still in registers
    const bool self_first = __all_sync(0xffffffffu, (mask_n >> n) & 1u);
    {
      const uint32_t rest = self_first ? (mask_n & ~(1u << n)) : mask_n;
#pragma unroll
      for (int d = 0; d < DMAX; ++d) {
        const int dd = self_first ? d - 1 : d;                // index into the remaining neighbours
        int j = n;
        if (!(self_first && d == 0) && d < cnt) j = (int)__fns(rest, 0, dd + 1);
        nidx_packed |= (uint64_t)(lane_env0 + j) << (8 * d);
      }
    }
    // adjacency row of this thread for GEMM g (element j at arow[j * astride]); rows past the batch read A_n
    auto row_of = [&](int g, const float*& rowp, int& stride) {
      const float* adj = (g == 1) ? P.A_ts : (g == 2) ? P.A_cs : (g == 3 || g == 6) ? P.A_s : nullptr;
      if (adj != nullptr && env_valid) { rowp = adj + ((size_t)(env0 + e) * NODES + n) * NODES; stride = 1; }
      else { rowp = AnT + n; stride = NODES; }
    };
    auto load_coefs = [&](int g, float* dst) {
      const float* rowp; int stride;
      row_of(g, rowp, stride);
#pragma unroll
      for (int d = 0; d < DMAX; ++d) {
        const int j = (int)((nidx_packed >> (8 * d)) & 0xff) - lane_env0;
        dst[d] = (d < cnt) ? rowp[j * stride] : 0.f;
      }
    };
    const float* arow = nullptr;
    int astride = 1;
    // Hand-off of an A stage, decoupled from the chunk that filled it: the stage is acquired (a_empty) only right before
    // its tcgen05.st, and published (wait::st, fence, arrive on a_full) in the middle of the NEXT chunk's arithmetic, so
    // neither the barrier round trip nor the tensor-memory store latency sits on the warp's dependent chain.
    int pending = -1;                                        // stage written but not yet published
    auto hand_off = [&]() {
      if (pending < 0) return;                               // warp-uniform
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(a_full + 8 * pending);
        else mbar_arrive_remote(a_full + 8 * pending, 0);
      }
      pending = -1;
    };
    auto acquire = [&](uint32_t u, uint32_t sa) {
      PDBG_T(ta);
      if (u >= PAST) ok = mbar_wait(a_empty + 8 * sa, ((u / PAST) - 1) & 1) && ok;        // chunk u-PAST consumed
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      PDBG_T(tb);
      PDBG_ACC(2, tb - ta);
    };
    for (int g = 0; g < NGEMM; ++g) {
      PDBG_T(tg0);
      // ---- per-GEMM setup ----
      load_coefs(g, coef);
      row_of(g, arow, astride);
      if (g + 1 < NGEMM) {                                   // prefetch for the next GEMM
        const float* rowp; int stride;
        row_of(g + 1, rowp, stride);
        if (stride == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(rowp));
        if (NODES == 32 && stride == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(rowp + 16));
      }
      PDBG_T(tg1);
      if (g == 5) { ok = mbar_wait(h_ready, (uint32_t)(it & 1)) && ok; }      // H complete (epilogue of GEMM 4)
      PDBG_T(tg2);
      PDBG_ACC(0, tg1 - tg0); PDBG_ACC(1, tg2 - tg1);
      // ---- this warp's chunks of the GEMM: generate, mix with the adjacency, split, store to tensor memory ----
      for (int c = 0; c < NCH; ++c) {
        const uint32_t u = ubase + (uint32_t)(g * NCH + c), sa = u % PAST;
        if ((int)(u & 1u) != par) continue;                  // the other parity's warps own this chunk
        PDBG_T(t1);
        const int k0 = c * KCH + KPT * kh;
        const uint32_t lane_sel = (uint32_t)(32 * q) << 16;
        const uint32_t col = (uint32_t)(ACOLS * sa + (KPT / 2) * kh);      // packed column of k0 inside the A stage
        if (k0 < KH) {                                       // warp-uniform; KH is a multiple of 8: all 8 k are real
          float y[KPT];
#pragma unroll
          for (int t = 0; t < KPT; ++t) y[t] = 0.f;
          if (g <= 4) {
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {                 // the two float4 halves of the thread's 8 k, one after the other
              const int kk = k0 + 4 * hf;
              float4 x;
              if (g <= 3) {
                const float* W1s = W1all + ((g == 0) ? 0 : (g <= 2 ? 1 : 2)) * 14 * 208;   // gcn_l1_1 | gcn_l1_2 (g = 1, 2) | gcn_l1_3
                x = *reinterpret_cast<const float4*>(W1s + 13 * 208 + kk);
                float4 wa[7], wb[6];                         // two batches of (broadcast) loads in flight ahead of the FMAs
#pragma unroll
                for (int i = 0; i < 7; ++i) wa[i] = *reinterpret_cast<const float4*>(W1s + i * 208 + kk);
#pragma unroll
                for (int i = 0; i < 6; ++i) wb[i] = *reinterpret_cast<const float4*>(W1s + (7 + i) * 208 + kk);
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                  x.x = fmaf(z[i], wa[i].x, x.x); x.y = fmaf(z[i], wa[i].y, x.y); x.z = fmaf(z[i], wa[i].z, x.z); x.w = fmaf(z[i], wa[i].w, x.w);
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                  x.x = fmaf(z[7 + i], wb[i].x, x.x); x.y = fmaf(z[7 + i], wb[i].y, x.y); x.z = fmaf(z[7 + i], wb[i].z, x.z); x.w = fmaf(z[7 + i], wb[i].w, x.w);
                }
                x = make_float4(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f));
              } else {
                // the reference's stack-and-reshape scramble: x14b[b,n,h] = pooled[b,(n*200+h)/N] (truss2D_RL.py:89-95)
                const float* pl = Pl + e * 208;
                const int f = n * KH + kk;
                x = make_float4(pl[f / NODES], pl[(f + 1) / NODES], pl[(f + 2) / NODES], pl[(f + 3) / NODES]);
              }
              *reinterpret_cast<float4*>(xt + (hf * 32 + lane) * LDX) = x;
            }
            PDBG_T(tx);
            PDBG_ACC(5, tx - t1);
            hand_off();                                      // publish the previous chunk's stage
            __syncwarp();
            if (!dense) {
#pragma unroll
              for (int d = 0; d < DMAX; ++d) {
                if (d >= dcnt) break;                        // warp-uniform
                const int j = (int)((nidx_packed >> (8 * d)) & 0xff);
                const float4 t0 = *reinterpret_cast<const float4*>(xt + j * LDX);
                const float4 t1v = *reinterpret_cast<const float4*>(xt + (32 + j) * LDX);
                y[0] = fmaf(coef[d], t0.x, y[0]); y[1] = fmaf(coef[d], t0.y, y[1]); y[2] = fmaf(coef[d], t0.z, y[2]); y[3] = fmaf(coef[d], t0.w, y[3]);
                y[4] = fmaf(coef[d], t1v.x, y[4]); y[5] = fmaf(coef[d], t1v.y, y[5]); y[6] = fmaf(coef[d], t1v.z, y[6]); y[7] = fmaf(coef[d], t1v.w, y[7]);
              }
            } else {
#pragma unroll 1
              for (int j = 0; j < NODES; ++j) {
                const float a = arow[j * astride];
                const float4 t0 = *reinterpret_cast<const float4*>(xt + (lane_env0 + j) * LDX);
                const float4 t1v = *reinterpret_cast<const float4*>(xt + (32 + lane_env0 + j) * LDX);
                y[0] = fmaf(a, t0.x, y[0]); y[1] = fmaf(a, t0.y, y[1]); y[2] = fmaf(a, t0.z, y[2]); y[3] = fmaf(a, t0.w, y[3]);
                y[4] = fmaf(a, t1v.x, y[4]); y[5] = fmaf(a, t1v.y, y[5]); y[6] = fmaf(a, t1v.z, y[6]); y[7] = fmaf(a, t1v.w, y[7]);
              }
            }
            __syncwarp();                                    // the tiles are rewritten by this warp's next chunk
          } else {
            // layer 3: the operand rows are rows of H (shared memory), no exchange tile needed
            hand_off();                                      // publish the previous chunk's stage
            const float* hb = H + (32 * q) * LDH + k0;
            if (!dense) {
#pragma unroll
              for (int d = 0; d < DMAX; ++d) {
                if (d >= dcnt) break;                        // warp-uniform
                const float* hr = hb + (int)((nidx_packed >> (8 * d)) & 0xff) * LDH;
                const float4 t0 = *reinterpret_cast<const float4*>(hr);
                const float4 t1v = *reinterpret_cast<const float4*>(hr + 4);
                y[0] = fmaf(coef[d], t0.x, y[0]); y[1] = fmaf(coef[d], t0.y, y[1]); y[2] = fmaf(coef[d], t0.z, y[2]); y[3] = fmaf(coef[d], t0.w, y[3]);
                y[4] = fmaf(coef[d], t1v.x, y[4]); y[5] = fmaf(coef[d], t1v.y, y[5]); y[6] = fmaf(coef[d], t1v.z, y[6]); y[7] = fmaf(coef[d], t1v.w, y[7]);
              }
            } else {
#pragma unroll 1
              for (int j = 0; j < NODES; ++j) {
                const float a = arow[j * astride];
                const float* hr = hb + (lane_env0 + j) * LDH;
                const float4 t0 = *reinterpret_cast<const float4*>(hr);
                const float4 t1v = *reinterpret_cast<const float4*>(hr + 4);
                y[0] = fmaf(a, t0.x, y[0]); y[1] = fmaf(a, t0.y, y[1]); y[2] = fmaf(a, t0.z, y[2]); y[3] = fmaf(a, t0.w, y[3]);
                y[4] = fmaf(a, t1v.x, y[4]); y[5] = fmaf(a, t1v.y, y[5]); y[6] = fmaf(a, t1v.z, y[6]); y[7] = fmaf(a, t1v.w, y[7]);
              }
            }
          }
          PDBG_T(t2);
          PDBG_ACC(3, t2 - t1);
          // y = hi + lo in fp16, two k per 32-bit column (even k in the low half)
          uint32_t hi[KPT / 2], lo[KPT / 2];
#pragma unroll
          for (int t = 0; t < KPT / 2; ++t) {
            amax = fmaxf(amax, fmaxf(fabsf(y[2 * t]), fabsf(y[2 * t + 1])));
            const __half2 h = __floats2half2_rn(y[2 * t], y[2 * t + 1]);
            const float2 f = __half22float2(h);
            hi[t] = h2_bits(h);
            lo[t] = h2_bits(__floats2half2_rn(y[2 * t] - f.x, y[2 * t + 1] - f.y));
          }
          acquire(u, sa);
          tmem_st4u(tmem_base + lane_sel + (uint32_t)TM_AHI + col, hi);
          tmem_st4u(tmem_base + lane_sel + (uint32_t)TM_ALO + col, lo);
          PDBG_T(t4);
          PDBG_ACC(6, t4 - t2);
        } else {
          // k0 == 200 (tail chunk, upper k-half): column 200 of A is the constant 1 that multiplies the bias row of the
          // W image, columns 201..207 are zero
          const uint32_t one[4] = {0x00003C00u, 0u, 0u, 0u}, zero[4] = {0u, 0u, 0u, 0u};
          hand_off();
          acquire(u, sa);
          tmem_st4u(tmem_base + lane_sel + (uint32_t)TM_AHI + col, one);
          tmem_st4u(tmem_base + lane_sel + (uint32_t)TM_ALO + col, zero);
        }
        pending = (int)sa;
        PDBG_T(t3);
        PDBG_ACC(4, t3 - t1);
      }
      hand_off();                                            // the accumulator of this GEMM must not wait for the next one
      if (warp == 0 && it == 0) { PDBG_FLUSH(0, g); }
    }
    }  // items
    if (!(amax <= F16_MAX) && P.error_flag) atomicOr(P.error_flag, 2);   // |A.X| left the fp16 range (or NaN input)
  };
  if (warp < NGENW) {
    generator_role();
  } else if (warp < NGENW + NEPIW) {
    // =================================================== epilogue =============================================
    const int q = warp & 3;
    const int r = 32 * q + lane;
    const int n = r % NODES;
    const int lane_env0 = lane & ~(NODES - 1);
    for (int item = blockIdx.x, it = 0; item < P.n_items; item += gridDim.x, ++it) {
    int row0, rows_here;
    item_rows(item, row0, rows_here);
    const bool live = 32 * q < rows_here;                    // dead row group of a split tile: barriers only
    for (int g = 0; g < NGEMM; ++g) {
      const int G = it * NGEMM + g, b = G & 1;               // GEMM counter across items: accumulator and parity
      if (!live) {
        ok = mbar_wait(acc_full + 8 * b, (uint32_t)((G >> 1) & 1)) && ok;
        __syncwarp();
        if (lane == 0) {
          if (is_leader) mbar_arrive(acc_empty + 8 * b);
          else mbar_arrive_remote(acc_empty + 8 * b, 0);
          if (g == 4) mbar_arrive(h_ready);
        }
        continue;
      }
      PDBG_T(t0);
      ok = mbar_wait(acc_full + 8 * b, (uint32_t)((G >> 1) & 1)) && ok;
      PDBG_T(t1);
      PDBG_ACC(0, t1 - t0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tacc = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(b ? TM_ACC1 : 0);
      const float wsi = P.wscale_inv[g];                     // undoes the power-of-two scale folded into W (exact)
      float* hrow = H + r * LDH;
      const float* wh = Wh + (g >= 5 ? g - 5 : 0) * 201 * 4; // only used for g >= 5
      float u0 = 0.f, u1 = 0.f, u2 = 0.f;
      uint32_t vr[2][8];
      tmem_ld8_async(tacc, vr[0]);
#pragma unroll 2
      for (int cb = 0; cb < KH / 8; ++cb) {
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float v[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = __uint_as_float(vr[cb & 1][t]);
        if (cb + 1 < KH / 8) tmem_ld8_async(tacc + (uint32_t)(8 * (cb + 1)), vr[(cb + 1) & 1]);   // in flight while cb is processed
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = fmaxf(v[t] * wsi, 0.f);   // the bias rode along as row 200 of the W image
        if (g <= 4) {
          float4* dst = reinterpret_cast<float4*>(hrow + 8 * cb);
          float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
          if (g > 0) { o0 = dst[0]; o1 = dst[1]; }
          dst[0] = make_float4(v[0] + o0.x, v[1] + o0.y, v[2] + o0.z, v[3] + o0.w);
          dst[1] = make_float4(v[4] + o1.x, v[5] + o1.y, v[6] + o1.z, v[7] + o1.w);
        } else {
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            const float4 w = *reinterpret_cast<const float4*>(wh + (8 * cb + t) * 4);
            u0 = fmaf(v[t], w.x, u0); u1 = fmaf(v[t], w.y, u1); u2 = fmaf(v[t], w.z, u2);
          }
        }
      }
      PDBG_T(t2);
      PDBG_ACC(1, t2 - t1);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (is_leader) mbar_arrive(acc_empty + 8 * b);
        else mbar_arrive_remote(acc_empty + 8 * b, 0);
        if (g == 4) mbar_arrive(h_ready);
      }
      if (g >= 5) {
        // output heads (truss2D_RL.py:121-125): sigmoid(A_n (x3 W4) + b4); the 16/32 rows of an environment are
        // lanes of this warp
        const int hd = g - 5, nout = 2 + hd;
        *reinterpret_cast<float4*>(Us + r * 4) = make_float4(u0, u1, u2, 0.f);
        __syncwarp();
        float o[3] = {wh[KH * 4 + 0], wh[KH * 4 + 1], wh[KH * 4 + 2]};
        for (int j = 0; j < NODES; ++j) {
          const float a = AnT[j * NODES + n];
          const float4 uj = *reinterpret_cast<const float4*>(Us + (32 * q + lane_env0 + j) * 4);
          o[0] = fmaf(a, uj.x, o[0]); o[1] = fmaf(a, uj.y, o[1]); o[2] = fmaf(a, uj.z, o[2]);
        }
        __syncwarp();
        const int row = row0 + r;
        if (row < M) {
          float* dst = (hd == 0 ? P.geo : P.topo) + (size_t)row * nout;
          uint64_t seed = P.seed, call = P.call;
          if (P.noise && P.seed_call) { seed = P.seed_call[0]; call += P.seed_call[1]; }
          for (int t = 0; t < nout; ++t) {
            float v = 1.f / (1.f + expf(-o[t]));
            if (P.noise) v = ou_step(v, P.mu, P.theta, P.sigma, seed, call, (uint64_t)(hd + 1), (uint64_t)row * nout + t);
            dst[t] = v;
          }
        }
      }
      PDBG_T(t3);
      PDBG_ACC(2, t3 - t2);
      if (warp == NGENW && it == 0) { PDBG_FLUSH(1, g); }
    }
    }  // items
  } else if (warp == W_ISSUER) {
    // =================================================== MMA issuer / W forwarder =============================
    if (lane == 0) {
      if (is_leader) {
        uint64_t db[WST][2];                                 // [stage][hi, lo]: one K = 16 step spans the chunk's two core-matrix columns
#pragma unroll
        for (int st = 0; st < WST; ++st) {
          const uint32_t b_hi = smem_u32(smem + st * STAGE_BYTES);
          db[st][0] = make_desc(b_hi, B_LBO);
          db[st][1] = make_desc(b_hi + NKB * B_LBO, B_LBO);
        }
        for (int item = blockIdx.x, it = 0; item < P.n_items; item += gridDim.x, ++it) {
        const uint32_t ubase = (uint32_t)it * (uint32_t)(NGEMM * NCH);
        for (int g = 0; g < NGEMM; ++g) {
          const int G = it * NGEMM + g, b = G & 1;
          const uint32_t dacc = tmem_base + (uint32_t)(b ? TM_ACC1 : 0);
          PDBG_T(ta0);
          if (G >= 2) {                                      // the epilogue of GEMM G-2 has drained this accumulator
            ok = mbar_wait_cluster(acc_empty + 8 * b, (uint32_t)(((G >> 1) - 1) & 1)) && ok;
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          }
          PDBG_T(ta1);
          PDBG_ACC(3, ta1 - ta0);
          for (int c = 0; c < NCH; ++c) {
            const uint32_t u = ubase + (uint32_t)(g * NCH + c), sw = u % WST, sa = u % PAST;
            PDBG_T(t0);
            if constexpr (NCTA == 2) ok = mbar_wait_cluster(a_full + 8 * sa, (u / PAST) & 1) && ok;
            else ok = mbar_wait(a_full + 8 * sa, (u / PAST) & 1) && ok;
            PDBG_T(t1);
            ok = mbar_wait(w_full + 8 * sw, (u / WST) & 1) && ok;
            if constexpr (NCTA == 2) ok = mbar_wait_cluster(w_peer + 8 * sw, (u / WST) & 1) && ok;
            PDBG_T(t2);
            PDBG_ACC(0, t1 - t0); PDBG_ACC(1, t2 - t1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = tmem_base + (uint32_t)(TM_AHI + ACOLS * sa), a_lo = tmem_base + (uint32_t)(TM_ALO + ACOLS * sa);
#pragma unroll
            for (int st = 0; st < WST; ++st) {
              if (st != (int)sw) continue;                   // compile-time stage index keeps the descriptors in registers
              mma_split<NCTA>(dacc, a_hi, db[st][0], c != 0);  // Yhi.Whi + Yhi.Wlo + Ylo.Whi (the tail chunk is zero-padded)
              mma_split<NCTA>(dacc, a_hi, db[st][1], 1);
              mma_split<NCTA>(dacc, a_lo, db[st][0], 1);
            }
            mma_commit<NCTA>(w_empty + 8 * sw);
            mma_commit<NCTA>(a_empty + 8 * sa);
            if (c + 1 == NCH) mma_commit<NCTA>(acc_full + 8 * b);
            PDBG_T(t3);
            PDBG_ACC(2, t3 - t2);
          }
          if (it == 0) { PDBG_FLUSH(2, g); }
        }
        }  // items
      } else {
        for (int item = blockIdx.x, it = 0; item < P.n_items; item += gridDim.x, ++it)
        for (int uu = 0; uu < NGEMM * NCH; ++uu) {           // peer CTA: tell the leader that W chunk u has landed here
          const uint32_t u = (uint32_t)it * (uint32_t)(NGEMM * NCH) + (uint32_t)uu, sw = u % WST;
          ok = mbar_wait(w_full + 8 * sw, (u / WST) & 1) && ok;
          mbar_arrive_remote(w_peer + 8 * sw, 0);
        }
      }
    }
    __syncwarp();
  } else {
    // =================================================== W producer ===========================================
    if (lane == 0) {
      for (int item = blockIdx.x, it = 0; item < P.n_items; item += gridDim.x, ++it)
      for (int g = 0; g < NGEMM; ++g)
        for (int c = 0; c < NCH; ++c) {
          const uint32_t u = (uint32_t)it * (uint32_t)(NGEMM * NCH) + (uint32_t)(g * NCH + c), s = u % WST;
          if (u >= WST) ok = mbar_wait(w_empty + 8 * s, ((u / WST) - 1) & 1) && ok;       // chunk u-WST consumed
          const uint32_t bytes = (uint32_t)STAGE_BYTES;                                    // this CTA's half: hi then lo
          const unsigned char* src = reinterpret_cast<const unsigned char*>(P.wimg[g]) +
                                     (size_t)c * Cfg<NCTA>::CHUNK_IMG_BYTES + (size_t)cta_rank * bytes;
          mbar_expect_tx(w_full + 8 * s, bytes);
          bulk_g2s(smem_u32(smem + s * STAGE_BYTES), src, bytes, w_full + 8 * s);
        }
    }
    __syncwarp();
  }

  if (!ok && P.error_flag) atomicOr(P.error_flag, 1);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();               // neither CTA leaves while the pair's TMEM / barriers are in use
  if (warp == 0) {
    if constexpr (NCTA == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

}  // namespace pipe
}  // namespace tc
}  // namespace tactor
