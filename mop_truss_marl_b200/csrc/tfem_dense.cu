// Dense blocked Cholesky solve of K d = P with FP64 tensor-core (DMMA) trailing updates.
//
// BASELINE.json's north_star names this variant for the large meshes (ndof = 60).  The production path
// (tfem_kernels.cu) instead renumbers the DOFs by chord column, which makes K banded with half-bandwidth 7 for
// every num_x, and factors the band.  This file builds the named variant so that the choice rests on a
// measurement and not on an argument (scripts/solver_compare.py, profiles/): K is assembled DENSE in the
// reference's own free-DOF order (FEM_2Dtruss.py:227-261, 310-324) and factored by a right-looking blocked
// Cholesky, block size 8:
//
//   per panel p:   L11 = chol(A11)                         8x8, one warp, IEEE sqrt / divide
//                  L21 = A21 L11^-T                        one thread per row
//                  A22 -= L21 L21^T                        8x8 tiles, two mma.sync.m8n8k4.f64 (SASS DMMA) each,
//                                                          lower-triangle tiles spread over the CTA's four warps
//   then           L z = P,  L^T d = z                     column-oriented, one warp
//
// One CTA (128 threads) per environment; K (64 x 65 doubles, 33 KB) lives in shared memory.
#include <cuda_runtime.h>
#include <stdint.h>

#include "tfem_family.h"

namespace tfem {

constexpr int DN = 64;        // padded system size (ndof <= 60)
constexpr int DLD = 65;       // leading dimension (odd: conflict-free column access)
constexpr int DNB = 8;        // block size

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(128)
dense_dmma_solve_kernel(const FamilyTables* __restrict__ fam, int B, const double* __restrict__ y,
                        const int32_t* __restrict__ section, double* __restrict__ d_out, int32_t* __restrict__ status) {
  extern __shared__ __align__(16) double sm[];
  double* K = sm;                       // [DN][DLD]
  double* rhs = K + DN * DLD;           // [DN]
  __shared__ int flag;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = fam->N, E = fam->E, n = fam->ndof;
  for (int b = blockIdx.x; b < B; b += gridDim.x) {
    // ---- dense assembly in the reference's DOF order (gen_global_k / gen_ssm, FEM_2Dtruss.py:284-324) ----
    for (int i = tid; i < DN * DLD; i += 128) K[i] = 0.0;
    if (tid < DN) rhs[tid] = 0.0;
    if (tid == 0) flag = 0;
    __syncthreads();
    for (int e = tid; e < E; e += 128) {
      const int n0 = fam->conn[e][0], n1 = fam->conn[e][1];
      const double dx = fam->x[n1] - fam->x[n0], dy = y[(size_t)b * N + n1] - y[(size_t)b * N + n0];
      const double L = sqrt(dx * dx + dy * dy);
      const double c = dx / L, s = dy / L;
      const double k = fam->young * fam->sec_area[section[(size_t)b * E + e]] / L;
      const double kk[2][2] = {{k * c * c, k * c * s}, {k * c * s, k * s * s}};
      const int dof[4] = {fam->dof[n0][0] - 1, fam->dof[n0][1] - 1, fam->dof[n1][0] - 1, fam->dof[n1][1] - 1};
      for (int p = 0; p < 4; ++p)
        for (int q = 0; q < 4; ++q) {
          if (dof[p] >= n || dof[q] >= n) continue;         // restrained DOFs are numbered after the free ones
          const double v = kk[p & 1][q & 1] * (((p >> 1) == (q >> 1)) ? 1.0 : -1.0);
          atomicAdd(&K[dof[p] * DLD + dof[q]], v);
        }
    }
    for (int i = tid; i < N; i += 128) {
      const int dy_ = fam->dof[i][1] - 1;
      if (dy_ < n) rhs[dy_] = fam->fy[i];
    }
    __syncthreads();
    for (int i = n + tid; i < DN; i += 128) K[i * DLD + i] = 1.0;   // padding rows: identity
    __syncthreads();

    // ---- blocked right-looking Cholesky ----
    for (int p = 0; p < DN / DNB; ++p) {
      const int j0 = p * DNB;
      if (warp == 0) {                                      // L11 = chol(A11), unblocked, lanes = rows
        for (int j = 0; j < DNB; ++j) {
          const double djj = K[(j0 + j) * DLD + j0 + j];
          if (!(djj > 0.0) && lane == 0) flag = 1;
          const double ljj = sqrt(djj);
          __syncwarp();
          if (lane == j) K[(j0 + j) * DLD + j0 + j] = ljj;
          if (lane > j && lane < DNB) K[(j0 + lane) * DLD + j0 + j] /= ljj;
          __syncwarp();
          if (lane > j && lane < DNB) {
            const double lij = K[(j0 + lane) * DLD + j0 + j];
            for (int c2 = j + 1; c2 <= lane; ++c2) K[(j0 + lane) * DLD + j0 + c2] -= lij * K[(j0 + c2) * DLD + j0 + j];
          }
          __syncwarp();
        }
      }
      __syncthreads();
      const int m0 = j0 + DNB, m = DN - m0;                 // rows below the panel
      if (tid < m) {                                        // L21 = A21 L11^-T, one row per thread
        double* row = K + (m0 + tid) * DLD + j0;
        for (int j = 0; j < DNB; ++j) {
          double v = row[j];
          for (int c2 = 0; c2 < j; ++c2) v -= row[c2] * K[(j0 + j) * DLD + j0 + c2];
          row[j] = v / K[(j0 + j) * DLD + j0 + j];
        }
      }
      __syncthreads();
      const int mt = m / DNB;                               // trailing tiles per side
      const int ntile = mt * (mt + 1) / 2;
      for (int t = warp; t < ntile; t += 4) {               // A22 -= L21 L21^T on the FP64 tensor cores
        int ti = 0, acc = 0;
        while (acc + ti + 1 <= t) { acc += ti + 1; ++ti; }  // tile (ti, tj), tj <= ti
        const int tj = t - acc;
        const int r0 = m0 + ti * DNB, c0 = m0 + tj * DNB;
        double* cp = K + (r0 + (lane >> 2)) * DLD + c0 + 2 * (lane & 3);
        double c_0 = cp[0], c_1 = cp[1];
#pragma unroll
        for (int ks = 0; ks < DNB; ks += 4) {
          const double a = -K[(r0 + (lane >> 2)) * DLD + j0 + ks + (lane & 3)];
          const double bb = K[(c0 + (lane >> 2)) * DLD + j0 + ks + (lane & 3)];
          dmma_884(c_0, c_1, a, bb);
        }
        cp[0] = c_0; cp[1] = c_1;
      }
      __syncthreads();
    }

    // ---- L z = P, L^T d = z (column oriented, one warp) ----
    if (warp == 0) {
      for (int j = 0; j < n; ++j) {
        const double zj = rhs[j] / K[j * DLD + j];
        __syncwarp();
        if (lane == 0) rhs[j] = zj;
        for (int i = j + 1 + lane; i < n; i += 32) rhs[i] -= K[i * DLD + j] * zj;
        __syncwarp();
      }
      for (int j = n - 1; j >= 0; --j) {
        const double dj = rhs[j] / K[j * DLD + j];
        __syncwarp();
        if (lane == 0) rhs[j] = dj;
        for (int i = lane; i < j; i += 32) rhs[i] -= K[j * DLD + i] * dj;
        __syncwarp();
      }
    }
    __syncthreads();
    if (tid < n) d_out[(size_t)b * n + tid] = rhs[tid];
    if (tid == 0 && status) {
      int st = flag ? 1 : 0;
      status[b] = st;
    }
    __syncthreads();
  }
}

int dense_solve_launch(const FamilyTables* d_tables, int B, const double* y, const int32_t* section, double* d,
                       int32_t* status, cudaStream_t stream) {
  const int smem = (DN * DLD + DN) * (int)sizeof(double);
  cudaError_t e = cudaFuncSetAttribute(dense_dmma_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;                       // per call: the attribute is per device
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int grid = B < sms * 6 ? B : sms * 6;               // 6 CTAs of 33 KB per SM
  dense_dmma_solve_kernel<<<grid, 128, smem, stream>>>(d_tables, B, y, section, d, status);
  return (int)cudaGetLastError();
}

}  // namespace tfem
