// Launch interface between the C ABI (tfem_capi.cu) and the device code (tfem_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include "tfem_family.h"

namespace tfem {

enum StepMode { MODE_STEP = 0, MODE_RESET = 1, MODE_SOLVE_ONLY = 2, MODE_GENES = 3 };

struct StepArgs {
  const FamilyTables* fam;     // device copy
  const uint16_t* maps;        // device copy of Family::maps
  int map_entries;             // padded to a multiple of 8
  int B;
  int mode;
  tfem_step_in in;             // device pointers (MODE_STEP)
  tfem_step_out out;           // device pointers
  const double* so_y;          // MODE_SOLVE_ONLY: [B,N]
  const int32_t* so_sec;       // MODE_SOLVE_ONLY: [B,E]
  float* reset_move_range;     // MODE_RESET: [B,N,2]
  const double* genes;         // MODE_GENES: [B,N+E]
  double max_height;           // MODE_GENES
  int32_t* sec_out;            // MODE_GENES: decoded sections [B,E] (optional)
  float int_obj1, int_obj2;    // > 0: normalisers of point[0], point[1] instead of the family's
};

struct LaunchInfo {
  int grid, block, smem_bytes, ctas_per_sm, grid_genes;
};

// one-time per-handle setup (opt-in shared memory, occupancy query); returns cudaError_t as int
int step_kernel_configure(int nx, int device, int map_entries, LaunchInfo* info);
int step_kernel_launch(int nx, const StepArgs& args, const LaunchInfo& info, cudaStream_t stream);

// dense blocked Cholesky with DMMA trailing updates (tfem_dense.cu); returns cudaError_t as int
int dense_solve_launch(const FamilyTables* d_tables, int B, const double* y, const int32_t* section, double* d,
                       int32_t* status, cudaStream_t stream);

}  // namespace tfem
