// Constant per-family tables: built once on the host (tfem_family.cc), copied to the device and staged
// into shared memory by every CTA.  Restates the reference's generator output as flat arrays:
//   mesh / supports / loads / targets .. truss2D_GEN.py:181-190, 241-434
//   DOF numbering ....................... FEM_2Dtruss.py:227-261
//   symmetry lists ...................... truss2D_ENV.py (small :460-553, large :460-673)
#pragma once
#include <stdint.h>
#include <vector>
#include <string>

#include "../../include/tfem.h"

namespace tfem {

constexpr int MAXN = 2 * TFEM_MAX_NX;            // 32 nodes
constexpr int MAXE = 5 * TFEM_MAX_NX - 4;        // 76 elements
constexpr int MAXADJ = 6;                        // max elements meeting at a node (5) + sentinel

// ---- value pool ------------------------------------------------------------------------------------
// Every float32 output tensor is written as out[f] = pool[map[f]]: `pool` is a per-warp shared-memory
// array holding each distinct value once, `map` a per-family uint16 table.
//   pool[0] = 0, pool[1] = 1
//   constant part  : 6 normalised x_n columns (0,2,3,4,5,6) and 6 raw columns, N entries each
//   dynamic part   : 7 normalised x_n columns (1,7,8,9,10,11,12), 6 raw columns (1,7,8,9,10,11),
//                    10 element columns (sec, A, L, tens, comp, q, viol, A_s, ts, cs)
struct PoolLayout {
  int n, e;
  __host__ __device__ int xn_const(int k) const { return 2 + k * n; }                 // k in 0..5
  __host__ __device__ int raw_const(int k) const { return 2 + (6 + k) * n; }          // k in 0..5
  __host__ __device__ int dyn_base() const { return 2 + 12 * n; }
  __host__ __device__ int xn_dyn(int k) const { return dyn_base() + k * n; }          // k in 0..6
  __host__ __device__ int raw_dyn(int k) const { return dyn_base() + (7 + k) * n; }   // k in 0..5
  __host__ __device__ int el(int k) const { return dyn_base() + 13 * n + k * e; }     // k in 0..9
  __host__ __device__ int size() const { return dyn_base() + 13 * n + 10 * e; }
};

enum ElCol { EL_SEC = 0, EL_A, EL_L, EL_TENS, EL_COMP, EL_Q, EL_VIOL, EL_AS, EL_TS, EL_CS };

// Flat POD copied verbatim to the device.
struct alignas(16) FamilyTables {
  int32_t nx, N, E, ndof, nres, truss_type, symmetry, pad0;
  // scalars ("weak" python numbers of the reference, kept in float64)
  double y_max, y_min, d_min, ymax_minus_dmin, max_def, young, allow, load_y;
  float maxdef32, int_obj1, int_obj2, pad1;
  double sec_area[TFEM_NSEC];        // truss[s][0]*1e-4
  float sec_area32[TFEM_NSEC];       // float32(area)
  float sec_as32[TFEM_NSEC];         // float32(area / max area)   (truss2D_ENV.py:89-90)
  float pad2[2];
  double x[MAXN];                    // node x
  double y0[MAXN];                   // generated y
  double target[MAXN];               // tar_y on top nodes
  double fy[MAXN];                   // load on the node's y DOF
  int16_t dof[MAXN][2];              // 1-based reference DOF id
  uint8_t res[MAXN];                 // bit0 = x restrained, bit1 = y restrained
  uint8_t top[MAXN];
  uint8_t pair[MAXN];
  uint8_t loaded[MAXN];
  int8_t sym_src[2][MAXN];           // [coin][node]
  int8_t conn[MAXE][2];
  int8_t sym_elem[MAXE];
  int8_t adj[MAXN][MAXADJ];          // adjacent elements, -1 terminated
  int16_t react_slot[MAXN][2];       // index into reactions[] for restrained DOFs, -1 otherwise
  float pool_const[2 + 12 * MAXN];   // initial content of the constant part of the pool
  // output maps (offsets into `maps`, counted in uint16 entries)
  int32_t map_xn, map_as, map_ts, map_cs, map_rawn, map_rawe, map_total, pad3;
};
static_assert(sizeof(FamilyTables) % 16 == 0, "FamilyTables is staged with 16-byte copies");

struct Family {
  tfem_family_desc desc;
  FamilyTables t;
  std::vector<uint16_t> maps;        // concatenated output maps
  std::vector<int32_t> conn, tnsc, res, top, pair, loaded, sym_src, sym_elem;
  std::vector<double> loadvec, x, y0, target;
  std::vector<float> A_n, mask, nC_e;
  std::string error;
};

// returns false (and fills f.error) on unsupported descriptions
bool build_family(const tfem_family_desc& d, Family& f);

// numpy's float32 pairwise sum for n <= 128 (8 accumulators), see numpy/core/src/umath/loops_utils.h
float pairwise_sum_f32(const float* a, int n);

}  // namespace tfem
