/* tpareto.h -- C ABI of the batched Pareto-front bookkeeping (libtfem.so): the step on the far side of the env-step.
 *
 * For B independent environments at once, what the reference driver does per Pareto solution with
 *   utils.simple_cull(points)                        test/00_small_bridge/code/utils.py:11-217
 *   utils.union_rectangles_fastest(front, +1, -1, ref_point)                          :463-530
 * (master_DDPG_truss2D_MO.py:263-368): feasibility cut (con1 > 1 or con2 > 1), strict two-objective dominance, the
 * front sorted by obj1, its spread statistics and its hypervolume against a moving reference point.
 * Each environment holds at most TPARETO_MAX_POINTS points (the reference keeps fronts of <= 50).
 */
#ifndef TPARETO_H_
#define TPARETO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TPARETO_MAX_POINTS 64

const char* tpareto_last_error(void);

/* points    [B,P,4] float32 device: obj1, obj2, con1, con2 per point (the `point` tensors of tfem_step)
 * counts    [B] int32 device, valid points per environment (NULL = all P)
 * ref_point host double[2] (NULL = {1, 1})
 * front_idx [B,P] int32: indices of the front members in the reference's order (obj1 ascending), -1 padded
 * front_len [B]   int32: 0 when no point is feasible (the reference raises IndexError there)
 * stats     [B,5] double: max_distance, dis_distance, p_norm_inv_cd, sum_distance, std_cd (simple_cull's tuple)
 * hv        [B]   double: union_rectangles_fastest(front, ref_point)
 * Any output may be NULL.  Runs on the current device, on `stream`. */
int tpareto_front_hv(int B, int P, const float* points, const int32_t* counts, const double* ref_point,
                     int32_t* front_idx, int32_t* front_len, double* stats, double* hv, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TPARETO_H_ */
