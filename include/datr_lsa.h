/*
 * datr_lsa.h -- C ABI of the GPU linear-sum-assignment solver of libdatr_b200.so (sm_100a).
 *
 * Replaces scipy.optimize.linear_sum_assignment as the reference's matcher calls it (models/dino/matcher.py:91, per
 * prediction set and image on a host copy of the cost matrix): same algorithm (Crouse's shortest augmenting path method
 * as implemented in scipy's rectangular_lsap.cpp), fp64 arithmetic on the fp32 costs in the same order, same tie rule --
 * identical assignments, without the device->host synchronisation.
 *
 *   cost      fp32 cost entries in device memory
 *   problems  int64 [n_problems, 5] in device memory: {offset of the problem's entry (query 0, box 0) in `cost`, elements
 *             between consecutive queries, number of queries, number of boxes (<= queries), offset of its output in `out`}
 *   out       int64 buffer in device memory: per problem 2 * boxes entries at its offset: the matched query indices in
 *             ascending order, then the box index matched to each (what scipy returns as (row_ind, col_ind))
 *   max_queries / max_boxes: maxima over the problems (shared-memory sizing); one thread block per problem.
 * Returns 0, -1 (bad argument), -3 (CUDA error) or -4 (shape outside the kernel: boxes > queries, or too large for one block).
 */
#ifndef DATR_LSA_H_
#define DATR_LSA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int datr_lsa_solve(const float* cost, const int64_t* problems, int n_problems, int max_queries, int max_boxes, int64_t* out,
                   void* stream);
const char* datr_lsa_last_error(void);
uint64_t datr_lsa_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* DATR_LSA_H_ */
