// obca_variant.cu - one kernel variant of the batched OBCA-MPC solver per compilation:
//   nvcc -c -DKV_SYM=obca_kv_cfg3 -DKV_E=4 -DKV_T=128 -DKV_B=3 -DKV_N=20 -DKV_O=4 -DKV_R=16 -DKV_FULL=0 obca_variant.cu
// KV_E: max edges per obstacle, KV_T: threads per block, KV_B: blocks per SM, KV_N/KV_O/KV_R: horizon, obstacles and rows
// compiled in (0: generic - read from the parameter block), KV_FULL: 0 first-pass kernel, 1 recovery kernel.  The list of variants is in _lib.py.
#include "obca_kernel.cuh"

extern "C" const void* KV_SYM(void) {
  return (const void*)&obca::obca_solve_kernel<KV_E, KV_T, KV_B, KV_N, KV_O, KV_R, (KV_FULL != 0)>;
}
