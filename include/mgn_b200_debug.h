/* mgn_b200_debug.h -- profiling hooks of libmgn_b200.so.  NOT part of the drop-in surface and NOT in the product
 * library: they exist only when the library is built with -DMGN_DEBUG_HOOKS
 * (MGN_NVCC_EXTRA="-DMGN_DEBUG_HOOKS" python -m modulus_b200.build), which tools/prof_kernels.py does for the per-phase
 * cycle breakdowns under profiles/.  They are process-global switches, which is why the product build leaves them out. */
#ifndef MGN_B200_DEBUG_H_
#define MGN_B200_DEBUG_H_
#ifdef __cplusplus
extern "C" {
#endif
/* dev_buf = 96 x int64 that CTA 0 of subsequent launches of the named kernel fills with per-role, per-phase cycle
 * counts; NULL disables. */
int mgn_debug_set_fwd2_timing(void* dev_buf);
int mgn_debug_set_bwd_timing(void* dev_buf);
int mgn_debug_set_edge_bwd2_timing(void* dev_buf);
#ifdef __cplusplus
}
#endif
#endif /* MGN_B200_DEBUG_H_ */
