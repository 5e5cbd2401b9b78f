// routing_fused2_wide.cu -- the packed cluster-resident routing kernels of routing_fused2.cu compiled a second
// time with 512 threads per CTA: one CTA of 16 warps per SM holding 8 192 votes, clusters of up to 4 (see the
// note at the end of routing_fused2.cu).  Exports qec_fused2_wide_fwd / qec_fused2_wide_bwd, which the public
// entry points qec_routing_fused_* call for capsules that would otherwise need an 8-CTA cluster.
#define QEC_PT 512
#define QEC_VARIANT_NS pt512
#define QEC_WIDE 1
#include "routing_fused2.cu"
