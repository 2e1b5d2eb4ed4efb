// tinybvh_b200/csrc/trace_cwbvh.cu - CWBVH traversal (placeholder until the kernel lands).
#include "common.cuh"
int cwbvh_trace_launch( tbvh_bvh b, const void* d_rays, uint32_t stride, void* d_hits, uint32_t hit_stride, uint32_t* d_bits, uint64_t n, bool anyhit, cudaStream_t s )
{
	tbvh_set_error( "CWBVH traversal not implemented yet" );
	return TBVH_E_UNSUPPORTED;
}
