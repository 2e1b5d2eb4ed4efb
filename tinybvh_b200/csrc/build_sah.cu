// tinybvh_b200/csrc/build_sah.cu - binned-SAH builder (placeholder until the kernels land).
#include "common.cuh"
int build_sah_launch( tbvh_bvh b, float c_trav, float c_int )
{
	tbvh_set_error( "GPU binned-SAH build not implemented yet" );
	return TBVH_E_UNSUPPORTED;
}
