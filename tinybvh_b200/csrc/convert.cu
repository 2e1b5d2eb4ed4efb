// tinybvh_b200/csrc/convert.cu - layout transforms on the device.
//   make_leaf_tris   : primIdx + verts -> leaf-ordered (v0|primIdx, e1, e2) records for BVH2 traversal
//   bvh_gpu_to_bvh   : BVH_GPU (Aila-Laine 64-byte, tiny_bvh.h:1095-1105) -> child-pair traversal array
//   bvh_to_bvh_gpu   : BVH_GPU::ConvertFrom (tiny_bvh.h:4612-4655)
//   bvh_to_cwbvh     : BVH8_CWBVH::Build's conversion chain (tiny_bvh.h:5827-5834)
#include "common.cuh"

// one thread per primitive reference: 4 B index read, 3 x 16 B gathered vertex reads, 3 x 16 B coalesced writes
__global__ void k_make_leaf_tris( const float4* __restrict__ verts, const uint32_t* __restrict__ prim_idx, float4* __restrict__ out, const uint32_t idx_count )
{
	const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= idx_count) return;
	const uint32_t pi = __ldg( prim_idx + p );
	const float4 v0 = __ldg( verts + (size_t)pi * 3 ), v1 = __ldg( verts + (size_t)pi * 3 + 1 ), v2 = __ldg( verts + (size_t)pi * 3 + 2 );
	// e1 = v1 - v0, e2 = v2 - v0 exactly as IntersectTri computes them per test (tiny_bvh.h:8510)
	out[(size_t)p * 3] = make_float4( v0.x, v0.y, v0.z, __uint_as_float( pi ) );
	out[(size_t)p * 3 + 1] = make_float4( __fsub_rn( v1.x, v0.x ), __fsub_rn( v1.y, v0.y ), __fsub_rn( v1.z, v0.z ), 0.0f );
	out[(size_t)p * 3 + 2] = make_float4( __fsub_rn( v2.x, v0.x ), __fsub_rn( v2.y, v0.y ), __fsub_rn( v2.z, v0.z ), 0.0f );
}

int make_leaf_tris( tbvh_bvh b, cudaStream_t s )
{
	const uint32_t n = b->info.idx_count;
	if (b->d_leaf_tris) cudaFree( b->d_leaf_tris );
	b->d_leaf_tris = 0;
	CUDA_TRY( cudaMalloc( &b->d_leaf_tris, (size_t)n * 48 ) );
	k_make_leaf_tris<<<(n + 255) / 256, 256, 0, s>>>( b->d_verts, b->d_prim_idx, b->d_leaf_tris, n );
	LAUNCHED();
	return TBVH_OK;
}

// one thread per Aila-Laine node i; an interior node writes its children as the pair at slots 2i, 2i+1:
//   {lmin, ref(L), lmax, cnt(L)}, {rmin, ref(R), rmax, cnt(R)},  ref(c) = leaf ? firstTri : 2*c,  cnt(c) = triCount(c)
__global__ void k_bvh_gpu_to_pairs( const float4* __restrict__ g, float4* __restrict__ pairs, const uint32_t used )
{
	const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= used) return;
	const float4 n0 = g[(size_t)i * 4], n1 = g[(size_t)i * 4 + 1], n2 = g[(size_t)i * 4 + 2], n3 = g[(size_t)i * 4 + 3];
	float4 o0 = make_float4( 0, 0, 0, 0 ), o1 = o0, o2 = o0, o3 = o0;
	if (__float_as_uint( n2.w ) == 0) // interior
	{
		const uint32_t L = __float_as_uint( n0.w ), R = __float_as_uint( n1.w );
		const uint32_t lc = __float_as_uint( g[(size_t)L * 4 + 2].w ), rc = __float_as_uint( g[(size_t)R * 4 + 2].w );
		const uint32_t lr = lc ? __float_as_uint( g[(size_t)L * 4 + 3].w ) : 2 * L, rr = rc ? __float_as_uint( g[(size_t)R * 4 + 3].w ) : 2 * R;
		o0 = make_float4( n0.x, n0.y, n0.z, __uint_as_float( lr ) ), o1 = make_float4( n1.x, n1.y, n1.z, __uint_as_float( lc ) );
		o2 = make_float4( n2.x, n2.y, n2.z, __uint_as_float( rr ) ), o3 = make_float4( n3.x, n3.y, n3.z, __uint_as_float( rc ) );
	}
	pairs[(size_t)i * 4] = o0, pairs[(size_t)i * 4 + 1] = o1, pairs[(size_t)i * 4 + 2] = o2, pairs[(size_t)i * 4 + 3] = o3;
}

int bvh_gpu_to_bvh( tbvh_bvh b, uint32_t used, cudaStream_t s )
{
	float4* pairs = 0;
	CUDA_TRY( cudaMalloc( &pairs, (size_t)used * 64 ) );
	k_bvh_gpu_to_pairs<<<(used + 255) / 256, 256, 0, s>>>( b->d_nodes_gpu, pairs, used );
	LAUNCHED();
	b->d_trav = pairs;
	return TBVH_OK;
}

int bvh_to_bvh_gpu( tbvh_bvh b, cudaStream_t s ) { tbvh_set_error( "BVH -> BVH_GPU conversion not implemented yet" ); return TBVH_E_UNSUPPORTED; }
int bvh_to_cwbvh( tbvh_bvh b, cudaStream_t s ) { tbvh_set_error( "BVH -> CWBVH conversion not implemented yet" ); return TBVH_E_UNSUPPORTED; }
