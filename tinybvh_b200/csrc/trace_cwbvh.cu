// tinybvh_b200/csrc/trace_cwbvh.cu - compressed wide BVH (CWBVH, Ylitie et al. 2017) closest-hit / any-hit for sm_100a.
//
// Replaces the OpenCL kernels traverse_cwbvh / isoccluded_cwbvh (traverse_cwbvh.cl:123,343) and follows the semantics
// of the reference's own CPU walk of the same data, BVH8_CWBVH::Intersect (tiny_bvh.h:7046-7154): node group
// (child base, hit bits | imask), triangle group (triangle base, 24 hit bits), children visited in octant order
// (slot ^ octinv, highest bit first), quantised slabs  t = q * 2^e * rD + (p - O) * rD  with  cmin = max(.., 0),
// cmax = min(.., tmax),  hit iff cmin <= cmax.
// Node = 5 x float4 (80 B), triangle = 3 x float4 (48 B: v2-v0, v1-v0, v0|primIdx), byte layout in SURVEY.md 8(a).
// The triangle test is the oracle's Moeller-Trumbore (common.cuh mt_test), so a ray that ends on the same primitive as
// BVH::Intersect carries bit-identical t,u,v.
#include "common.cuh"

#define CW_STACK 48

__device__ __forceinline__ uint32_t sign_extend_s8x4( const uint32_t x )
{
	// every byte becomes 0xff when its top bit is set (prmt with the sign-replicate selector)
	uint32_t v;
	asm( "prmt.b32 %0, %1, 0x0, 0x0000BA98;" : "=r"( v ) : "r"( x ) );
	return v;
}

// (float)byte i of w without the quarter-rate I2F: splice the byte under the mantissa of 2^23 (one PRMT) and subtract 2^23
// (exact).  The node step converts 48 bytes; with I2F the conversion unit, not the issue slots, bounded the kernel.
__device__ __forceinline__ float byte_f( const uint32_t w, const int i )
{
	return __fsub_rn( __uint_as_float( __byte_perm( w, 0x4b000000u, 0x7650u | (uint32_t)i ) ), 8388608.0f );
}

// 4 children: quantised bounds words (lo/hi per axis, already swizzled by ray sign) -> hit bits
__device__ __forceinline__ uint32_t slab4( const uint32_t meta4, const uint32_t octinv4, const uint32_t lox, const uint32_t loy, const uint32_t loz,
	const uint32_t hix, const uint32_t hiy, const uint32_t hiz, const float ax, const float ay, const float az,
	const float bx, const float by, const float bz, const float tmax )
{
	const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
	const uint32_t inner_mask4 = sign_extend_s8x4( is_inner4 << 3 );
	const uint32_t bit_index4 = (meta4 ^ (octinv4 & inner_mask4)) & 0x1f1f1f1fu;
	const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
	uint32_t hitmask = 0;
	#pragma unroll
	for (int i = 0; i < 4; i++)
	{
		const float tminx = __fmaf_rn( byte_f( lox, i ), ax, bx ), tminy = __fmaf_rn( byte_f( loy, i ), ay, by ), tminz = __fmaf_rn( byte_f( loz, i ), az, bz );
		const float tmaxx = __fmaf_rn( byte_f( hix, i ), ax, bx ), tmaxy = __fmaf_rn( byte_f( hiy, i ), ay, by ), tmaxz = __fmaf_rn( byte_f( hiz, i ), az, bz );
		const float cmin = fmaxf( fmaxf( fmaxf( tminx, tminy ), tminz ), 0.0f );
		const float cmax = fminf( fminf( fminf( tmaxx, tmaxy ), tmaxz ), tmax );
		if (cmin <= cmax) hitmask |= ((child_bits4 >> (8 * i)) & 0xffu) << ((bit_index4 >> (8 * i)) & 0xffu);
	}
	return hitmask;
}

template <bool ANYHIT, bool STATS>
__global__ void __launch_bounds__( 128 ) k_trace_cwbvh( const float4* __restrict__ nodes, const float4* __restrict__ tris,
	const char* rays, const uint32_t stride, char* hits, const uint32_t hit_stride, uint32_t* __restrict__ bits, const uint64_t n,
	unsigned long long* __restrict__ stats )
{
	const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
	bool occluded = false;
	if (i < n)
	{
		const float4* rp = (const float4*)(rays + i * stride);
		const float4 ro4 = rp[0], rd4 = rp[1], rr4 = rp[2], rh4 = rp[3];
		const float ox = ro4.x, oy = ro4.y, oz = ro4.z, dx = rd4.x, dy = rd4.y, dz = rd4.z;
		const float rdx = rr4.x, rdy = rr4.y, rdz = rr4.z;
		float tmax = rh4.x, hu = rh4.y, hv = rh4.z;
		uint32_t hprim = __float_as_uint( rh4.w );
		const uint32_t octinv4 = (7u - ((dx < 0 ? 4u : 0u) | (dy < 0 ? 2u : 0u) | (dz < 0 ? 1u : 0u))) * 0x01010101u;
		uint2 stack[CW_STACK];
		int sp = 0;
		uint2 ngroup = make_uint2( 0u, 0x80000000u ), tgroup = make_uint2( 0u, 0u );
		unsigned long long nsteps = 0, ntris = 0;
		while (true)
		{
			if (ngroup.y > 0x00ffffffu)
			{
				const uint32_t hitsm = ngroup.y;
				const uint32_t bit = 31u - __clz( hitsm );
				ngroup.y &= ~(1u << bit);
				if (ngroup.y > 0x00ffffffu) stack[sp++] = ngroup;
				const uint32_t slot = (bit - 24u) ^ (octinv4 & 255u);
				const uint32_t rel = __popc( hitsm & ~(0xffffffffu << slot) );
				const float4* np = nodes + (size_t)(ngroup.x + rel) * 5;
				const float4 n0 = __ldg( np ), n1 = __ldg( np + 1 ), n2 = __ldg( np + 2 ), n3 = __ldg( np + 3 ), n4 = __ldg( np + 4 );
				if (STATS) nsteps++;
				const uint32_t n0w = __float_as_uint( n0.w );
				// exponents are signed bytes: scale = 2^e as a float bit pattern (tiny_bvh.h:7072-7074)
				const int ex = (int)(int8_t)(n0w & 0xff), ey = (int)(int8_t)((n0w >> 8) & 0xff), ez = (int)(int8_t)((n0w >> 16) & 0xff);
				const float ax = __fmul_rn( __uint_as_float( (uint32_t)(ex + 127) << 23 ), rdx );
				const float ay = __fmul_rn( __uint_as_float( (uint32_t)(ey + 127) << 23 ), rdy );
				const float az = __fmul_rn( __uint_as_float( (uint32_t)(ez + 127) << 23 ), rdz );
				const float bx = __fmul_rn( -__fsub_rn( ox, n0.x ), rdx ), by = __fmul_rn( -__fsub_rn( oy, n0.y ), rdy ), bz = __fmul_rn( -__fsub_rn( oz, n0.z ), rdz );
				const bool nx = rdx < 0, ny = rdy < 0, nz = rdz < 0;
				// words: n2.x,n2.y = qlox[0..7]; n2.z,n2.w = qloy; n3.x,n3.y = qloz; n3.z,n3.w = qhix; n4.x,n4.y = qhiy; n4.z,n4.w = qhiz
				const uint32_t qlox0 = __float_as_uint( n2.x ), qlox1 = __float_as_uint( n2.y ), qloy0 = __float_as_uint( n2.z ), qloy1 = __float_as_uint( n2.w );
				const uint32_t qloz0 = __float_as_uint( n3.x ), qloz1 = __float_as_uint( n3.y ), qhix0 = __float_as_uint( n3.z ), qhix1 = __float_as_uint( n3.w );
				const uint32_t qhiy0 = __float_as_uint( n4.x ), qhiy1 = __float_as_uint( n4.y ), qhiz0 = __float_as_uint( n4.z ), qhiz1 = __float_as_uint( n4.w );
				uint32_t hitmask = slab4( __float_as_uint( n1.z ), octinv4, nx ? qhix0 : qlox0, ny ? qhiy0 : qloy0, nz ? qhiz0 : qloz0,
					nx ? qlox0 : qhix0, ny ? qloy0 : qhiy0, nz ? qloz0 : qhiz0, ax, ay, az, bx, by, bz, tmax );
				hitmask |= slab4( __float_as_uint( n1.w ), octinv4, nx ? qhix1 : qlox1, ny ? qhiy1 : qloy1, nz ? qhiz1 : qloz1,
					nx ? qlox1 : qhix1, ny ? qloy1 : qhiy1, nz ? qloz1 : qhiz1, ax, ay, az, bx, by, bz, tmax );
				ngroup = make_uint2( __float_as_uint( n1.x ), (hitmask & 0xff000000u) | (n0w >> 24) );
				tgroup = make_uint2( __float_as_uint( n1.y ), hitmask & 0x00ffffffu );
			}
			else
			{
				tgroup = ngroup;
				ngroup = make_uint2( 0u, 0u );
			}
			while (tgroup.y != 0)
			{
				const uint32_t ti = 31u - __clz( tgroup.y );
				tgroup.y -= 1u << ti;
				const float4* tp = tris + (size_t)tgroup.x + ti * 3;
				const float4 e2 = __ldg( tp ), e1 = __ldg( tp + 1 ), v0 = __ldg( tp + 2 );
				if (STATS) ntris++;
				float t, u, v;
				if (mt_test( ox, oy, oz, dx, dy, dz, v0, e1, e2, tmax, t, u, v ))
				{
					if (ANYHIT) { occluded = true; break; }
					tmax = t, hu = u, hv = v, hprim = __float_as_uint( v0.w );
				}
			}
			if (ANYHIT && occluded) break;
			if (ngroup.y > 0x00ffffffu) continue;
			if (sp == 0) break;
			ngroup = stack[--sp];
		}
		if (!ANYHIT)
		{
			float4* hp = (float4*)(hits + i * hit_stride);
			*hp = make_float4( tmax, hu, hv, __uint_as_float( hprim ) );
		}
		if (STATS) { atomicAdd( &stats[0], nsteps ); atomicAdd( &stats[1], ntris ); }
	}
	if (ANYHIT)
	{
		const uint32_t m = __ballot_sync( 0xffffffffu, occluded );
		if ((threadIdx.x & 31) == 0 && (i & ~31ull) < n) bits[i >> 5] = m;
	}
}

int cwbvh_trace_launch( tbvh_bvh b, const void* d_rays, uint32_t stride, void* d_hits, uint32_t hit_stride, uint32_t* d_bits,
	uint64_t n, bool anyhit, cudaStream_t s )
{
	if (!b->d_cw_nodes || !b->d_cw_tris) { tbvh_set_error( "CWBVH layout not resident" ); return TBVH_E_STATE; }
	if (n == 0) return TBVH_OK;
	const uint32_t block = 128;
	const uint64_t grid = (n + block - 1) / block;
	if (grid > 0x7fffffffull) { tbvh_set_error( "ray batch too large for one launch" ); return TBVH_E_ARG; }
	if (b->stats) CUDA_TRY( cudaMemsetAsync( b->d_stats, 0, 16, s ) );
	#define LAUNCH( A, S ) k_trace_cwbvh<A, S><<<(uint32_t)grid, block, 0, s>>>( b->d_cw_nodes, b->d_cw_tris, (const char*)d_rays, stride, \
		(char*)d_hits, hit_stride, d_bits, n, b->d_stats )
	if (anyhit) { if (b->stats) LAUNCH( true, true ); else LAUNCH( true, false ); }
	else { if (b->stats) LAUNCH( false, true ); else LAUNCH( false, false ); }
	#undef LAUNCH
	LAUNCHED();
	return TBVH_OK;
}
