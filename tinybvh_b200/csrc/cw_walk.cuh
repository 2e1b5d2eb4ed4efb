// tinybvh_b200/csrc/cw_walk.cuh - the node step of the CWBVH traversal kernels (trace_cwbvh.cu: k_trace_wide; trace_tlas.cu: the
// CWBVH BLAS walk of the two-level kernel): one 160-byte traversal node (layout: trace_cwbvh.cu, cw_make_trav) against one ray ->
// the node's hit word.  Semantics: BVH8_CWBVH::Intersect, tiny_bvh.h:7046-7154.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

#define CW_STACK 128            // node groups a ray can have pending: one per level of the wide tree (the reference's own limit, tiny_bvh.h:7048)
#define CW_NODE_F4 10           // float4 per traversal node
// Quantised planes in the traversal nodes: 0 = half pairs (widened by two HADD2.F32), 1 = bfloat16 pairs (0..255 is exact in 8
// significant bits; widening is one IMAD shift for the low half and one mask for the high half).  Measured on Bistro, 16.8 M rays:
// half 5.13 / 12.55 / 0.92 Grays/s (camera / shadow / bounce), bfloat16 4.87 / 11.91 / 0.90 - the half form stays.
#ifndef CW_PLANES_BF16
#define CW_PLANES_BF16 0
#endif

__device__ __forceinline__ float2 widen( const uint32_t h2 )
{
#if CW_PLANES_BF16
	return make_float2( __uint_as_float( h2 * 65536u ), __uint_as_float( h2 & 0xffff0000u ) );
#else
	return __half22float2( *(const __half2*)&h2 );
#endif
}

// one pair of children against one ray: `near` / `far` words already chosen by the ray's signs
__device__ __forceinline__ uint32_t pair_hits( const uint32_t wnx, const uint32_t wny, const uint32_t wnz, const uint32_t wfx, const uint32_t wfy, const uint32_t wfz,
	const uint32_t bits_a, const uint32_t bits_b, const float2 ax, const float2 ay, const float2 az, const float2 bx, const float2 by, const float2 bz, const float t )
{
	const float2 tnx = __ffma2_rn( widen( wnx ), ax, bx ), tny = __ffma2_rn( widen( wny ), ay, by ), tnz = __ffma2_rn( widen( wnz ), az, bz );
	const float2 tfx = __ffma2_rn( widen( wfx ), ax, bx ), tfy = __ffma2_rn( widen( wfy ), ay, by ), tfz = __ffma2_rn( widen( wfz ), az, bz );
	const float in_a = fmaxf( fmaxf( fmaxf( tnx.x, tny.x ), tnz.x ), 0.0f ), out_a = fminf( fminf( fminf( tfx.x, tfy.x ), tfz.x ), t );
	const float in_b = fmaxf( fmaxf( fmaxf( tnx.y, tny.y ), tnz.y ), 0.0f ), out_b = fminf( fminf( fminf( tfx.y, tfy.y ), tfz.y ), t );
	return (in_a <= out_a ? bits_a : 0u) | (in_b <= out_b ? bits_b : 0u);
}

// bits 24..31 of `w` hold inner-child hits by slot; move slot s to position s ^ o (o = 7 - octant)
__device__ __forceinline__ uint32_t slots_to_order( const uint32_t w, const uint32_t o )
{
	uint32_t top = w >> 24;
	if (o & 1u) top = ((top & 0x55u) << 1) | ((top >> 1) & 0x55u);
	if (o & 2u) top = ((top & 0x33u) << 2) | ((top >> 2) & 0x33u);
	if (o & 4u) top = ((top & 0x0fu) << 4) | (top >> 4);
	return top << 24;
}

// All child pairs of one node against one ray -> the node's hit word in traversal order (inner children in bits 24..31 by
// s ^ o, triangles in bits 0..23).  OCT < 0: the ray's own signs (per lane); OCT = 0..7: every ray of the warp has negative
// x / y / z direction components as bits 2 / 1 / 0 of OCT say - plane choice and bit order are then compile-time.
template <int OCT> __device__ __forceinline__ uint32_t node_hits( const float4* __restrict__ np, const uint32_t pairs, const bool negx, const bool negy, const bool negz, const uint32_t o,
	const float ax1, const float ay1, const float az1, const float bx1, const float by1, const float bz1, const float t )
{
	const bool nx = OCT < 0 ? negx : (OCT & 4) != 0, ny = OCT < 0 ? negy : (OCT & 2) != 0, nz = OCT < 0 ? negz : (OCT & 1) != 0;
	const float2 ax = make_float2( ax1, ax1 ), ay = make_float2( ay1, ay1 ), az = make_float2( az1, az1 );
	const float2 bx = make_float2( bx1, bx1 ), by = make_float2( by1, by1 ), bz = make_float2( bz1, bz1 );
	// All four pair records, unconditionally: a node visited on the way to a hit is almost always full (3.83 of 4 pair steps per visited
	// node on Bistro camera rays), the records behind `pairs` are zero (no bits, so whatever their planes say contributes nothing), and
	// without the four branch regions the eight loads leave together.
	(void)pairs;
	float4 A[4], B[4];
	#pragma unroll
	for (int j = 0; j < 4; j++) A[j] = __ldg( np + 2 + 2 * j ), B[j] = __ldg( np + 3 + 2 * j );
	uint32_t got = 0;
	#pragma unroll
	for (int j = 0; j < 4; j++)
	{
		const uint32_t lx = __float_as_uint( A[j].x ), ly = __float_as_uint( A[j].y ), lz = __float_as_uint( A[j].z );
		const uint32_t hx = __float_as_uint( A[j].w ), hy = __float_as_uint( B[j].x ), hz = __float_as_uint( B[j].y );
		got |= pair_hits( nx ? hx : lx, ny ? hy : ly, nz ? hz : lz, nx ? lx : hx, ny ? ly : hy, nz ? lz : hz,
			__float_as_uint( B[j].z ), __float_as_uint( B[j].w ), ax, ay, az, bx, by, bz, t );
	}
	return slots_to_order( got, OCT < 0 ? o : (uint32_t)(7 - OCT) ) | (got & 0x00ffffffu);
}
