// harness/optimize_host_check.cpp - the host half of tinybvh_b200::BVH::Optimize, checked WITHOUT a GPU.
// The shim's Optimize is: download the GPU-built tree -> FillReference( tinybvh::BVH ) -> the reference's own BVH::Optimize (tiny_bvh.h:3043,
// BVH_Verbose::Optimize :4338) -> tbvh_upload_bvh.  Download / upload are covered by the GPU tests; this program checks the part in between:
// a tinybvh::BVH filled by FillReference from the plain arrays of a tree behaves, under the reference's Optimize, exactly like the object
// the reference's builder made (same node array, same SAH cost) - i.e. FillReference sets every member the optimiser reads.
// usage: optimize_host_check <scene.bin> [iterations]
#define TINYBVH_IMPLEMENTATION
#include "tiny_bvh.h"
#include "tinybvh_b200.hpp"
#include <cstdio>
#include <fstream>
#include <vector>

using namespace tinybvh;

int main( int argc, char** argv )
{
	if (argc < 2) { printf( "usage: optimize_host_check scene.bin [iterations]\n" ); return 2; }
	std::fstream s{ argv[1], s.binary | s.in };
	if (!s) { printf( "cannot open %s\n", argv[1] ); return 2; }
	int tris = 0;
	s.read( (char*)&tris, 4 );
	bvhvec4* verts = (bvhvec4*)tinybvh::malloc64( (size_t)tris * 3 * sizeof( bvhvec4 ) );
	s.read( (char*)verts, (size_t)tris * 48 );
	const uint32_t iterations = argc > 2 ? (uint32_t)atoi( argv[2] ) : 4;
	int fails = 0;
	for (int hq = 0; hq < 2; hq++)
	{
		BVH ref;
		if (hq) ref.BuildHQ( verts, tris ); else ref.Build( verts, tris );
		// what the engine's download hands over: plain arrays + counters (the GPU tree is the reference's byte for byte, tests/test_build*_gpu.py)
		std::vector<BVH::BVHNode> nodes( ref.bvhNode, ref.bvhNode + ref.usedNodes );
		std::vector<uint32_t> idx( ref.primIdx, ref.primIdx + ref.idxCount );
		tbvh_info i;
		memset( &i, 0, sizeof( i ) );
		i.used_nodes = ref.usedNodes, i.idx_count = ref.idxCount, i.prim_count = ref.triCount;
		i.aabb_min[0] = ref.aabbMin.x, i.aabb_min[1] = ref.aabbMin.y, i.aabb_min[2] = ref.aabbMin.z;
		i.aabb_max[0] = ref.aabbMax.x, i.aabb_max[1] = ref.aabbMax.y, i.aabb_max[2] = ref.aabbMax.z;
		BVH filled;
		tinybvh_b200::BVH::FillReference( filled, i, nodes.data(), idx.data(), ref.c_trav, ref.c_int, verts, (uint32_t)tris, (uint32_t)sizeof( bvhvec4 ) );
		const float before = filled.SAHCost();
		const bool sameBefore = before == ref.SAHCost();
		ref.Optimize( iterations, false, false );
		filled.Optimize( iterations, false, false );
		const bool sameNodes = filled.usedNodes == ref.usedNodes && memcmp( filled.bvhNode, ref.bvhNode, (size_t)ref.usedNodes * 32 ) == 0;
		const bool sameIdx = filled.idxCount == ref.idxCount && memcmp( filled.primIdx, ref.primIdx, (size_t)ref.idxCount * 4 ) == 0;
		const float after = filled.SAHCost();
		printf( "%s: SAH %.4f -> %.4f after Optimize( %u ); before %s, nodes %s, primIdx %s the reference object's (%u nodes)\n", hq ? "BuildHQ" : "Build", before, after,
			iterations, sameBefore ? "identical to" : "DIFFERS from", sameNodes ? "identical to" : "DIFFER from", sameIdx ? "identical to" : "DIFFERS from", ref.usedNodes );
		fails += !(sameBefore && sameNodes && sameIdx); // (whether the cost falls is the reference optimiser's business: on the bunny a few iterations raise it)
	}
	printf( fails ? "FAILED\n" : "host half of Optimize ok\n" );
	return fails ? 1 : 0;
}
