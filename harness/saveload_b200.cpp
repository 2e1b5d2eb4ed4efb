// harness/saveload_b200.cpp - BVH::Save / Load and BVH8_CWBVH::Save / Load between the B200 engine and the reference.
// Compiled where the reference header exists (oracle/Makefile `saveload`), run on the GPU box by tests/test_harness.py:
//   1. GPU-built SBVH -> shim Save -> the REFERENCE's BVH::Load -> node / primIdx arrays identical to the download, CPU rays identical
//   2. reference BVH::BuildHQ -> reference Save -> shim Load (upload) -> GPU hits identical to the reference's CPU hits
//   3. + 4. the same for BVH8_CWBVH (bvh8Data / bvh8Tris)
// usage: saveload_b200 <scene.bin> <tmpdir>
#define TINYBVH_IMPLEMENTATION
#include "tiny_bvh.h"
#include "tinybvh_b200.hpp"
#include <cstdio>
#include <string>
#include <vector>
#include <fstream>

using namespace tinybvh;

static int fails = 0;
#define CHECK( c, ... ) do { if (!(c)) { printf( "FAIL: " __VA_ARGS__ ); printf( "\n" ); fails++; } } while (0)

int main( int argc, char** argv )
{
	if (argc < 3) { printf( "usage: saveload_b200 scene.bin tmpdir\n" ); return 2; }
	std::fstream s{ argv[1], s.binary | s.in };
	if (!s) { printf( "cannot open %s\n", argv[1] ); return 2; }
	int tris = 0;
	s.read( (char*)&tris, 4 );
	bvhvec4* verts = (bvhvec4*)tinybvh::malloc64( (size_t)tris * 3 * sizeof( bvhvec4 ) );
	s.read( (char*)verts, (size_t)tris * 48 );
	const std::string dir = argv[2];
	// a few thousand rays through the scene's bounds
	bvhvec3 lo( 1e30f ), hi( -1e30f );
	for (int i = 0; i < tris * 3; i++) lo = tinybvh_min( lo, bvhvec3( verts[i] ) ), hi = tinybvh_max( hi, bvhvec3( verts[i] ) );
	const uint32_t N = 64 * 64;
	Ray* rays = (Ray*)tinybvh_b200::malloc_pinned( N * sizeof( Ray ) ), * want = (Ray*)tinybvh::malloc64( N * sizeof( Ray ) );
	const bvhvec3 eye = (lo + hi) * 0.5f + bvhvec3( 0.1f, 0.2f, 0.05f ) * (hi - lo);
	for (uint32_t i = 0; i < N; i++)
	{
		const float u = (float)(i & 63) / 64.0f - 0.5f, v = (float)(i >> 6) / 64.0f - 0.5f;
		new (&rays[i]) Ray( eye, tinybvh_normalize( bvhvec3( u, v * 0.6f, 0.7f ) ) );
	}
	auto same_hits = [&]( const Ray* a, const Ray* b ) { uint32_t d = 0; for (uint32_t i = 0; i < N; i++) d += a[i].hit.t != b[i].hit.t || a[i].hit.prim != b[i].hit.prim; return d; };
	auto reset = [&]( Ray* r ) { for (uint32_t i = 0; i < N; i++) r[i].hit.t = 1e30f, r[i].hit.u = r[i].hit.v = 0, r[i].hit.prim = 0; };

	// ---- 1. engine -> file -> reference
	tinybvh_b200::BVH gpu;
	gpu.BuildHQ( verts, tris );
	const std::string f1 = dir + "/gpu_sbvh.bvh";
	gpu.Save( f1.c_str() );
	BVH loaded;
	CHECK( loaded.Load( f1.c_str(), verts, tris ), "reference BVH::Load rejected the engine's file" );
	const tbvh_info gi = gpu.Info();
	std::vector<BVH::BVHNode> dn( gi.used_nodes ); std::vector<uint32_t> di( gi.idx_count );
	gpu.Download( dn.data(), di.data() );
	CHECK( loaded.usedNodes == gi.used_nodes && loaded.idxCount == gi.idx_count && loaded.triCount == (uint32_t)tris, "counts differ after Load" );
	CHECK( memcmp( loaded.bvhNode, dn.data(), (size_t)gi.used_nodes * 32 ) == 0 && memcmp( loaded.primIdx, di.data(), (size_t)gi.idx_count * 4 ) == 0, "arrays differ after Load" );
	memcpy( want, rays, N * sizeof( Ray ) );
	for (uint32_t i = 0; i < N; i++) loaded.Intersect( want[i] );
	reset( rays ); gpu.Intersect( rays, N );
	CHECK( same_hits( rays, want ) == 0, "CPU walk of the loaded tree differs from the GPU walk on %u rays", same_hits( rays, want ) );
	printf( "1. engine BVH::BuildHQ -> Save -> reference BVH::Load: %u nodes, arrays identical, %u rays identical\n", loaded.usedNodes, N );

	// ---- 2. reference -> file -> engine
	BVH ref;
	ref.BuildHQ( verts, tris );
	const std::string f2 = dir + "/ref_sbvh.bvh";
	ref.Save( f2.c_str() );
	tinybvh_b200::BVH gpu2;
	CHECK( gpu2.Load( f2.c_str(), verts, (uint32_t)tris ), "shim BVH::Load rejected the reference's file" );
	reset( want ); for (uint32_t i = 0; i < N; i++) ref.Intersect( want[i] );
	reset( rays ); gpu2.Intersect( rays, N );
	CHECK( same_hits( rays, want ) == 0, "GPU walk of the reference's saved tree differs on %u rays", same_hits( rays, want ) );
	printf( "2. reference BVH::BuildHQ -> Save -> engine Load: %u rays identical\n", N );

	// ---- 3. CWBVH engine -> file -> reference
	tinybvh_b200::BVH8_CWBVH gcw;
	gcw.BuildHQ( verts, tris );
	const std::string f3 = dir + "/gpu.cwbvh";
	gcw.Save( f3.c_str() );
	BVH8_CWBVH lcw;
	CHECK( lcw.Load( f3.c_str(), (uint32_t)tris ), "reference BVH8_CWBVH::Load rejected the engine's file" );
	const tbvh_info ci = gcw.Info();
	std::vector<bvhvec4> d8( ci.used_blocks ), t8( (size_t)ci.cwbvh_tri_count * 3 );
	gcw.Download( d8.data(), t8.data() );
	CHECK( lcw.usedBlocks == ci.used_blocks && memcmp( lcw.bvh8Data, d8.data(), (size_t)ci.used_blocks * 16 ) == 0, "bvh8Data differs after Load" );
	CHECK( memcmp( lcw.bvh8Tris, t8.data(), (size_t)ci.cwbvh_tri_count * 48 ) == 0, "bvh8Tris differs after Load" );
	reset( want ); for (uint32_t i = 0; i < N; i++) lcw.Intersect( want[i] );
	reset( rays ); gcw.Intersect( rays, N );
	CHECK( same_hits( rays, want ) == 0, "CPU walk of the loaded CWBVH differs from the GPU walk on %u rays", same_hits( rays, want ) );
	printf( "3. engine BVH8_CWBVH::BuildHQ -> Save -> reference Load: %u blocks identical, %u rays identical\n", lcw.usedBlocks, N );

	// ---- 4. CWBVH reference -> file -> engine
	BVH8_CWBVH rcw;
	rcw.BuildHQ( verts, tris );
	const std::string f4 = dir + "/ref.cwbvh";
	rcw.Save( f4.c_str() );
	tinybvh_b200::BVH8_CWBVH gcw2;
	CHECK( gcw2.Load( f4.c_str(), (uint32_t)tris ), "shim BVH8_CWBVH::Load rejected the reference's file" );
	reset( want ); for (uint32_t i = 0; i < N; i++) rcw.Intersect( want[i] );
	reset( rays ); gcw2.Intersect( rays, N );
	CHECK( same_hits( rays, want ) == 0, "GPU walk of the reference's saved CWBVH differs on %u rays", same_hits( rays, want ) );
	printf( "4. reference BVH8_CWBVH::BuildHQ -> Save -> engine Load: %u rays identical\n", N );
	printf( fails ? "saveload: %d FAILURES\n" : "saveload: all round trips ok\n", fails );
	return fails ? 1 : 0;
}
